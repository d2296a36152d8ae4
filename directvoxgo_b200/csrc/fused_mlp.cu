// fused_mlp.cu -- the rgbnet MLP (lib/dvgo.py:123-131, applied at :524-539) on the 5th-generation
// tensor cores: tcgen05.mma kind::f16 (FP16 operands, fp32 accumulators in TMEM), operands staged in
// shared memory in the canonical no-swizzle core-matrix layout (tc_common.cuh).
//
// Numerics: fp32 master weights / features / outputs; GEMM operands rounded to FP16 on the way into shared memory /
// tensor memory; products accumulated in fp32.  FP16 shares TF32's 10-bit MANTISSA (the rounding error of what the
// reference's nn.Linear does on any tensor-core GPU under its pinned PyTorch 1.8.1, where allow_tf32 defaulted to True;
// README.md:32) but NOT its range: 5 exponent bits, largest finite value 65504, subnormal below 6e-5.  Hence
//   * every conversion saturates (cvt.rn.satfinite: a value beyond +-65504 clamps instead of becoming inf) and a NaN /
//     inf reaching an output sets bit 1 of the device status word, which check_status() / sync_to_model() /
//     checkpointing raise on (the reference's fp32 path has no such range limit: mlp="torch" is the exact mode);
//   * backward operands are scaled by a power of two (grad_scale) so that small gradients stay in FP16's normal
//     range, and unscaled exactly in the fp32 epilogue.
// Stated tolerance vs the exact-fp32 path: 2e-3 abs on rgb, 5e-2 relative L2 on gradients (tests/test_gpu_mlp.py,
// incl. a 300-step convergence comparison against the fp32 mode).
#include "common.cuh"
#include "tc_common.cuh"
#include "fused_scene.cuh"
#include "k0_tiles.cuh"
#include "../../include/dvgo_b200_fused.h"

namespace dvgo {

using namespace tc;

// ----------------------------------------------------------------------------------------------------
// rgbnet:  x = [k0 features (C) | view embedding (P) | 1]  ->  relu(W1 x)  ->  relu(W2 h + b2)  ->
//          sigmoid(W3 h + b3),  hidden width 128, 3 outputs  (lib/dvgo.py:123-131, :536-539).
// One tile = 128 survivors of the sample stream.  256 threads: thread (row = 32*(warp%4)+lane,
// half = warp/4) owns one sample row of the tile and half of the 128 hidden columns (a warp can only
// read its own 32-lane quarter of TMEM, so warps w and w+4 split the columns of the same rows).
//   layer 1     : tcgen05.mma, A = X~ tile [sample][feature] (K-major, shared memory), B = weights [out][in]
//   layers 2, 3 : tcgen05.mma with the A operand (the previous layer's activations, packed fp16) in TENSOR MEMORY;
//                 layer 3 is an N = 16 MMA (3 outputs used)
//   b1 is folded into W1 as column d_in of the augmented input (the constant-1 feature), b2 into W2 as column 128.
// ----------------------------------------------------------------------------------------------------
constexpr int kHid = 128;
constexpr int kHidA = kHid + 16;  // hidden width augmented with the constant-1 column (carries b2 / db2 in bwd)
constexpr int kTile = 128;
constexpr int kMlpThreads = 256;

struct MlpW {            // fp32 master weights, torch nn.Linear layout [out][in]
  const float* W1; const float* b1; const float* W2; const float* b2; const float* W3; const float* b3;
};
struct MlpG {            // fp32 gradient accumulators, same shapes
  float* W1; float* b1; float* W2; float* b2; float* W3; float* b3;
};

__device__ __forceinline__ __half half_sat(float x) {
  return __low2half(pack2_sat(x, 0.f));
}
__device__ __forceinline__ uint4 pack8(const float* v) {
  __half2 h[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) h[q] = pack2_sat(v[2 * q], v[2 * q + 1]);
  return *reinterpret_cast<uint4*>(h);
}
__device__ __forceinline__ void unpack8(const uint4& u, float* v) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int q = 0; q < 4; ++q) { const float2 f = __half22float2(h[q]); v[2 * q] = f.x; v[2 * q + 1] = f.y; }
}

// fp32 x8 -> relu -> fp16 x8, with the max done on packed halves (HMNMX2: one instruction per two values)
__device__ __forceinline__ uint4 pack8_relu(const float* v) {
  __half2 h[4];
  const __half2 z = __float2half2_rn(0.f);
#pragma unroll
  for (int q = 0; q < 4; ++q) h[q] = __hmax2(pack2_sat(v[2 * q], v[2 * q + 1]), z);
  return *reinterpret_cast<uint4*>(h);
}
// out = (act > 0) ? v : 0 on packed halves: HSET2.GT gives 1.0/0.0, one HMUL2 applies the mask
__device__ __forceinline__ uint4 pack8_masked(const float* v, const uint4& act) {
  __half2 h[4];
  const __half2* a = reinterpret_cast<const __half2*>(&act);
  const __half2 z = __float2half2_rn(0.f);
#pragma unroll
  for (int q = 0; q < 4; ++q) h[q] = __hmul2(pack2_sat(v[2 * q], v[2 * q + 1]), __hgt2(a[q], z));
  return *reinterpret_cast<uint4*>(h);
}

// ---- survivor tiles from fp32 streams -----------------------------------------------------------------
// The rgbnet kernels take their per-survivor inputs as fp16 TILES in the operand layout (tc_common.cuh: tiles_for):
//   X~ tile  [128][K1]: cols [0,C) k0 features, [C,C+pe_stride) the ray's row of the padded view-embedding table
//                       (P embedding values, the constant 1 that carries b1, zeros), rest 0
//   dZ3 tile [128][16]: cols 0..2 = S * d_rgb * rgb (1 - rgb), rest 0
// In the fused step they are written by their producers (k0_gather / sample_grad) and pulled by one bulk copy per tile;
// the two kernels below build them from fp32 streams for callers that hold those (tests, tools, the C ABI's users).
__global__ void __launch_bounds__(256) mlp_pack_x_kernel(
    const float* __restrict__ feat, int C, const int32_t* __restrict__ s_ray, const float* __restrict__ pe,
    int pe_stride, const int32_t* __restrict__ counters, int64_t surv_cap, int K1, uint8_t* __restrict__ xt) {
  int64_t n = counters[0];
  if (n > surv_cap) n = surv_cap;
  const int units = K1 >> 2;                      // 8-byte units (4 halves) per row
  const int64_t rows = (n + 255) / 256 * 256;     // zero rows up to the end of the last tile PAIR
  const uint32_t xb = tile_bytes(kTile, K1);
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < rows * units;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t p = i / units;
    const int u = static_cast<int>(i - p * units);
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (p < n) {
      const float* __restrict__ e = pe ? pe + static_cast<int64_t>(s_ray[p]) * pe_stride : nullptr;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = 4 * u + j;
        v[j] = c < C ? __ldg(feat + p * C + c) : ((e && c < C + pe_stride) ? __ldg(e + (c - C)) : 0.f);
      }
    }
    *reinterpret_cast<uint2*>(xt + (p >> 7) * xb + tile_off(static_cast<int>(p & 127), 4 * u, K1)) =
        pack4(make_float4(v[0], v[1], v[2], v[3]));
  }
}
__global__ void __launch_bounds__(256) mlp_pack_dz_kernel(
    const float* __restrict__ rgb, const float* __restrict__ d_rgb, float scale,
    const int32_t* __restrict__ counters, int64_t surv_cap, uint8_t* __restrict__ dzt) {
  int64_t n = counters[0];
  if (n > surv_cap) n = surv_cap;
  const int64_t rows = (n + 255) / 256 * 256;
  for (int64_t p = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; p < rows;
       p += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (p < n) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float x = rgb[3 * p + k];
        z[k] = d_rgb[3 * p + k] * x * (1.f - x) * scale;
      }
    }
    uint8_t* t = dzt + (p >> 7) * tile_bytes(kTile, 16) + tile_off(static_cast<int>(p & 127), 0, 16);
    *reinterpret_cast<uint4*>(t) = pack8(z);
    *reinterpret_cast<uint4*>(t + 128) = make_uint4(0u, 0u, 0u, 0u);   // columns 8..15
  }
}


// ---- fp16 weight tiles, converted once per call by mlp_pack_weights_kernel and bulk-copied by every CTA ----
// (round 1: every CTA converted 22 K fp32 weights with scalar loads and integer divisions: 84 K cycles = 43 us of
// prologue in the forward kernel, 24 us in the backward one -- tools/mlp_timeline.py)
//   [W1~ : 128 x K1 canonical tile, column d_in = b1][W2~ : 128 x 144, column 128 = b2][W3 : 16 x 128 K-major tile]
//   [W3 column pairs for the SIMT dZ2 = dZ3 W3: pair jj -> half2 {W3[c][2jj], W3[c][2jj+1]} for c = 0,1,2, pad]
//   [b3: 3 floats + pad]
struct WPack {
  uint32_t offW2, offW3, offW3p, offW3t, offB3, total;
};
__host__ __device__ inline WPack wpack_layout(int K1) {
  WPack l;
  l.offW2 = (tile_bytes(kHid, K1) + 1023u) & ~1023u;
  l.offW3 = l.offW2 + ((tile_bytes(kHid, kHidA) + 1023u) & ~1023u);
  l.offW3p = l.offW3 + ((tile_bytes(16, kHid) + 1023u) & ~1023u);
  l.offW3t = l.offW3p + 1024u;                                          // [128 j][16] = W3^T: K-major B of dH2 = dZ3 W3
  l.offB3 = l.offW3t + ((tile_bytes(kHid, 16) + 1023u) & ~1023u);
  l.total = l.offB3 + 16u;
  return l;
}
__global__ void __launch_bounds__(256) mlp_pack_weights_kernel(MlpW w, int d_in, int K1, uint8_t* __restrict__ out) {
  const WPack l = wpack_layout(K1);
  const int n1 = kHid * K1, n2 = kHid * kHidA, n3 = 16 * kHid, n4 = (kHid / 2) * 8, n5 = kHid * 16;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n1 + n2 + n3 + n4 + n5 + 4; i += gridDim.x * blockDim.x) {
    if (i < n1) {
      const int n = i / K1, c = i % K1;
      const float v = c < d_in ? w.W1[n * d_in + c] : (c == d_in ? w.b1[n] : 0.f);
      *reinterpret_cast<__half*>(out + tile_off(n, c, K1)) = half_sat(v);
    } else if (i < n1 + n2) {
      const int k = i - n1, n = k / kHidA, c = k % kHidA;
      const float v = c < kHid ? w.W2[n * kHid + c] : (c == kHid ? w.b2[n] : 0.f);
      *reinterpret_cast<__half*>(out + l.offW2 + tile_off(n, c, kHidA)) = half_sat(v);
    } else if (i < n1 + n2 + n3) {
      const int k = i - n1 - n2, n = k / kHid, c = k % kHid;
      *reinterpret_cast<__half*>(out + l.offW3 + tile_off(n, c, kHid)) = half_sat(n < 3 ? w.W3[n * kHid + c] : 0.f);
    } else if (i < n1 + n2 + n3 + n4) {
      const int k = i - n1 - n2 - n3, jj = k >> 3, e = k & 7;      // 8 halves per pair: (c = e/2, which = e%2)
      const int c = e >> 1, j = 2 * jj + (e & 1);
      *reinterpret_cast<__half*>(out + l.offW3p + jj * 16 + e * 2) = half_sat(c < 3 ? w.W3[c * kHid + j] : 0.f);
    } else if (i < n1 + n2 + n3 + n4 + n5) {
      const int k = i - n1 - n2 - n3 - n4, j = k / 16, c = k % 16;
      *reinterpret_cast<__half*>(out + l.offW3t + tile_off(j, c, 16)) = half_sat(c < 3 ? w.W3[c * kHid + j] : 0.f);
    } else {
      const int k = i - n1 - n2 - n3 - n4 - n5;
      reinterpret_cast<float*>(out + l.offB3)[k] = k < 3 ? w.b3[k] : 0.f;
    }
  }
}
// cooperative 16-byte copy global -> shared (bytes % 16 == 0)
__device__ __forceinline__ void copy_to_smem(uint8_t* dst, const uint8_t* __restrict__ src, uint32_t bytes) {
  for (uint32_t i = threadIdx.x; i < bytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(dst)[i] = __ldg(reinterpret_cast<const uint4*>(src) + i);
}
__device__ __forceinline__ void group_sync(int ctx) {   // named barrier of one 256-thread epilogue group
  asm volatile("bar.sync %0, %1;" ::"r"(1 + ctx), "r"(256) : "memory");
}

struct MmaCtx {
  uint32_t bar;     // mbarrier shared address
  uint32_t phase;   // parity to wait for next
};
__device__ __forceinline__ void mma_wait(MmaCtx& c) {
  mbar_wait(c.bar, c.phase);
  c.phase ^= 1u;
  fence_after_sync();
}
// smem written by all threads -> visible to the tensor core; TMEM reads ordered before later MMAs.
__device__ __forceinline__ void sync_for_mma() {
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
}

// GEMM issue helpers (ONE thread).  A/B tiles in the canonical layout; `cols` = tile row length.  The
// descriptors of consecutive K steps differ only in the 14-bit start-address field, so they are built once
// and advanced with one 64-bit add per step (descriptor construction was ~1k cycles of single-thread issue
// per GEMM before).
__device__ __forceinline__ void gemm_issue(uint32_t d, uint64_t a_desc, uint32_t a_step, uint64_t b_desc,
                                           uint32_t b_step, uint32_t idesc, int ksteps, bool accumulate) {
  const uint64_t a_inc = a_step >> 4, b_inc = b_step >> 4;
#pragma unroll 4
  for (int k = 0; k < ksteps; ++k) {
    mma_f16(d, a_desc, b_desc, idesc, (accumulate || k) ? 1u : 0u);
    a_desc += a_inc;
    b_desc += b_inc;
  }
}
__device__ __forceinline__ void gemm_kk(uint32_t d, uint32_t a, int a_cols, uint32_t b, int b_cols, int N, int K,
                                        bool accumulate) {  // A K-major, B K-major
  gemm_issue(d, desc_kmajor(a, a_cols), 256u, desc_kmajor(b, b_cols), 256u, make_idesc_f16(128, N, 0, 0),
             K / kMmaK, accumulate);
}
__device__ __forceinline__ void gemm_km(uint32_t d, uint32_t a, int a_cols, uint32_t b, int b_cols, int N, int K,
                                        bool accumulate) {  // A K-major, B MN-major (B tile rows = K)
  gemm_issue(d, desc_kmajor(a, a_cols), 256u, desc_mnmajor(b, b_cols), 2u * group_stride(b_cols),
             make_idesc_f16(128, N, 0, 1), K / kMmaK, accumulate);
}
__device__ __forceinline__ void gemm_mm(uint32_t d, uint32_t a, int a_cols, uint32_t b, int b_cols, int N, int K,
                                        bool accumulate) {  // A MN-major, B MN-major (rows of both = K)
  gemm_issue(d, desc_mnmajor(a, a_cols), 2u * group_stride(a_cols), desc_mnmajor(b, b_cols),
             2u * group_stride(b_cols), make_idesc_f16(128, N, 1, 1), K / kMmaK, accumulate);
}

__device__ __forceinline__ size_t align1k(size_t x) { return (x + 1023) & ~static_cast<size_t>(1023); }

// GEMM with the A operand in tensor memory (fp16 pairs packed along K, 8 columns per K = 16 step) and B in shared
// memory.  Measured on B200 (tools/ts_probe.py): N/2 cycles per MMA (the math rate) against 36 + N/4 when A is read
// from shared memory -- the A tile (4 KB per K step) is what bounds a small-N MMA from shared memory.
__device__ __forceinline__ void gemm_ts(uint32_t d, uint32_t a_tmem, uint64_t b_desc, uint32_t b_step, uint32_t idesc,
                                        int ksteps, bool accumulate) {
  const uint64_t b_inc = b_step >> 4;
#pragma unroll 4
  for (int k = 0; k < ksteps; ++k) {
    mma_f16_ts(d, a_tmem + 8u * k, b_desc, idesc, (accumulate || k) ? 1u : 0u);
    b_desc += b_inc;
  }
}

// TMEM [this thread's row][c_begin, c_begin + 32) fp32 -> relu -> fp16 pairs -> 16 packed registers
__device__ __forceinline__ void relu_pack32(uint32_t taddr, uint32_t* u) {
  float v[32];
  tmem_ld16(taddr, v);
  tmem_ld16(taddr + 16, v + 16);
  tmem_ld_wait();
  const __half2 z = __float2half2_rn(0.f);
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const __half2 h = __hmax2(pack2_sat(v[2 * j], v[2 * j + 1]), z);
    u[j] = *reinterpret_cast<const uint32_t*>(&h);
  }
}

// ---- forward ---------------------------------------------------------------------------------------
// One tile = 128 survivors; 256 threads: thread (row = 32*(warp%4)+lane, half = warp/4) owns one sample row and 64
// of the 128 hidden columns.  The X~ tile arrives by one bulk copy (double-buffered, two tiles ahead) and the
// activations never touch shared memory: each layer's epilogue writes relu(D) as packed fp16 straight back into
// tensor memory, where it is the A operand of the next layer's MMAs (TS form):
//   layer 1 : D = X~ W1~^T           A = X~ tile in shared memory (K1 = 48: k0 | view PE | 1 carries b1)
//   layer 2 : D = [H1 | 1] W2~^T     A = H1 in TMEM, K = 144 (the constant-1 column carries b2)
//   layer 3 : D3 = H2 W3^T           A = H2 in TMEM, N = 16 (3 used): 10 cycles per MMA instead of 39 from smem and
//                                     no SIMT dot products (round 1 spent ~5 instructions per hidden unit on them)
// TMEM columns (256 per CTA, 2 CTAs per SM): [0,128) accumulator D, [128,200) H (64 packed columns + 8 for the
// constant-1 K step), [224,240) layer-3 accumulator.
// No thread of this kernel issues a global LOAD inside the loop: round 2's register-staged version lost 20-30 % of
// every tile to scoreboard-shared stalls around its prefetch registers.
constexpr uint32_t kFwdTH = 128, kFwdTD3 = 224;
constexpr int kFwdIssuer = 4;   // the MMA-issuing warp: one of the four (half 1) that have no rgb epilogue to do

__global__ void __launch_bounds__(kMlpThreads, 2) mlp_fwd_kernel(
    const uint8_t* __restrict__ xt, int32_t* __restrict__ counters, int64_t surv_cap,
    const uint8_t* __restrict__ wpack, int K1, float* __restrict__ rgb, long long* __restrict__ dbg) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[4];   // chain (L1 / L2), out3 (L3), xfull[0], xfull[1]
  __shared__ uint32_t tmem_base_s;
  __shared__ float sB3[4];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, half = warp >> 2, row = q * 32 + lane;
  int64_t count = counters[0];
  if (count > surv_cap) count = surv_cap;
  const int64_t n_tiles = (count + kTile - 1) / kTile;
  if (static_cast<int64_t>(blockIdx.x) >= n_tiles) return;  // uniform per CTA, before any allocation
  if (dbg && blockIdx.x == 0 && (tid == 0 || tid == 255)) dbg[tid ? 64 : 0] = clock64();   // timeline: kernel entry

  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sW1 = base;                                        // [128 out][K1]   K-major B of layer 1
  uint8_t* sW2 = sW1 + align1k(tile_bytes(kHid, K1));         // [128 out][144]  K-major B of layer 2 (col 128 = b2)
  uint8_t* sW3 = sW2 + align1k(tile_bytes(kHid, kHidA));      // [16 out][128]   K-major B of layer 3 (rows >= 3 zero)
  uint8_t* sXb = sW3 + align1k(tile_bytes(16, kHid));         // 2 x [128 samples][K1] (double-buffered)
  const uint32_t x_bytes = tile_bytes(kTile, K1);             // K1 * 256: a multiple of 1024

  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 256);
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bars[i]), 1);
    mbar_init_fence();
  }
  const WPack wl = wpack_layout(K1);
  copy_to_smem(sW1, wpack, wl.offW3p);      // W1~ | W2~ | W3 tiles, same offsets in shared memory as in the pack
  if (tid < 3) sB3[tid] = reinterpret_cast<const float*>(wpack + wl.offB3)[tid];
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t lane_sel = static_cast<uint32_t>(q * 32) << 16;
  const uint32_t tD = tmem, tH = tmem + kFwdTH, tD3 = tmem + kFwdTD3;
  if (half == 0) {   // K = 128..143 of H: the constant 1 (fp16 0x3C00 in the low half of column 64), written once
    const uint32_t one[8] = {0x3C00u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    tmem_st8(tH + lane_sel + 64, one);
    tmem_st_wait();
  }
  MmaCtx chain{smem_u32(&bars[0]), 0u};   // L1 / L2 completions (strictly alternating)
  MmaCtx out3{smem_u32(&bars[1]), 0u};    // L3 completions
  const uint32_t xbar = smem_u32(&bars[2]);
  uint32_t xphase = 0u;                    // bit b: parity of the next fill of X buffer b to wait for
  const uint32_t idesc128 = make_idesc_f16(128, kHid, 0, 0), idesc16 = make_idesc_f16(128, 16, 0, 0);

  int dbg_n = 1;
  auto stamp = [&]() {  // optional in-kernel timeline (tools/mlp_timeline.py): CTA 0, threads 0 and 255
    if (dbg && blockIdx.x == 0 && (tid == 0 || tid == 255) && dbg_n < 64) dbg[(tid ? 64 : 0) + dbg_n++] = clock64();
  };
  // one elected thread: arm the buffer's barrier and start the bulk copy of tile t into X buffer b
  auto load_x = [&](int64_t t, uint32_t b) {
    if (warp == kFwdIssuer) {
      if (elect_one()) {
        mbar_expect_tx(xbar + 8u * b, x_bytes);
        bulk_g2s(smem_u32(sXb) + b * x_bytes, xt + t * static_cast<int64_t>(x_bytes), x_bytes, xbar + 8u * b);
      }
      __syncwarp();
    }
  };
  // relu(D[row][64 half .. +64)) -> packed fp16 -> H columns [32 half, +32); then a CTA barrier (the MMA issuer
  // may read H / overwrite D)
  auto epilogue_to_h = [&]() {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t u[16];
      relu_pack32(tD + lane_sel + half * 64 + c * 32, u);
      tmem_st16(tH + lane_sel + half * 32 + c * 16, u);
    }
    tmem_st_wait();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
  };
  auto issue_l1 = [&](uint32_t b) {       // waits for the tile in X buffer b, then queues layer 1 on it
    if (warp == kFwdIssuer) {
      if (elect_one()) {
        mbar_wait(xbar + 8u * b, (xphase >> b) & 1u);
        gemm_kk(tD, smem_u32(sXb) + b * x_bytes, K1, smem_u32(sW1), K1, kHid, K1, false);
        mma_commit(chain.bar);
      }
      __syncwarp();
    }
    xphase ^= 1u << b;
  };

  // layer-3 epilogue of the tile at sample offset s_prev: sigmoid(D3 + b3) -> rgb (half 0 only; warp-uniform)
  auto epilogue_rgb = [&](int64_t s_prev) {
    if (half == 0) {                                         // tcgen05.ld is warp-collective
      float z[16];
      tmem_ld16(tD3 + lane_sel, z);
      tmem_ld_wait();
      if (s_prev + row < count) {
        const float z0 = z[0] + sB3[0], z1 = z[1] + sB3[1], z2 = z[2] + sB3[2];
        float* __restrict__ o = rgb + (s_prev + row) * 3;
        o[0] = __frcp_rn(1.f + expf(-z0));   // correctly rounded reciprocal == 1.f / x without the division's range fix-ups
        o[1] = __frcp_rn(1.f + expf(-z1));
        o[2] = __frcp_rn(1.f + expf(-z2));
        if (!(fabsf(z0) + fabsf(z1) + fabsf(z2) < 3.0e38f)) counters[1] = counters[1] | 2;  // NaN / inf logit
      }
    }
  };

  const int64_t step = gridDim.x;
  int64_t tile = blockIdx.x;
  load_x(tile, 0u);
  if (tile + step < n_tiles) load_x(tile + step, 1u);
  issue_l1(0u);
  uint32_t buf = 0u;
  int64_t s_prev = -1;
  for (; tile < n_tiles; tile += step, buf ^= 1u) {
    const bool has_next = tile + step < n_tiles;
    stamp();
    mma_wait(chain);                                       // L1(tile): X buffer `buf` is free again
    if (tile + 2 * step < n_tiles) load_x(tile + 2 * step, buf);
    stamp();
    epilogue_to_h();                                       // H1
    stamp();
    if (warp == kFwdIssuer) {
      if (elect_one()) {
        gemm_ts(tD, tH, desc_kmajor(smem_u32(sW2), kHidA), 256u, idesc128, kHidA / kMmaK, false);
        mma_commit(chain.bar);
      }
      __syncwarp();
    }
    stamp();
    // The previous tile's rgb epilogue runs HERE, in the shadow of this tile's layer-2 MMAs (576 cycles of tensor time
    // during which the CTA had nothing to do), by the four warps that do not issue.  It used to sit between L3 and the
    // next tile's first epilogue: ~670 cycles of every tile's dependency chain (tools/mlp_timeline.py).  L3 of THIS
    // tile, the next writer of D3, is issued after the CTA barrier of the H2 epilogue below.
    if (s_prev >= 0) {
      mma_wait(out3);                                      // L3(previous tile): long done
      epilogue_rgb(s_prev);
    }
    stamp();
    mma_wait(chain);                                       // L2(tile)
    stamp();
    epilogue_to_h();                                       // H2 (b2 came through the constant-1 K step)
    stamp();
    if (warp == kFwdIssuer) {
      if (elect_one()) {
        gemm_ts(tD3, tH, desc_kmajor(smem_u32(sW3), kHid), 256u, idesc16, kHid / kMmaK, false);
        mma_commit(out3.bar);
      }
      __syncwarp();
    }
    if (has_next) issue_l1(buf ^ 1u);                      // L1 of the next tile queues behind L3: D is free, H is not touched
    s_prev = tile * kTile;
    stamp();
  }
  mma_wait(out3);                                          // L3 of the CTA's last tile
  epilogue_rgb(s_prev);
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// ---- forward with the k0 gather fused in --------------------------------------------------------------
// mlp_fwd_kernel plus four PRODUCER warps per CTA that build the X~ tile straight into the shared-memory buffer the
// layer-1 MMA reads (the row code of k0_gather_tiles_kernel: k0_tiles.cuh): the trilinear k0 gather (8 corners x C
// floats per survivor, L1-wavefront bound: ncu l1tex data-pipe 71 % in the stand-alone kernel) runs in the shadow of
// the tensor-core chain of the previous tiles instead of in front of it, and the tiles no longer travel through HBM
// (96 B written + 96 B read per survivor; 800x800 dense render: 18 GB per frame each way).
//   producers (warps 8..11): wait xempty[b] -> rows of tile j into X buffer b = j & 1 -> fence.proxy.async -> one
//                            arrival per warp on xfull[b]
//   issuer (warp 4)        : wait xfull[b] -> [training: one bulk copy shared -> global of the finished tile, the
//                            backward kernel's input] -> layer-1 MMAs -> commit on `chain` and on xempty[b]
// The eight consumer warps run mlp_fwd_kernel's loop unchanged; their CTA-wide barriers become a named barrier of 256.
constexpr int kFgProducerWarps = 4;
constexpr int kFgThreads = kMlpThreads + 32 * kFgProducerWarps;

__device__ __forceinline__ void consumer_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(kMlpThreads) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src_saddr, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_saddr), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_s2g_wait_read() {   // the source buffer may be overwritten
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

template <int C>
__global__ void __launch_bounds__(kFgThreads, 2) mlp_fwd_gather_kernel(
    SceneArgs a, const float* __restrict__ k0, const float4* __restrict__ s_pos, const uint8_t* __restrict__ pe16,
    int pe_stride, int32_t* __restrict__ counters, int64_t surv_cap, const uint8_t* __restrict__ wpack, int K1,
    float* __restrict__ rgb, uint8_t* __restrict__ xt_out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[6];   // chain (L1 / L2), out3 (L3), xfull[0,1], xempty[0,1]
  __shared__ uint32_t tmem_base_s;
  __shared__ float sB3[4];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int64_t count = counters[0];
  if (count > surv_cap) count = surv_cap;
  const int64_t n_tiles = (count + kTile - 1) / kTile;
  // training: the backward kernel consumes tile PAIRS, so an odd tile count is completed by one all-zero tile
  const int64_t n_proc = xt_out ? 2 * ((n_tiles + 1) / 2) : n_tiles;
  if (static_cast<int64_t>(blockIdx.x) >= n_proc) return;  // uniform per CTA, before any allocation

  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sW1 = base;                                        // [128 out][K1]   K-major B of layer 1
  uint8_t* sW2 = sW1 + align1k(tile_bytes(kHid, K1));         // [128 out][144]  K-major B of layer 2 (col 128 = b2)
  uint8_t* sW3 = sW2 + align1k(tile_bytes(kHid, kHidA));      // [16 out][128]   K-major B of layer 3 (rows >= 3 zero)
  uint8_t* sXb = sW3 + align1k(tile_bytes(16, kHid));         // 2 x [128 samples][K1] (double-buffered)
  const uint32_t x_bytes = tile_bytes(kTile, K1);             // K1 * 256: a multiple of 1024

  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 256);
  if (tid == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    mbar_init(smem_u32(&bars[2]), kFgProducerWarps);
    mbar_init(smem_u32(&bars[3]), kFgProducerWarps);
    mbar_init(smem_u32(&bars[4]), 1);
    mbar_init(smem_u32(&bars[5]), 1);
    mbar_init_fence();
  }
  const WPack wl = wpack_layout(K1);
  copy_to_smem(sW1, wpack, wl.offW3p);      // W1~ | W2~ | W3 tiles, same offsets in shared memory as in the pack
  for (uint32_t i = tid; i < 2u * x_bytes / 16u; i += blockDim.x)   // padding chunks of the rows are never rewritten
    reinterpret_cast<uint4*>(sXb)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid < 3) sB3[tid] = reinterpret_cast<const float*>(wpack + wl.offB3)[tid];
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t xfull = smem_u32(&bars[2]), xempty = smem_u32(&bars[4]);
  const int64_t step = gridDim.x;

  if (warp >= kMlpThreads / 32) {
    // ---- producers ----
    constexpr int G = K0TileShape<C>::G, SPW = K0TileShape<C>::SPW;
    const SceneDev sc = load_scene(a);
    const int pw = warp - kMlpThreads / 32;
    const int q = lane % G, sub = lane / G;
    const int used_chunks = (C + pe_stride + 7) >> 3;
    uint32_t ephase = 0u, j = 0u;
    for (int64_t tile = blockIdx.x; tile < n_proc; tile += step, ++j) {
      const uint32_t b = j & 1u;
      if (j >= 2u) {                       // the layer-1 MMAs (and the copy-out) of the buffer's previous tile are done
        mbar_wait(xempty + 8u * b, (ephase >> b) & 1u);
        ephase ^= 1u << b;
      }
      uint8_t* sx = sXb + b * x_bytes;
      for (int pass = pw; pass * SPW < kTile; pass += kFgProducerWarps) {   // warp-uniform trip count
        const int row = pass * SPW + sub;
        const bool active = sub < SPW && row < kTile;
        const int64_t p = tile * kTile + row;
        const bool live = active && p < count;
        int r = 0;
        Corner8 cn;
        cn.valid = 0u;
        if (live) {
          const float4 rec = __ldg(s_pos + p);
          r = __float_as_int(rec.w);
          cn = corner8_idx(sc, rec.x, rec.y, rec.z);
        }
        k0_tile_row<C, true>(k0, cn, r, active, live, pe16, K1, used_chunks, sx + tile_off(active ? row : 0, 0, K1), q);
      }
      fence_async_smem();                  // this lane's rows -> visible to the tensor core / the bulk copy
      __syncwarp();
      if (lane == 0) mbar_arrive(xfull + 8u * b);
    }
    return;
  }

  // ---- consumers: mlp_fwd_kernel's loop ----
  const int q = warp & 3, half = warp >> 2, row = q * 32 + lane;
  const uint32_t tmem = tmem_base_s;
  const uint32_t lane_sel = static_cast<uint32_t>(q * 32) << 16;
  const uint32_t tD = tmem, tH = tmem + kFwdTH, tD3 = tmem + kFwdTD3;
  if (half == 0) {   // K = 128..143 of H: the constant 1 (fp16 0x3C00 in the low half of column 64), written once
    const uint32_t one[8] = {0x3C00u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    tmem_st8(tH + lane_sel + 64, one);
    tmem_st_wait();
  }
  MmaCtx chain{smem_u32(&bars[0]), 0u};
  MmaCtx out3{smem_u32(&bars[1]), 0u};
  uint32_t xphase = 0u;
  const uint32_t idesc128 = make_idesc_f16(128, kHid, 0, 0), idesc16 = make_idesc_f16(128, 16, 0, 0);

  auto epilogue_to_h = [&]() {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t u[16];
      relu_pack32(tD + lane_sel + half * 64 + c * 32, u);
      tmem_st16(tH + lane_sel + half * 32 + c * 16, u);
    }
    tmem_st_wait();
    fence_before_sync();
    consumer_sync();
    fence_after_sync();
  };
  auto issue_l1 = [&](uint32_t b, int64_t t) {   // waits for tile t in X buffer b, then queues layer 1 on it
    if (warp == kFwdIssuer) {
      if (elect_one()) {
        mbar_wait(xfull + 8u * b, (xphase >> b) & 1u);
        if (xt_out) bulk_s2g(xt_out + t * static_cast<int64_t>(x_bytes), smem_u32(sXb) + b * x_bytes, x_bytes);
        gemm_kk(tD, smem_u32(sXb) + b * x_bytes, K1, smem_u32(sW1), K1, kHid, K1, false);
        mma_commit(chain.bar);
        if (xt_out) bulk_s2g_wait_read();
        mma_commit(xempty + 8u * b);       // second arrival of the same completion: the producers may refill the buffer
      }
      __syncwarp();
    }
    xphase ^= 1u << b;
  };
  auto epilogue_rgb = [&](int64_t s_prev) {
    if (half == 0) {                                         // tcgen05.ld is warp-collective
      float z[16];
      tmem_ld16(tD3 + lane_sel, z);
      tmem_ld_wait();
      if (s_prev + row < count) {
        const float z0 = z[0] + sB3[0], z1 = z[1] + sB3[1], z2 = z[2] + sB3[2];
        float* __restrict__ o = rgb + (s_prev + row) * 3;
        o[0] = __frcp_rn(1.f + expf(-z0));
        o[1] = __frcp_rn(1.f + expf(-z1));
        o[2] = __frcp_rn(1.f + expf(-z2));
        if (!(fabsf(z0) + fabsf(z1) + fabsf(z2) < 3.0e38f)) counters[1] = counters[1] | 2;  // NaN / inf logit
      }
    }
  };

  int64_t tile = blockIdx.x;
  issue_l1(0u, tile);
  uint32_t buf = 0u;
  int64_t s_prev = -1;
  for (; tile < n_proc; tile += step, buf ^= 1u) {
    const bool has_next = tile + step < n_proc;
    mma_wait(chain);                                       // L1(tile)
    epilogue_to_h();                                       // H1
    if (warp == kFwdIssuer) {
      if (elect_one()) {
        gemm_ts(tD, tH, desc_kmajor(smem_u32(sW2), kHidA), 256u, idesc128, kHidA / kMmaK, false);
        mma_commit(chain.bar);
      }
      __syncwarp();
    }
    if (s_prev >= 0) {                                     // the previous tile's rgb, in the shadow of layer 2
      mma_wait(out3);
      epilogue_rgb(s_prev);
    }
    mma_wait(chain);                                       // L2(tile)
    epilogue_to_h();                                       // H2
    if (warp == kFwdIssuer) {
      if (elect_one()) {
        gemm_ts(tD3, tH, desc_kmajor(smem_u32(sW3), kHid), 256u, idesc16, kHid / kMmaK, false);
        mma_commit(out3.bar);
      }
      __syncwarp();
    }
    if (has_next) issue_l1(buf ^ 1u, tile + step);
    s_prev = tile * kTile;
  }
  mma_wait(out3);
  epilogue_rgb(s_prev);
  if (xt_out && warp == kFwdIssuer) {      // the tile copies must have landed in global memory before the CTA exits
    if (elect_one()) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncwarp();
  }
  fence_before_sync();
  consumer_sync();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// ---- backward (with forward recompute) ---------------------------------------------------------------
// Per tile: recompute H1, H2 (as in forward); dZ3 = d_rgb * rgb (1-rgb) * S; then
//   dW3^T += H2^T dZ3      (MMA, A = H2 MN-major, B = dZ3 MN-major, N = 16)
//   dH2    = dZ3 W3 (MMA, K = 16) ; dZ2 = (H2 > 0) * dH2           (epilogue, in place over H2)
//   dW2   += dZ2^T H1 ;  db2 += dZ2^T 1 ;  dH1 = dZ2 W2            (MMAs)
//   dZ1    = (H1 > 0) * dH1                                        (epilogue, in place over H1)
//   dW1~  += dZ1^T X~   (column d_in of the augmented input gives db1) ;  dX = dZ1 W1[:, :16]   (MMAs)
//   d_feat = dX[:, :C] / S
// The four weight-gradient accumulators stay in TMEM across all tiles of the CTA and are flushed to
// global memory with atomics once at the end.  S = grad_scale (a power of two) keeps the FP16
// operands of the backward GEMMs in normal range.
//
// Pipelining: the five MMA batches of a tile are separated by SIMT epilogues that depend on them, so a
// single tile leaves the tensor pipe idle ~85% of the time (measured, profiles/r01_*).  Each CTA
// therefore works on TWO tiles (contexts A and B, each with its own activation buffers, TMEM work
// columns and mbarriers) and is warp-specialised: 16 epilogue warps alternate between the two contexts
// while a dedicated issuer warp feeds the tensor core, so descriptor building / MMA issue of one
// context overlaps the epilogue of the other.  Hand-offs are mbarriers only (no __syncthreads in the
// loop): ready[ctx] (512 epilogue arrivals -> issuer may read the smem tiles / overwrite the TMEM work
// columns) and full[ctx] (tcgen05.commit -> epilogue may read TMEM / overwrite the tiles in place).
// One thread issues every MMA in program order, so both contexts can accumulate into the same TMEM
// weight-gradient columns.
constexpr int kBwdEpiThreads = 512;               // 16 epilogue warps: (row quarter q) x (32-column part)
constexpr int kBwdThreads = kBwdEpiThreads + 32;  // + one MMA-issuer warp

struct TileCtx {
  uint8_t* sX; uint8_t* sH1; uint8_t* sH2; uint8_t* sdZ3;
  uint32_t tWork;
  MmaCtx bar;     // full[ctx]: MMA batch complete
  MmaCtx ready;   // ready[ctx]: epilogue output in place
  uint32_t xbar;  // tiles[ctx]: the X~ and dZ3 tiles of the pair have landed (bulk copies)
  uint32_t xfree; // xfree[ctx]: the last batch of the pair (the only late reader of X~ / dZ3) has completed
  int64_t s0;
  bool valid;
};

// Inputs: X~ tiles and dZ3 tiles (see mlp_pack_x_kernel).  One bulk copy per tile puts them where the tensor core
// reads them; no thread issues a global load inside the loop.  (Round 2's register-staged version left the tensor
// pipe idle ~4500 of ~12300 cycles per tile pair at the pair boundary: the loads could only be issued after the last
// epilogue, because registers held across the epilogues stalled every tcgen05.wait::ld on the shared scoreboards.)
__global__ void __launch_bounds__(kBwdThreads, 1) mlp_bwd_kernel(
    const uint8_t* __restrict__ xt, const uint8_t* __restrict__ dzt, int C, int d_in,
    int32_t* __restrict__ counters, int64_t surv_cap, const uint8_t* __restrict__ wpack, int K1, float grad_scale,
    float* __restrict__ d_feat, MlpG g, long long* __restrict__ dbg) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[8];  // full[0,1], ready[0,1], tiles[0,1], xfree[0,1]
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool is_issuer = warp == kBwdEpiThreads / 32;
  const int q = warp & 3, part = (warp >> 2) & 3, row = q * 32 + lane;  // part in 0..3: 32 hidden columns each
  int64_t count = counters[0];
  if (count > surv_cap) count = surv_cap;
  const int64_t n_tiles = (count + kTile - 1) / kTile;
  const int64_t n_pairs = (n_tiles + 1) / 2;
  if (static_cast<int64_t>(blockIdx.x) >= n_pairs) return;

  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sW1 = base;
  uint8_t* sW2 = sW1 + align1k(tile_bytes(kHid, K1));
  uint8_t* sW3t = sW2 + align1k(tile_bytes(kHid, kHidA));   // [128 hidden j][16: c < 3] = W3^T, K-major B of dH2
  uint8_t* ctx_base = sW3t + align1k(tile_bytes(kHid, 16));
  const size_t ctx_bytes = align1k(tile_bytes(kTile, K1)) + align1k(tile_bytes(kTile, kHidA)) +
                           align1k(tile_bytes(kTile, kHid)) + align1k(tile_bytes(kTile, 16));
  const uint32_t x_bytes = tile_bytes(kTile, K1), dz_bytes = tile_bytes(kTile, 16);

  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 512);
  if (tid == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    mbar_init(smem_u32(&bars[2]), kBwdEpiThreads);
    mbar_init(smem_u32(&bars[3]), kBwdEpiThreads);
    for (int i = 4; i < 8; ++i) mbar_init(smem_u32(&bars[i]), 1);
    mbar_init_fence();
  }
  const WPack wl = wpack_layout(K1);
  copy_to_smem(sW1, wpack, wl.offW3);                         // W1~ | W2~ tiles (fp16, packed once per step)
  copy_to_smem(sW3t, wpack + wl.offW3t, tile_bytes(kHid, 16));
  for (int i = tid; i < kTile * 16; i += blockDim.x) {
    const int j = i / 16, c = i % 16;
    // columns 128..143 of both H1 tiles: the constant 1 (b2 rides the layer-2 GEMM, db2 the dW2 GEMM), then 0;
    // the epilogues only ever write columns < 128
    for (int cx_i = 0; cx_i < 2; ++cx_i) {
      uint8_t* h1 = ctx_base + cx_i * ctx_bytes + align1k(tile_bytes(kTile, K1));
      *reinterpret_cast<__half*>(h1 + tile_off(j, kHid + c, kHidA)) = __float2half_rn(c == 0 ? 1.f : 0.f);
    }
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t tdW2 = tmem + 256, tdW1 = tmem + 400, tdW3 = tmem + 448;  // dW2|db2: 144 cols, dW1~: 48, dW3^T: 16
  TileCtx cx[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    uint8_t* b = ctx_base + i * ctx_bytes;
    cx[i].sX = b;
    cx[i].sH1 = b + align1k(tile_bytes(kTile, K1));
    cx[i].sH2 = cx[i].sH1 + align1k(tile_bytes(kTile, kHidA));
    cx[i].sdZ3 = cx[i].sH2 + align1k(tile_bytes(kTile, kHid));
    cx[i].tWork = tmem + 128 * i;
    cx[i].bar = MmaCtx{smem_u32(&bars[i]), 0u};
    cx[i].ready = MmaCtx{smem_u32(&bars[2 + i]), 0u};
    cx[i].xbar = smem_u32(&bars[4 + i]);
    cx[i].xfree = smem_u32(&bars[6 + i]);
  }
  const float inv_scale = 1.f / grad_scale;
  const uint32_t lane_sel = static_cast<uint32_t>(q * 32) << 16;
  float db3[3] = {0.f, 0.f, 0.f};
  bool first = true;

  // ---- per-context stages (all 512 epilogue threads unless noted) ----
  // TMEM work[row][32*part, +32) -> relu(v) -> fp16 -> smem
  auto epi_relu = [&](TileCtx& c, uint8_t* sOut, int out_cols) {
    float v[32];
    tmem_ld16(c.tWork + lane_sel + part * 32, v);
    tmem_ld16(c.tWork + lane_sel + part * 32 + 16, v + 16);
    tmem_ld_wait();
#pragma unroll
    for (int cc = 0; cc < 4; ++cc)
      *reinterpret_cast<uint4*>(sOut + tile_off(row, part * 32 + cc * 8, out_cols)) = pack8_relu(v + cc * 8);
  };
  // out = (act > 0) * TMEM work, in place over the activation tile (dZ2 over H2, dZ1 over H1)
  auto epi_mask = [&](TileCtx& c, uint8_t* sAct, int act_cols) {
    float v[32];
    tmem_ld16(c.tWork + lane_sel + part * 32, v);
    tmem_ld16(c.tWork + lane_sel + part * 32 + 16, v + 16);
    tmem_ld_wait();
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      uint4* p = reinterpret_cast<uint4*>(sAct + tile_off(row, part * 32 + cc * 8, act_cols));
      *p = pack8_masked(v + cc * 8, *p);
    }
  };
  auto epi_dx = [&](TileCtx& c) {
    if (part == 0) {
      float v[16];
      tmem_ld16(c.tWork + lane_sel, v);
      tmem_ld_wait();
      if (c.valid) {
        float chk = 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) chk += fabsf(v[k]);
        if (!(chk < 3.0e38f)) counters[1] = counters[1] | 2;   // NaN / inf feature gradient
        float* __restrict__ o = d_feat + (c.s0 + row) * C;
        if ((C & 3) == 0) {   // static register indexing + 16-byte stores (a runtime-indexed loop spills v[])
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            if (k4 * 4 < C)
              *reinterpret_cast<float4*>(o + k4 * 4) = make_float4(v[k4 * 4] * inv_scale, v[k4 * 4 + 1] * inv_scale,
                                                                   v[k4 * 4 + 2] * inv_scale, v[k4 * 4 + 3] * inv_scale);
        } else {
#pragma unroll
          for (int k = 0; k < 16; ++k)
            if (k < C) o[k] = v[k] * inv_scale;
        }
      }
    }
  };
  // db3 += this row's (scaled, fp16) dZ3, read back from the tile the bulk copy delivered (part 0: one row per thread)
  auto add_db3 = [&](TileCtx& c, uint32_t parity) {
    if (part == 0) {
      mbar_wait(c.xbar, parity);
      const uint2 h = *reinterpret_cast<const uint2*>(c.sdZ3 + tile_off(row, 0, 16));
      const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&h.x));
      const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&h.y));
      db3[0] += a.x; db3[1] += a.y; db3[2] += b.x;
    }
  };
  // MMA batches (one elected thread)
  auto issue_l1 = [&](TileCtx& c) {
    gemm_kk(c.tWork, smem_u32(c.sX), K1, smem_u32(sW1), K1, kHid, K1, false);
    mma_commit(c.bar.bar);
  };
  auto issue_l2 = [&](TileCtx& c) {
    gemm_kk(c.tWork, smem_u32(c.sH1), kHidA, smem_u32(sW2), kHidA, kHid, kHidA, false);  // K = 144: + b2
    mma_commit(c.bar.bar);
  };
  auto issue_dw3 = [&](TileCtx& c, bool acc) {
    gemm_mm(tdW3, smem_u32(c.sH2), kHid, smem_u32(c.sdZ3), 16, 16, kTile, acc);   // dW3^T += H2^T dZ3
    gemm_kk(c.tWork, smem_u32(c.sdZ3), 16, smem_u32(sW3t), 16, kHid, 16, false);  // dH2 = dZ3 W3  (K = 16)
    mma_commit(c.bar.bar);
  };
  auto issue_l2b = [&](TileCtx& c, bool acc) {
    gemm_mm(tdW2, smem_u32(c.sH2), kHid, smem_u32(c.sH1), kHidA, kHidA, kTile, acc);   // [dW2 | db2] += dZ2^T [H1 | 1]
    gemm_km(c.tWork, smem_u32(c.sH2), kHid, smem_u32(sW2), kHidA, kHid, kHid, false);  // dH1 = dZ2 W2
    mma_commit(c.bar.bar);
  };
  auto issue_l1b = [&](TileCtx& c, bool acc) {
    gemm_mm(tdW1, smem_u32(c.sH1), kHidA, smem_u32(c.sX), K1, K1, kTile, acc);        // dW1~ += dZ1^T X~
    gemm_km(c.tWork, smem_u32(c.sH1), kHidA, smem_u32(sW1), K1, 16, kHid, false);     // dX = dZ1 W1[:, :16]
    mma_commit(c.bar.bar);
    mma_commit(c.xfree);          // second arrival of the same completion: the issuer's "tiles may be refilled"
  };
  // issuer: start the bulk copies of tile `t` (X~ and dZ3) into a context
  auto load_tiles = [&](TileCtx& c, int64_t t) {
    mbar_expect_tx(c.xbar, x_bytes + dz_bytes);
    bulk_g2s(smem_u32(c.sX), xt + t * static_cast<int64_t>(x_bytes), x_bytes, c.xbar);
    bulk_g2s(smem_u32(c.sdZ3), dzt + t * static_cast<int64_t>(dz_bytes), dz_bytes, c.xbar);
  };
  TileCtx& A = cx[0];
  TileCtx& B = cx[1];
  // epilogue side: publish this thread's smem writes / TMEM reads of a context to the issuer
  auto publish = [&](TileCtx& c) {
    fence_async_smem();
    fence_before_sync();
    mbar_arrive(c.ready.bar);
  };
  // issuer side: wait until all 512 epilogue threads have published the context
  auto acquire = [&](TileCtx& c) {
    mbar_wait(c.ready.bar, c.ready.phase);
    c.ready.phase ^= 1u;
    fence_after_sync();
  };

  int dbg_n = 0;
  auto stamp = [&]() {  // optional in-kernel timeline: CTA 0, epilogue thread 0 -> [0,64), issuer lane 0 -> [64,128)
    if (dbg && blockIdx.x == 0 && dbg_n < 64 && (tid == 0 || (is_issuer && lane == 0)))
      dbg[(is_issuer ? 64 : 0) + dbg_n++] = clock64();
  };
  if (is_issuer) {
    if (elect_one()) {
      uint32_t par = 0u;     // parity of this pair's tiles[] / xfree[] phases (one phase of each per pair)
      load_tiles(A, 2 * static_cast<int64_t>(blockIdx.x));
      load_tiles(B, 2 * static_cast<int64_t>(blockIdx.x) + 1);   // <= n_tiles: producers zero-fill up to the pair's end
      for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x, par ^= 1u) {
        const int64_t next = pair + gridDim.x;
        const bool has_next = next < n_pairs;
        if (has_next) {      // pull the next pair's tiles (contiguous in global memory) into L2 a whole pair ahead
          bulk_prefetch_l2(xt + 2 * next * static_cast<int64_t>(x_bytes), 2u * x_bytes);
          bulk_prefetch_l2(dzt + 2 * next * static_cast<int64_t>(dz_bytes), 2u * dz_bytes);
        }
        stamp();
        acquire(A); mbar_wait(A.xbar, par); stamp(); issue_l1(A); stamp();
        acquire(B); mbar_wait(B.xbar, par); stamp(); issue_l1(B); stamp();
        acquire(A); stamp(); issue_l2(A); stamp();
        acquire(B); stamp(); issue_l2(B); stamp();
        acquire(A); stamp(); issue_dw3(A, !first); stamp();
        acquire(B); stamp(); issue_dw3(B, true); stamp();
        acquire(A); stamp(); issue_l2b(A, !first); stamp();
        acquire(B); stamp(); issue_l2b(B, true); stamp();
        acquire(A); stamp(); issue_l1b(A, !first); stamp();
        acquire(B); stamp(); issue_l1b(B, true); stamp();
        first = false;
        if (has_next) {      // the pair's last batch was the only remaining reader of its X~ / dZ3 tiles
          mbar_wait(A.xfree, par); load_tiles(A, 2 * next);
          mbar_wait(B.xfree, par); load_tiles(B, 2 * next + 1);
        }
      }
    }
  } else {
    publish(A);   // nothing of the first pair to wait for: the TMEM work columns are free
    publish(B);
    uint32_t par = 0u;
    for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x, par ^= 1u) {
      A.s0 = 2 * pair * kTile;
      B.s0 = A.s0 + kTile;              // may lie past the count: then every row is invalid (an all-zero tile)
      A.valid = A.s0 + row < count;
      B.valid = B.s0 + row < count;
      stamp();
      mma_wait(A.bar); stamp(); epi_relu(A, A.sH1, kHidA); publish(A); add_db3(A, par); stamp();
      mma_wait(B.bar); stamp(); epi_relu(B, B.sH1, kHidA); publish(B); add_db3(B, par); stamp();
      mma_wait(A.bar); stamp(); epi_relu(A, A.sH2, kHid); publish(A); stamp();
      mma_wait(B.bar); stamp(); epi_relu(B, B.sH2, kHid); publish(B); stamp();
      mma_wait(A.bar); stamp(); epi_mask(A, A.sH2, kHid); publish(A); stamp();
      mma_wait(B.bar); stamp(); epi_mask(B, B.sH2, kHid); publish(B); stamp();
      mma_wait(A.bar); stamp(); epi_mask(A, A.sH1, kHidA); publish(A); stamp();
      mma_wait(B.bar); stamp(); epi_mask(B, B.sH1, kHidA); publish(B); stamp();
      mma_wait(A.bar); stamp(); epi_dx(A); publish(A);   // work columns read: layer 1 of the next pair may overwrite them
      mma_wait(B.bar); stamp(); epi_dx(B); publish(B); stamp();
    }
  }
  fence_before_sync();
  __syncthreads();

  // flush the TMEM-resident weight-gradient accumulators (lane = output feature n)
  fence_after_sync();
  if (!is_issuer) {
    const int n = row;
    float v[32];
    tmem_ld16(tdW2 + lane_sel + part * 32, v);
    tmem_ld16(tdW2 + lane_sel + part * 32 + 16, v + 16);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) atomicAdd(g.W2 + n * kHid + part * 32 + j, v[j] * inv_scale);
    if (part == 0) {
      for (int c0 = 0; c0 < K1; c0 += 16) {
        tmem_ld16(tdW1 + lane_sel + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int c = c0 + j;
          if (c < d_in) atomicAdd(g.W1 + n * d_in + c, v[j] * inv_scale);
          else if (c == d_in) atomicAdd(g.b1 + n, v[j] * inv_scale);
        }
      }
    } else if (part == 1) {
      tmem_ld16(tdW3 + lane_sel, v);  // dW3^T [j = n][c]
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 3; ++c) atomicAdd(g.W3 + c * kHid + n, v[c] * inv_scale);
    } else if (part == 2) {
      tmem_ld16(tdW2 + lane_sel + kHid, v);   // column 128 of [dW2 | db2]
      tmem_ld_wait();
      atomicAdd(g.b2 + n, v[0] * inv_scale);
    }
    if (part == 0) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float x = db3[c];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(0xffffffffu, x, off);
        if (lane == 0) atomicAdd(g.b3 + c, x * inv_scale);
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

static inline size_t mlp_fwd_smem(int K1) {
  auto a1k = [](size_t x) { return (x + 1023) & ~static_cast<size_t>(1023); };
  return 1024 + a1k(tile_bytes(kHid, K1)) + a1k(tile_bytes(kHid, kHidA)) + a1k(tile_bytes(16, kHid)) + 2 * a1k(tile_bytes(kTile, K1)) + 64;
}
static inline size_t mlp_bwd_smem(int K1) {
  auto a1k = [](size_t x) { return (x + 1023) & ~static_cast<size_t>(1023); };
  const size_t ctx = a1k(tile_bytes(kTile, K1)) + a1k(tile_bytes(kTile, kHidA)) + a1k(tile_bytes(kTile, kHid)) +
                     a1k(tile_bytes(kTile, 16));
  return 1024 + a1k(tile_bytes(kHid, K1)) + a1k(tile_bytes(kHid, kHidA)) + a1k(tile_bytes(kHid, 16)) + 2 * ctx + 64;
}

// ----------------------------------------------------------------------------------------------------
// Self test: D[128,N] = A[128,K] * B[N,K]^T for every operand orientation the MLP kernels use.
//   a_mn = 0: A given as [128][K] (K contiguous)   -> K-major tile (rows = m, cols = k)
//   a_mn = 1: A given as [K][128] (M contiguous)   -> MN-major tile (rows = k, cols = m)
//   b_mn = 0: B given as [N][K]                    -> K-major tile (rows = n, cols = k)
//   b_mn = 1: B given as [K][N]                    -> MN-major tile (rows = k, cols = n)
// ----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) tc_selftest_kernel(const float* __restrict__ A,
                                                          const float* __restrict__ B,
                                                          float* __restrict__ D, int N, int K, int a_mn,
                                                          int b_mn) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int a_rows = a_mn ? K : 128, a_cols = a_mn ? 128 : K;
  const int b_rows = b_mn ? K : N, b_cols = b_mn ? N : K;
  uint8_t* sA = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);  // 1 KiB-aligned operand tiles
  uint8_t* sB = sA + ((tile_bytes(a_rows, a_cols) + 1023) / 1024) * 1024;

  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 256);
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); mbar_init_fence(); }
  for (int i = tid; i < a_rows * a_cols; i += blockDim.x) {
    const int r = i / a_cols, c = i % a_cols;
    *reinterpret_cast<__half*>(sA + tile_off(r, c, a_cols)) = __float2half_rn(A[i]);
  }
  for (int i = tid; i < b_rows * b_cols; i += blockDim.x) {
    const int r = i / b_cols, c = i % b_cols;
    *reinterpret_cast<__half*>(sB + tile_off(r, c, b_cols)) = __float2half_rn(B[i]);
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (warp == 0 && elect_one()) {
    const uint32_t idesc = make_idesc_f16(128, N, a_mn, b_mn);
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    const uint32_t a_step = a_mn ? 2 * group_stride(a_cols) : 256u;
    const uint32_t b_step = b_mn ? 2 * group_stride(b_cols) : 256u;
    for (int k = 0; k < K / kMmaK; ++k) {
      const uint64_t da = a_mn ? desc_mnmajor(a0 + k * a_step, a_cols) : desc_kmajor(a0 + k * a_step, a_cols);
      const uint64_t db = b_mn ? desc_mnmajor(b0 + k * b_step, b_cols) : desc_kmajor(b0 + k * b_step, b_cols);
      mma_f16(tmem, da, db, idesc, k > 0 ? 1u : 0u);
    }
    mma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  fence_after_sync();
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tmem_ld16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) D[row * N + c0 + j] = v[j];
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// Descriptor probe (debug aid): A is staged K-major (validated orientation); the B region of shared
// memory is filled VERBATIM from Braw (word i -> byte 4*i) and described with caller-chosen
// LBO / SBO / per-k-step advance, so with A = unit vectors and Braw = ramp the output D reveals which
// shared-memory word the tensor core reads for each logical B[n][k].
__global__ void __launch_bounds__(128) tc_probe_kernel(const float* __restrict__ A,
                                                       const float* __restrict__ Braw, float* __restrict__ D,
                                                       int N, int K, int b_mn, uint32_t lbo, uint32_t sbo,
                                                       uint32_t kstep, int nwords) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* sA = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  uint8_t* sB = sA + ((tile_bytes(128, K) + 1023) / 1024) * 1024;
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 256);
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); mbar_init_fence(); }
  for (int i = tid; i < 128 * K; i += blockDim.x)
    *reinterpret_cast<__half*>(sA + tile_off(i / K, i % K, K)) = __float2half_rn(A[i]);
  for (int i = tid; i < nwords; i += blockDim.x) reinterpret_cast<__half*>(sB)[i] = __float2half_rn(Braw[i]);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (warp == 0 && elect_one()) {
    const uint32_t idesc = make_idesc_f16(128, N, 0, b_mn);
    for (int k = 0; k < K / kMmaK; ++k)
      mma_f16(tmem, desc_kmajor(smem_u32(sA) + k * 256u, K), make_desc(smem_u32(sB) + k * kstep, lbo, sbo),
               idesc, k > 0 ? 1u : 0u);
    mma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  fence_after_sync();
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tmem_ld16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) D[(warp * 32 + lane) * N + c0 + j] = v[j];
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// Issue-rate probe: one thread issues `reps` x `ksteps` MMAs (M = 128, N given) from zero-filled shared memory with
// caller-chosen descriptors; out[cta] = clock64 cycles from the first issue to the completion of the last one.
__global__ void __launch_bounds__(128) tc_rate_kernel(int N, int ksteps, int reps, int a_mn, int b_mn, uint32_t a_lbo,
                                                      uint32_t a_sbo, uint32_t a_kstep, uint32_t b_lbo, uint32_t b_sbo,
                                                      uint32_t b_kstep, uint32_t layout, int n_accum,
                                                      long long* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* sA = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  uint8_t* sB = sA + 64 * 1024;
  for (int i = tid; i < 32 * 1024; i += blockDim.x) reinterpret_cast<uint32_t*>(sA)[i] = 0u;   // 128 KB of zeros
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 512);
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); mbar_init_fence(); }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (warp == 0 && elect_one()) {
    const uint32_t idesc = make_idesc_f16(128, N, a_mn, b_mn);
    const uint64_t lay = static_cast<uint64_t>(layout) << 61;
    uint64_t ad[8], bd[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      ad[k] = make_desc(smem_u32(sA) + k * a_kstep, a_lbo, a_sbo) | lay;
      bd[k] = make_desc(smem_u32(sB) + k * b_kstep, b_lbo, b_sbo) | lay;
    }
    const uint32_t d1 = tmem + static_cast<uint32_t>(((n_accum > 1) ? 1 : 0) * N);
    const long long t0 = clock64();
    for (int r = 0; r < reps; r += 2) {     // descriptors precomputed: the loop body is MMAs only
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (k < ksteps) mma_f16(tmem, ad[k], bd[k], idesc, k > 0 ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (k < ksteps) mma_f16(d1, ad[k], bd[k], idesc, k > 0 ? 1u : 0u);
    }
    mma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    out[blockIdx.x] = clock64() - t0;
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// TMEM-sourced A operand probe: D[128,N] = A[128,K] * B^T with A written to tensor memory by tcgen05.st (fp16 pairs
// packed along K) and B from shared memory; then `reps` x (K/16) MMAs of the same shape are timed in the TS form
// (A from TMEM) and in the SS form (A from shared memory), cycles[0] / cycles[1].  Answers (i) is the packing
// assumption right, (ii) what does an N = 16 MMA cost when A does not use shared-memory bandwidth.
__global__ void __launch_bounds__(128) tc_ts_probe_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                          float* __restrict__ D, int N, int K, int b_mn, int reps,
                                                          long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b_rows = b_mn ? K : N, b_cols = b_mn ? N : K;
  uint8_t* sA = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  uint8_t* sB = sA + ((tile_bytes(128, K) + 1023) / 1024) * 1024;
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 512);
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); mbar_init_fence(); }
  for (int i = tid; i < 128 * K; i += blockDim.x)
    *reinterpret_cast<__half*>(sA + tile_off(i / K, i % K, K)) = __float2half_rn(A[i]);
  for (int i = tid; i < b_rows * b_cols; i += blockDim.x)
    *reinterpret_cast<__half*>(sB + tile_off(i / b_cols, i % b_cols, b_cols)) = __float2half_rn(B[i]);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t tA = tmem + 256;                       // A operand: columns [256, 256 + K/2)
  const uint32_t lane_sel = static_cast<uint32_t>(warp * 32) << 16;
  for (int k0 = 0; k0 < K; k0 += 16) {                  // this thread's row, 16 halves -> 8 columns
    uint32_t u[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const __half2 h = __floats2half2_rn(A[tid * K + k0 + 2 * j], A[tid * K + k0 + 2 * j + 1]);
      u[j] = *reinterpret_cast<const uint32_t*>(&h);
    }
    tmem_st8(tA + lane_sel + k0 / 2, u);
  }
  tmem_st_wait();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t idesc = make_idesc_f16(128, N, 0, b_mn);
  const uint32_t b_step = b_mn ? 2 * group_stride(b_cols) : 256u;
  const int ksteps = K / kMmaK;
  uint32_t phase = 0;
  if (warp == 0 && elect_one()) {
    for (int k = 0; k < ksteps; ++k) {
      const uint64_t db = b_mn ? desc_mnmajor(smem_u32(sB) + k * b_step, b_cols) : desc_kmajor(smem_u32(sB) + k * b_step, b_cols);
      mma_f16_ts(tmem, tA + k * 8, db, idesc, k > 0 ? 1u : 0u);
    }
    mma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), phase); phase ^= 1u;
  fence_after_sync();
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tmem_ld16(tmem + lane_sel + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) D[(warp * 32 + lane) * N + c0 + j] = v[j];
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  if (warp == 0 && reps > 0 && elect_one()) {
    uint64_t bd[8];
    for (int k = 0; k < 8; ++k)
      bd[k] = b_mn ? desc_mnmajor(smem_u32(sB) + (k % ksteps) * b_step, b_cols) : desc_kmajor(smem_u32(sB) + (k % ksteps) * b_step, b_cols);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r)
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (k < ksteps) mma_f16_ts(tmem + 128, tA + k * 8, bd[k], idesc, (r | k) ? 1u : 0u);
    mma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), phase); phase ^= 1u;
    cycles[0] = clock64() - t0;
    uint64_t ad[8];
    for (int k = 0; k < 8; ++k) ad[k] = desc_kmajor(smem_u32(sA) + (k % ksteps) * 256u, K);
    t0 = clock64();
    for (int r = 0; r < reps; ++r)
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (k < ksteps) mma_f16(tmem + 128, ad[k], bd[k], idesc, (r | k) ? 1u : 0u);
    mma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), phase); phase ^= 1u;
    cycles[1] = clock64() - t0;
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// TMEM -> register read-rate probe: `nwarps` warps each issue `reps` x (4 x tcgen05.ld 32x32b.x16 = 8 KB) from their own
// lane quarter; out[0] = clock64 cycles of the slowest warp, out[1] = bytes read by the CTA.  mode 1: 2 x .x32, mode 2:
// 1 x .x64 (same bytes per iteration).
__global__ void __launch_bounds__(1024) tc_ldtm_rate_kernel(int reps, int mode, long long* __restrict__ out) {
  __shared__ uint32_t tmem_base_s;
  __shared__ long long t_max;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 512);
  if (threadIdx.x == 0) t_max = 0;
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t t = tmem_base_s + (static_cast<uint32_t>((warp & 3) * 32) << 16) + static_cast<uint32_t>((warp >> 2) & 7) * 64;
  float acc = 0.f;
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    if (mode == 0) {
      float v[64];
      tmem_ld16(t, v); tmem_ld16(t + 16, v + 16); tmem_ld16(t + 32, v + 32); tmem_ld16(t + 48, v + 48);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 64; j += 16) acc += v[j];
    } else if (mode == 1) {
      uint32_t u[64];
#pragma unroll
      for (int h = 0; h < 2; ++h)
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
            : "=r"(u[32 * h + 0]), "=r"(u[32 * h + 1]), "=r"(u[32 * h + 2]), "=r"(u[32 * h + 3]), "=r"(u[32 * h + 4]),
              "=r"(u[32 * h + 5]), "=r"(u[32 * h + 6]), "=r"(u[32 * h + 7]), "=r"(u[32 * h + 8]), "=r"(u[32 * h + 9]),
              "=r"(u[32 * h + 10]), "=r"(u[32 * h + 11]), "=r"(u[32 * h + 12]), "=r"(u[32 * h + 13]), "=r"(u[32 * h + 14]),
              "=r"(u[32 * h + 15]), "=r"(u[32 * h + 16]), "=r"(u[32 * h + 17]), "=r"(u[32 * h + 18]), "=r"(u[32 * h + 19]),
              "=r"(u[32 * h + 20]), "=r"(u[32 * h + 21]), "=r"(u[32 * h + 22]), "=r"(u[32 * h + 23]), "=r"(u[32 * h + 24]),
              "=r"(u[32 * h + 25]), "=r"(u[32 * h + 26]), "=r"(u[32 * h + 27]), "=r"(u[32 * h + 28]), "=r"(u[32 * h + 29]),
              "=r"(u[32 * h + 30]), "=r"(u[32 * h + 31])
            : "r"(t + 32 * h));
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 64; j += 16) acc += __uint_as_float(u[j]);
    } else {   // 16x256b.x4: 16 lanes x 256 bits per repeat -> this warp's 32 lanes need two (lane halves)
      uint32_t u[32];
#pragma unroll
      for (int h = 0; h < 2; ++h)
        asm volatile(
            "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
            : "=r"(u[16 * h + 0]), "=r"(u[16 * h + 1]), "=r"(u[16 * h + 2]), "=r"(u[16 * h + 3]), "=r"(u[16 * h + 4]),
              "=r"(u[16 * h + 5]), "=r"(u[16 * h + 6]), "=r"(u[16 * h + 7]), "=r"(u[16 * h + 8]), "=r"(u[16 * h + 9]),
              "=r"(u[16 * h + 10]), "=r"(u[16 * h + 11]), "=r"(u[16 * h + 12]), "=r"(u[16 * h + 13]), "=r"(u[16 * h + 14]),
              "=r"(u[16 * h + 15])
            : "r"(t + (static_cast<uint32_t>(16 * h) << 16)));
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j += 8) acc += __uint_as_float(u[j]);
    }
  }
  const long long dt = clock64() - t0;
  atomicMax(reinterpret_cast<unsigned long long*>(&t_max), static_cast<unsigned long long>(dt));
  if (acc == 123.456f) out[2] = 1;   // keep the loads alive
  fence_before_sync();
  __syncthreads();
  if (threadIdx.x == 0) {
    out[0] = t_max;
    out[1] = static_cast<long long>(blockDim.x / 32) * reps * (mode == 2 ? 4096 : 8192);
  }
  if (warp == 0) tmem_dealloc(tmem_base_s, 512);
}

// What does a running tensor pipe take away from the other warps?  Warp 0 issues `n_mma` MMAs back to back (mma_mode 0:
// none, 1: SS N=128, 2: TS N=128, 3: SS N=16, 4: TS N=16) while warps 1..8 each run `reps` iterations of simt_mode
// 0: dependent FFMAs, 1: 4 x st.shared.v4, 2: 4 x ld.shared.v4, 3: 4 x ld.global.v4 (L2-resident), 4: 4 x tcgen05.ld x16.
// out[0] = cycles of the slowest SIMT warp, out[1] = cycles of the MMA batch (issue to completion).
__global__ void __launch_bounds__(288) tc_contention_kernel(int mma_mode, int n_mma, int simt_mode, int reps,
                                                            const float4* __restrict__ gbuf, long long* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ long long t_simt, t_mma;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* sA = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  uint8_t* sB = sA + 32 * 1024;
  uint8_t* sS = sB + 32 * 1024;   // 32 KB scratch for the SIMT warps
  for (int i = tid; i < 24 * 1024; i += blockDim.x) reinterpret_cast<uint32_t*>(sA)[i] = 0u;
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 512);
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); mbar_init_fence(); t_simt = 0; t_mma = 0; }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (warp == 0) {
    if (mma_mode && elect_one()) {
      const int N = mma_mode <= 2 ? 128 : 16;
      const uint32_t idesc = make_idesc_f16(128, N, 0, 0);
      const uint64_t a0 = desc_kmajor(smem_u32(sA), 128), b0 = desc_kmajor(smem_u32(sB), 128);
      const long long t0 = clock64();
      for (int i = 0; i < n_mma; i += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (mma_mode & 1) mma_f16(tmem, a0 + k * 16, b0 + k * 16, idesc, 1u);
          else mma_f16_ts(tmem, tmem + 256 + 8 * k, b0 + k * 16, idesc, 1u);
        }
      }
      mma_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), 0);
      t_mma = clock64() - t0;
    }
    __syncwarp();
  } else {
    float acc = 0.f;
    uint4* sp = reinterpret_cast<uint4*>(sS) + (warp - 1) * 128 + lane;   // 2 KB per warp, conflict-free
    const uint32_t t = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + 384 + ((warp - 1) >> 2) * 64;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      if (simt_mode == 0) {
#pragma unroll
        for (int k = 0; k < 64; ++k) acc = fmaf(acc, 1.0001f, 0.5f);
      } else if (simt_mode == 1) {
#pragma unroll
        for (int k = 0; k < 4; ++k) sp[32 * k] = make_uint4(r, k, lane, warp);
      } else if (simt_mode == 2) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { const uint4 v = sp[32 * k]; acc += __uint_as_float(v.x ^ v.w); }
      } else if (simt_mode == 3) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { const float4 v = __ldg(gbuf + ((r * 4 + k) & 255) * 256 + (warp - 1) * 32 + lane); acc += v.x + v.w; }
      } else {
        float v[64];
        tmem_ld16(t, v); tmem_ld16(t + 16, v + 16); tmem_ld16(t + 32, v + 32); tmem_ld16(t + 48, v + 48);
        tmem_ld_wait();
        acc += v[0] + v[17] + v[34] + v[51];
      }
    }
    const long long dt = clock64() - t0;
    atomicMax(reinterpret_cast<unsigned long long*>(&t_simt), static_cast<unsigned long long>(dt));
    if (acc == 123.456f) out[3] = 1;
  }
  fence_before_sync();
  __syncthreads();
  if (tid == 0) { out[0] = t_simt; out[1] = t_mma; }
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace dvgo

using namespace dvgo;

DVGO_API int dvgo_tc_contention(int mma_mode, int n_mma, int simt_mode, int reps, const void* gbuf, long long* out,
                                dvgo_stream_t stream) {
  if (mma_mode < 0 || mma_mode > 4 || simt_mode < 0 || simt_mode > 4 || n_mma < 0 || reps < 1 || !gbuf || !out) return DVGO_EINVAL;
  const size_t bytes = 97 * 1024 + 1024;
  cudaError_t e = cudaFuncSetAttribute(tc_contention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
  if (e != cudaSuccess) return static_cast<int>(e);
  tc_contention_kernel<<<1, 288, bytes, as_stream(stream)>>>(mma_mode, n_mma, simt_mode, reps,
                                                             static_cast<const float4*>(gbuf), out);
  return launch_status();
}

DVGO_API int dvgo_tc_ldtm_rate(int nwarps, int reps, int mode, long long* out, dvgo_stream_t stream) {
  if (nwarps < 1 || nwarps > 32 || reps < 1 || mode < 0 || mode > 2 || !out) return DVGO_EINVAL;
  tc_ldtm_rate_kernel<<<1, nwarps * 32, 0, as_stream(stream)>>>(reps, mode, out);
  return launch_status();
}

DVGO_API int dvgo_tc_ts_probe(const float* A, const float* B, float* D, int N, int K, int b_mn, int reps,
                              long long* cycles, dvgo_stream_t stream) {
  if (!A || !B || !D || !cycles || N < 16 || N > 128 || N % 16 || K < 16 || K > 128 || K % 16 || reps < 0) return DVGO_EINVAL;
  const size_t bytes = ((tc::tile_bytes(128, K) + 1023) / 1024) * 1024 + tc::tile_bytes(b_mn ? K : N, b_mn ? N : K) + 2048;
  cudaError_t e = cudaFuncSetAttribute(tc_ts_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
  if (e != cudaSuccess) return static_cast<int>(e);
  tc_ts_probe_kernel<<<1, 128, bytes, as_stream(stream)>>>(A, B, D, N, K, b_mn, reps, cycles);
  return launch_status();
}

DVGO_API int dvgo_tc_selftest(const float* A, const float* B, float* D, int N, int K, int a_mn, int b_mn,
                              dvgo_stream_t stream) {
  if (!A || !B || !D || N < 16 || N > 256 || N % 16 || K < 16 || K > 128 || K % 16) return DVGO_EINVAL;
  const int a_rows = a_mn ? K : 128, a_cols = a_mn ? 128 : K;
  const int b_rows = b_mn ? K : N, b_cols = b_mn ? N : K;
  const size_t bytes = ((tc::tile_bytes(a_rows, a_cols) + 1023) / 1024) * 1024 + tc::tile_bytes(b_rows, b_cols) + 2048;
  cudaError_t e = cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(bytes));
  if (e != cudaSuccess) return static_cast<int>(e);
  tc_selftest_kernel<<<1, 128, bytes, as_stream(stream)>>>(A, B, D, N, K, a_mn, b_mn);
  return launch_status();
}

DVGO_API int dvgo_tc_probe(const float* A, const float* Braw, float* D, int N, int K, int b_mn, int lbo, int sbo,
                           int kstep, int nwords, dvgo_stream_t stream) {
  if (!A || !Braw || !D || N < 16 || N > 256 || N % 16 || K < 16 || K > 128 || K % 16 || nwords > 32768)
    return DVGO_EINVAL;
  const size_t bytes = ((tc::tile_bytes(128, K) + 1023) / 1024) * 1024 + nwords * 4 + 2048;
  cudaError_t e = cudaFuncSetAttribute(tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(bytes));
  if (e != cudaSuccess) return static_cast<int>(e);
  tc_probe_kernel<<<1, 128, bytes, as_stream(stream)>>>(A, Braw, D, N, K, b_mn, lbo, sbo, kstep, nwords);
  return launch_status();
}

static inline int mlp_k1(int C, int pe_stride) { return ((C + pe_stride + 15) / 16) * 16; }

DVGO_API int64_t dvgo_mlp_wpack_bytes(int C, int pe_stride) {
  if (C < 0 || pe_stride < 1 || C + pe_stride > 64) return DVGO_EINVAL;
  return static_cast<int64_t>(wpack_layout(mlp_k1(C, pe_stride)).total);
}

DVGO_API int dvgo_mlp_pack_weights(int C, int P, int pe_stride, const float* W1, const float* b1, const float* W2,
                                   const float* b2, const float* W3, const float* b3, int width, void* wpack,
                                   dvgo_stream_t stream) {
  if (width != kHid || C < 0 || P < 0 || C + P < 1 || C + P > 63 || pe_stride < P + 1) return DVGO_EINVAL;
  if (!W1 || !b1 || !W2 || !b2 || !W3 || !b3 || !wpack) return DVGO_EINVAL;
  const int K1 = mlp_k1(C, pe_stride);
  MlpW w{W1, b1, W2, b2, W3, b3};
  mlp_pack_weights_kernel<<<32, 256, 0, as_stream(stream)>>>(w, C + P, K1, static_cast<uint8_t*>(wpack));
  return launch_status();
}

static inline bool mlp_shape_ok(int C, int P, int pe_stride) {
  return C >= 0 && P >= 0 && C + P >= 1 && C + P <= 63 && pe_stride >= P + 1 && C + pe_stride <= 64;
}
static inline int pack_grid(int64_t work) {
  const int64_t want = (work + 255) / 256;
  return static_cast<int>(want < kNumSMs * 16 ? (want > 0 ? want : 1) : kNumSMs * 16);
}

DVGO_API int64_t dvgo_mlp_xtile_bytes(int64_t surv_cap, int C, int pe_stride) {
  if (surv_cap < 0 || C < 0 || pe_stride < 1 || C + pe_stride > 64) return DVGO_EINVAL;
  return tc::tiles_for(surv_cap) * static_cast<int64_t>(tile_bytes(kTile, mlp_k1(C, pe_stride)));
}
DVGO_API int64_t dvgo_mlp_dztile_bytes(int64_t surv_cap) {
  if (surv_cap < 0) return DVGO_EINVAL;
  return tc::tiles_for(surv_cap) * static_cast<int64_t>(tile_bytes(kTile, 16));
}

DVGO_API int dvgo_mlp_pack_x(const float* feat, int C, const int32_t* s_ray, const float* pe, int P, int pe_stride,
                             const int32_t* counters, int64_t surv_cap, void* xt, dvgo_stream_t stream) {
  if (!mlp_shape_ok(C, P, pe_stride) || surv_cap < 0) return DVGO_EINVAL;
  if (surv_cap == 0) return 0;
  if ((C > 0 && !feat) || !s_ray || (P > 0 && !pe) || !counters || !xt) return DVGO_EINVAL;
  const int K1 = mlp_k1(C, pe_stride);
  mlp_pack_x_kernel<<<pack_grid(surv_cap * (K1 / 4)), 256, 0, as_stream(stream)>>>(
      feat, C, s_ray, pe, pe_stride, counters, surv_cap, K1, static_cast<uint8_t*>(xt));
  return launch_status();
}

DVGO_API int dvgo_mlp_pack_dz(const float* rgb, const float* d_rgb, float grad_scale,
                              const int32_t* counters, int64_t surv_cap, void* dzt, dvgo_stream_t stream) {
  if (surv_cap < 0 || !(grad_scale > 0.f)) return DVGO_EINVAL;
  if (surv_cap == 0) return 0;
  if (!rgb || !d_rgb || !counters || !dzt) return DVGO_EINVAL;
  mlp_pack_dz_kernel<<<pack_grid(surv_cap), 256, 0, as_stream(stream)>>>(rgb, d_rgb, grad_scale, counters, surv_cap,
                                                                        static_cast<uint8_t*>(dzt));
  return launch_status();
}

DVGO_API int dvgo_mlp_fwd_timed(const void* xt, int C, int P, int pe_stride, int32_t* counters, int64_t surv_cap,
                                const void* wpack, float* rgb, long long* timeline, dvgo_stream_t stream) {
  if (!mlp_shape_ok(C, P, pe_stride) || surv_cap < 0) return DVGO_EINVAL;
  if (!xt || !counters || !wpack || !rgb) return DVGO_EINVAL;
  if (surv_cap == 0) return 0;
  const int K1 = mlp_k1(C, pe_stride);
  const size_t bytes = mlp_fwd_smem(K1);
  cudaError_t e = cudaFuncSetAttribute(mlp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
  if (e != cudaSuccess) return static_cast<int>(e);
  const int64_t tiles = (surv_cap + kTile - 1) / kTile;
  const int grid = static_cast<int>(tiles < 2 * kNumSMs ? tiles : 2 * kNumSMs);
  mlp_fwd_kernel<<<grid, kMlpThreads, bytes, as_stream(stream)>>>(static_cast<const uint8_t*>(xt), counters, surv_cap,
                                                                 static_cast<const uint8_t*>(wpack), K1, rgb, timeline);
  return launch_status();
}

DVGO_API int dvgo_mlp_fwd(const void* xt, int C, int P, int pe_stride, int32_t* counters, int64_t surv_cap,
                          const void* wpack, float* rgb, dvgo_stream_t stream) {
  return dvgo_mlp_fwd_timed(xt, C, P, pe_stride, counters, surv_cap, wpack, rgb, nullptr, stream);
}

DVGO_API int dvgo_mlp_fwd_gather(const dvgo_scene_t* scene, const float* k0_cl, const float* s_pos,
                                 const void* pe_rows16, int P, int pe_stride, int32_t* counters, int64_t surv_cap,
                                 const void* wpack, float* rgb, void* xt_out, dvgo_stream_t stream) {
  if (!scene || !mlp_shape_ok(scene->C, P, pe_stride) || scene->C < 1 || surv_cap < 0) return DVGO_EINVAL;
  if (!k0_cl || !s_pos || !pe_rows16 || !counters || !wpack || !rgb) return DVGO_EINVAL;
  if (surv_cap == 0) return 0;
  const int K1 = mlp_k1(scene->C, pe_stride);
  const size_t bytes = mlp_fwd_smem(K1);
  const int64_t tiles = (surv_cap + kTile - 1) / kTile + 1;
  const int grid = static_cast<int>(tiles < 2 * kNumSMs ? tiles : 2 * kNumSMs);
  DVGO_DISPATCH_C(scene->C, {
    cudaError_t e = cudaFuncSetAttribute(mlp_fwd_gather_kernel<kC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(bytes));
    if (e != cudaSuccess) return static_cast<int>(e);
    mlp_fwd_gather_kernel<kC><<<grid, kFgThreads, bytes, as_stream(stream)>>>(
        to_args(scene), k0_cl, reinterpret_cast<const float4*>(s_pos), static_cast<const uint8_t*>(pe_rows16), pe_stride,
        counters, surv_cap, static_cast<const uint8_t*>(wpack), K1, rgb, static_cast<uint8_t*>(xt_out));
  });
  return launch_status();
}

DVGO_API int dvgo_mlp_bwd_timed(const void* xt, const void* dzt, int C, int P, int pe_stride, int32_t* counters,
                                int64_t surv_cap, const void* wpack, float grad_scale, float* d_feat, float* gW1,
                                float* gb1, float* gW2, float* gb2, float* gW3, float* gb3, long long* timeline,
                                dvgo_stream_t stream) {
  if (!mlp_shape_ok(C, P, pe_stride) || C < 1 || C > 16 || surv_cap < 0 || !(grad_scale > 0.f)) return DVGO_EINVAL;
  if (!xt || !dzt || !counters || !wpack || !d_feat || !gW1 || !gb1 || !gW2 || !gb2 || !gW3 || !gb3) return DVGO_EINVAL;
  if (surv_cap == 0) return 0;
  const int K1 = mlp_k1(C, pe_stride);
  const size_t bytes = mlp_bwd_smem(K1);
  cudaError_t e = cudaFuncSetAttribute(mlp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
  if (e != cudaSuccess) return static_cast<int>(e);
  const int64_t tiles = (surv_cap + kTile - 1) / kTile;
  const int64_t pairs = (tiles + 1) / 2;
  const int grid = static_cast<int>(pairs < kNumSMs ? pairs : kNumSMs);
  MlpG g{gW1, gb1, gW2, gb2, gW3, gb3};
  mlp_bwd_kernel<<<grid, kBwdThreads, bytes, as_stream(stream)>>>(
      static_cast<const uint8_t*>(xt), static_cast<const uint8_t*>(dzt), C, C + P, counters, surv_cap,
      static_cast<const uint8_t*>(wpack), K1, grad_scale, d_feat, g, timeline);
  return launch_status();
}

DVGO_API int dvgo_mlp_bwd(const void* xt, const void* dzt, int C, int P, int pe_stride, int32_t* counters,
                          int64_t surv_cap, const void* wpack, float grad_scale, float* d_feat, float* gW1, float* gb1,
                          float* gW2, float* gb2, float* gW3, float* gb3, dvgo_stream_t stream) {
  return dvgo_mlp_bwd_timed(xt, dzt, C, P, pe_stride, counters, surv_cap, wpack, grad_scale, d_feat, gW1, gb1, gW2, gb2,
                            gW3, gb3, nullptr, stream);
}

DVGO_API int dvgo_tc_rate(int ctas, int N, int ksteps, int reps, int a_mn, int b_mn, int a_lbo, int a_sbo, int a_kstep,
                          int b_lbo, int b_sbo, int b_kstep, int layout, int n_accum, long long* out, dvgo_stream_t stream) {
  if (ctas < 1 || N < 16 || N > 256 || N % 16 || ksteps < 1 || reps < 1 || n_accum < 1 || n_accum * N > 512 || !out)
    return DVGO_EINVAL;
  const size_t bytes = 129 * 1024 + 1024;
  cudaError_t e = cudaFuncSetAttribute(tc_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
  if (e != cudaSuccess) return static_cast<int>(e);
  tc_rate_kernel<<<ctas, 128, bytes, as_stream(stream)>>>(N, ksteps, reps, a_mn, b_mn, a_lbo, a_sbo, a_kstep, b_lbo, b_sbo,
                                                          b_kstep, layout, n_accum, out);
  return launch_status();
}
