// fused_mlp.cu -- the rgbnet MLP (lib/dvgo.py:123-131, applied at :524-539) on the 5th-generation
// tensor cores: tcgen05.mma kind::f16 (FP16 operands, fp32 accumulators in TMEM), operands staged in
// shared memory in the canonical no-swizzle core-matrix layout (tc_common.cuh).
//
// Numerics: fp32 master weights / features / outputs; GEMM operands rounded to FP16 (10-bit mantissa,
// the same as TF32 -- what the reference's nn.Linear uses on any tensor-core GPU under its pinned
// PyTorch 1.8.1, where allow_tf32 defaulted to True; README.md:32) on the way into shared memory;
// products accumulated in fp32.  Backward operands are scaled by a power of two so that small
// gradients stay in FP16's normal range, and unscaled exactly in the fp32 epilogue.
// Stated tolerance vs the exact-fp32 path: 2e-3 abs on rgb (tests/test_gpu_mlp.py).
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/dvgo_b200_fused.h"

namespace dvgo {

using namespace tc;

// ----------------------------------------------------------------------------------------------------
// rgbnet:  x = [k0 features (C) | view embedding (P) | 1]  ->  relu(W1 x)  ->  relu(W2 h + b2)  ->
//          sigmoid(W3 h + b3),  hidden width 128, 3 outputs  (lib/dvgo.py:123-131, :536-539).
// One tile = 128 survivors of the sample stream.  256 threads: thread (row = 32*(warp%4)+lane,
// half = warp/4) owns one sample row of the tile and half of the 128 hidden columns (a warp can only
// read its own 32-lane quarter of TMEM, so warps w and w+4 split the columns of the same rows).
//   layers 1, 2 : tcgen05.mma, A = activations [sample][feature] (K-major), B = weights [out][in]
//   layer 3     : fp32 SIMT dot products fused into the layer-2 epilogue (3 outputs -> no MMA tile)
//   b1 is folded into W1 as column d_in of the augmented input (the constant-1 feature).
// ----------------------------------------------------------------------------------------------------
constexpr int kHid = 128;
constexpr int kTile = 128;
constexpr int kMlpThreads = 256;

struct MlpW {            // fp32 master weights, torch nn.Linear layout [out][in]
  const float* W1; const float* b1; const float* W2; const float* b2; const float* W3; const float* b3;
};
struct MlpG {            // fp32 gradient accumulators, same shapes
  float* W1; float* b1; float* W2; float* b2; float* W3; float* b3;
};

__device__ __forceinline__ uint4 pack8(const float* v) {
  __half2 h[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) h[q] = __floats2half2_rn(v[2 * q], v[2 * q + 1]);
  return *reinterpret_cast<uint4*>(h);
}
__device__ __forceinline__ void unpack8(const uint4& u, float* v) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int q = 0; q < 4; ++q) { const float2 f = __half22float2(h[q]); v[2 * q] = f.x; v[2 * q + 1] = f.y; }
}

// Weights -> fp16 canonical tiles (sW1 [128 x K1] incl. the bias column, sW2 [128 x 128]) and the
// small fp32 arrays used by the SIMT parts (W3 [3][128], b2 [128], b3 [3]).
__device__ void load_weights(const MlpW& w, int d_in, int K1, uint8_t* sW1, uint8_t* sW2, float* sW3, float* sB2,
                             float* sB3) {
  for (int i = threadIdx.x; i < kHid * K1; i += blockDim.x) {
    const int n = i / K1, c = i % K1;
    const float v = c < d_in ? w.W1[n * d_in + c] : (c == d_in ? w.b1[n] : 0.f);
    *reinterpret_cast<__half*>(sW1 + tile_off(n, c, K1)) = __float2half_rn(v);
  }
  for (int i = threadIdx.x; i < kHid * kHid; i += blockDim.x)
    *reinterpret_cast<__half*>(sW2 + tile_off(i / kHid, i % kHid, kHid)) = __float2half_rn(w.W2[i]);
  for (int i = threadIdx.x; i < 3 * kHid; i += blockDim.x) sW3[i] = w.W3[i];
  for (int i = threadIdx.x; i < kHid; i += blockDim.x) sB2[i] = w.b2[i];
  if (threadIdx.x < 3) sB3[threadIdx.x] = w.b3[threadIdx.x];
}

// Augmented input tile [128 x K1]: cols [0,C) k0 features, [C,C+P) view embedding of the sample's ray,
// col C+P = 1 (bias), rest 0; rows past the survivor count are zero.
__device__ __forceinline__ void stage_x(uint8_t* sX, int K1, int64_t base, int64_t count,
                                        const float* __restrict__ feat, int C,
                                        const int32_t* __restrict__ s_ray, const float* __restrict__ pe, int P) {
  const int r = threadIdx.x & 127, h = threadIdx.x >> 7;
  const int64_t s = base + r;
  const bool valid = s < count;
  const float* __restrict__ f = feat + s * C;
  const float* __restrict__ e = pe + static_cast<int64_t>(valid ? s_ray[s] : 0) * P;
  for (int ch = h; ch < K1 / 8; ch += 2) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = ch * 8 + j;
      float x = 0.f;
      if (valid) x = c < C ? __ldg(f + c) : (c < C + P ? __ldg(e + (c - C)) : (c == C + P ? 1.f : 0.f));
      v[j] = x;
    }
    *reinterpret_cast<uint4*>(sX + tile_off(r, ch * 8, K1)) = pack8(v);
  }
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

// GEMM issue helpers (ONE thread).  A/B tiles in the canonical layout; `cols` = tile row length.
__device__ __forceinline__ void gemm_kk(uint32_t d, uint32_t a, int a_cols, uint32_t b, int b_cols, int N, int K,
                                        bool accumulate) {  // A K-major, B K-major
  const uint32_t idesc = make_idesc_f16(128, N, 0, 0);
  for (int k = 0; k < K / kMmaK; ++k)
    mma_f16(d, desc_kmajor(a + k * 256u, a_cols), desc_kmajor(b + k * 256u, b_cols), idesc, (accumulate || k) ? 1u : 0u);
}
__device__ __forceinline__ void gemm_km(uint32_t d, uint32_t a, int a_cols, uint32_t b, int b_cols, int N, int K,
                                        bool accumulate) {  // A K-major, B MN-major (B tile rows = K)
  const uint32_t idesc = make_idesc_f16(128, N, 0, 1);
  for (int k = 0; k < K / kMmaK; ++k)
    mma_f16(d, desc_kmajor(a + k * 256u, a_cols), desc_mnmajor(b + k * 2u * group_stride(b_cols), b_cols), idesc,
            (accumulate || k) ? 1u : 0u);
}
__device__ __forceinline__ void gemm_mm(uint32_t d, uint32_t a, int a_cols, uint32_t b, int b_cols, int N, int K,
                                        bool accumulate) {  // A MN-major, B MN-major (rows of both = K)
  const uint32_t idesc = make_idesc_f16(128, N, 1, 1);
  for (int k = 0; k < K / kMmaK; ++k)
    mma_f16(d, desc_mnmajor(a + k * 2u * group_stride(a_cols), a_cols),
            desc_mnmajor(b + k * 2u * group_stride(b_cols), b_cols), idesc, (accumulate || k) ? 1u : 0u);
}

// TMEM [this thread's row][c0, c0+64) -> act(v + bias) -> fp16 -> smem tile row.  bias may be null.
__device__ __forceinline__ void epi_relu_to_smem(uint32_t tmem_d, int q, int row, int c_begin, const float* sBias,
                                                 uint8_t* sOut) {
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) {
    const int c0 = c_begin + cc * 16;
    float v[16];
    tmem_ld16(tmem_d + (static_cast<uint32_t>(q * 32) << 16) + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j] + (sBias ? sBias[c0 + j] : 0.f), 0.f);
    *reinterpret_cast<uint4*>(sOut + tile_off(row, c0, kHid)) = pack8(v);
    *reinterpret_cast<uint4*>(sOut + tile_off(row, c0 + 8, kHid)) = pack8(v + 8);
  }
}

__device__ __forceinline__ size_t align1k(size_t x) { return (x + 1023) & ~static_cast<size_t>(1023); }

// ---- forward ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kMlpThreads, 2) mlp_fwd_kernel(
    const float* __restrict__ feat, int C, const int32_t* __restrict__ s_ray, const float* __restrict__ pe, int P,
    const int32_t* __restrict__ counters, int64_t surv_cap, MlpW w, int K1, float* __restrict__ rgb) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, half = warp >> 2, row = q * 32 + lane;
  const int d_in = C + P;
  int64_t count = counters[0];
  if (count > surv_cap) count = surv_cap;
  const int64_t n_tiles = (count + kTile - 1) / kTile;
  if (static_cast<int64_t>(blockIdx.x) >= n_tiles) return;  // uniform per CTA, before any allocation

  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sW1 = base;
  uint8_t* sW2 = sW1 + align1k(tile_bytes(kHid, K1));
  uint8_t* sX = sW2 + align1k(tile_bytes(kHid, kHid));
  uint8_t* sH1 = sX + align1k(tile_bytes(kTile, K1));
  float* sW3 = reinterpret_cast<float*>(sH1 + align1k(tile_bytes(kTile, kHid)));
  float* sB2 = sW3 + 3 * kHid;
  float* sB3 = sB2 + kHid;
  float* sPart = sB3 + 4;  // [128][3] partial layer-3 sums of the upper column half

  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 256);
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); mbar_init_fence(); }
  load_weights(w, d_in, K1, sW1, sW2, sW3, sB2, sB3);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t tD1 = tmem, tD2 = tmem + 128;
  MmaCtx ctx{smem_u32(&bar), 0u};

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t s0 = tile * kTile;
    stage_x(sX, K1, s0, count, feat, C, s_ray, pe, P);
    sync_for_mma();
    if (tid == 0) {
      gemm_kk(tD1, smem_u32(sX), K1, smem_u32(sW1), K1, kHid, K1, false);
      mma_commit(ctx.bar);
    }
    mma_wait(ctx);
    epi_relu_to_smem(tD1, q, row, half * 64, nullptr, sH1);
    sync_for_mma();
    if (tid == 0) {
      gemm_kk(tD2, smem_u32(sH1), kHid, smem_u32(sW2), kHid, kHid, kHid, false);
      mma_commit(ctx.bar);
    }
    mma_wait(ctx);
    // layer-2 epilogue fused with layer 3 (fp32 SIMT): acc_c = sum_j relu(z2_j + b2_j) * W3[c][j]
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      const int c0 = half * 64 + cc * 16;
      float v[16];
      tmem_ld16(tD2 + (static_cast<uint32_t>(q * 32) << 16) + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float hv = fmaxf(v[j] + sB2[c0 + j], 0.f);
        a0 = fmaf(hv, sW3[c0 + j], a0);
        a1 = fmaf(hv, sW3[kHid + c0 + j], a1);
        a2 = fmaf(hv, sW3[2 * kHid + c0 + j], a2);
      }
    }
    if (half == 1) { sPart[row * 3] = a0; sPart[row * 3 + 1] = a1; sPart[row * 3 + 2] = a2; }
    fence_before_sync();
    __syncthreads();
    if (half == 0 && s0 + row < count) {
      const float z0 = a0 + sPart[row * 3] + sB3[0];
      const float z1 = a1 + sPart[row * 3 + 1] + sB3[1];
      const float z2 = a2 + sPart[row * 3 + 2] + sB3[2];
      float* __restrict__ o = rgb + (s0 + row) * 3;
      o[0] = 1.f / (1.f + expf(-z0));
      o[1] = 1.f / (1.f + expf(-z1));
      o[2] = 1.f / (1.f + expf(-z2));
    }
    // sPart / sX / sH1 / TMEM are re-used by the next tile only after the next sync_for_mma()
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// ---- backward (with forward recompute) ---------------------------------------------------------------
// Per tile: recompute H1, H2 (as in forward); dZ3 = d_rgb * rgb (1-rgb) * S; then
//   dW3^T += H2^T dZ3      (MMA, A = H2 MN-major, B = dZ3 MN-major, N = 16)
//   dZ2    = (H2 > 0) * (dZ3 W3)                                   (SIMT epilogue, in place over H2)
//   dW2   += dZ2^T H1 ;  db2 += dZ2^T 1 ;  dH1 = dZ2 W2            (MMAs)
//   dZ1    = (H1 > 0) * dH1                                        (epilogue, in place over H1)
//   dW1~  += dZ1^T X~   (column d_in of the augmented input gives db1) ;  dX = dZ1 W1[:, :16]   (MMAs)
//   d_feat = dX[:, :C] / S
// The four weight-gradient accumulators stay in TMEM across all tiles of the CTA and are flushed to
// global memory with atomics once at the end.  S = grad_scale (a power of two) keeps the FP16
// operands of the backward GEMMs in normal range.
__global__ void __launch_bounds__(kMlpThreads, 1) mlp_bwd_kernel(
    const float* __restrict__ feat, int C, const int32_t* __restrict__ s_ray, const float* __restrict__ pe, int P,
    const int32_t* __restrict__ counters, int64_t surv_cap, MlpW w, int K1, const float* __restrict__ rgb,
    const float* __restrict__ d_rgb, float grad_scale, float* __restrict__ d_feat, MlpG g) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, half = warp >> 2, row = q * 32 + lane;
  const int d_in = C + P;
  int64_t count = counters[0];
  if (count > surv_cap) count = surv_cap;
  const int64_t n_tiles = (count + kTile - 1) / kTile;
  if (static_cast<int64_t>(blockIdx.x) >= n_tiles) return;

  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sW1 = base;
  uint8_t* sW2 = sW1 + align1k(tile_bytes(kHid, K1));
  uint8_t* sX = sW2 + align1k(tile_bytes(kHid, kHid));
  uint8_t* sH1 = sX + align1k(tile_bytes(kTile, K1));
  uint8_t* sH2 = sH1 + align1k(tile_bytes(kTile, kHid));
  uint8_t* sdZ3 = sH2 + align1k(tile_bytes(kTile, kHid));
  uint8_t* sOnes = sdZ3 + align1k(tile_bytes(kTile, 16));
  float* sW3 = reinterpret_cast<float*>(sOnes + align1k(tile_bytes(kTile, 16)));
  float* sB2 = sW3 + 3 * kHid;
  float* sB3 = sB2 + kHid;

  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 512);
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); mbar_init_fence(); }
  load_weights(w, d_in, K1, sW1, sW2, sW3, sB2, sB3);
  for (int i = tid; i < kTile * 16; i += blockDim.x)  // ones in column 0: db2 = dZ2^T * ones
    *reinterpret_cast<__half*>(sOnes + tile_off(i / 16, i % 16, 16)) = __float2half_rn((i % 16) == 0 ? 1.f : 0.f);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t tWork0 = tmem, tWork1 = tmem + 128, tdW2 = tmem + 256, tdW1 = tmem + 384, tdW3 = tmem + 432,
                 tdB2 = tmem + 448;
  MmaCtx ctx{smem_u32(&bar), 0u};
  const float inv_scale = 1.f / grad_scale;
  float db3[3] = {0.f, 0.f, 0.f};
  bool first = true;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t s0 = tile * kTile;
    const int64_t s = s0 + row;
    const bool valid = s < count;
    // dZ3 (scaled) for this thread's row, kept in registers and staged for the dW3 GEMM
    float dz[3] = {0.f, 0.f, 0.f};
    if (valid) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float o = rgb[s * 3 + c];
        dz[c] = d_rgb[s * 3 + c] * o * (1.f - o) * grad_scale;
      }
    }
    if (half == 0) {
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = j < 3 ? dz[j] : 0.f;
      *reinterpret_cast<uint4*>(sdZ3 + tile_off(row, 0, 16)) = pack8(v);
      *reinterpret_cast<uint4*>(sdZ3 + tile_off(row, 8, 16)) = pack8(v + 8);
#pragma unroll
      for (int c = 0; c < 3; ++c) db3[c] += dz[c];
    }
    stage_x(sX, K1, s0, count, feat, C, s_ray, pe, P);
    sync_for_mma();
    if (tid == 0) {
      gemm_kk(tWork0, smem_u32(sX), K1, smem_u32(sW1), K1, kHid, K1, false);
      mma_commit(ctx.bar);
    }
    mma_wait(ctx);
    epi_relu_to_smem(tWork0, q, row, half * 64, nullptr, sH1);
    sync_for_mma();
    if (tid == 0) {
      gemm_kk(tWork1, smem_u32(sH1), kHid, smem_u32(sW2), kHid, kHid, kHid, false);
      mma_commit(ctx.bar);
    }
    mma_wait(ctx);
    epi_relu_to_smem(tWork1, q, row, half * 64, sB2, sH2);
    sync_for_mma();
    if (tid == 0) {  // dW3^T [hidden j][c] += sum_s H2[s][j] dZ3[s][c]
      gemm_mm(tdW3, smem_u32(sH2), kHid, smem_u32(sdZ3), 16, 16, kTile, !first);
      mma_commit(ctx.bar);
    }
    mma_wait(ctx);
    // dZ2 in place over H2: thread (row, half) handles its 64 columns
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) {
      const int c0 = half * 64 + ch * 8;
      uint8_t* p = sH2 + tile_off(row, c0, kHid);
      float h2[8], o[8];
      unpack8(*reinterpret_cast<const uint4*>(p), h2);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float gsum = dz[0] * sW3[c0 + j] + dz[1] * sW3[kHid + c0 + j] + dz[2] * sW3[2 * kHid + c0 + j];
        o[j] = h2[j] > 0.f ? gsum : 0.f;
      }
      *reinterpret_cast<uint4*>(p) = pack8(o);
    }
    sync_for_mma();
    if (tid == 0) {
      gemm_mm(tdW2, smem_u32(sH2), kHid, smem_u32(sH1), kHid, kHid, kTile, !first);   // dW2 += dZ2^T H1
      gemm_mm(tdB2, smem_u32(sH2), kHid, smem_u32(sOnes), 16, 16, kTile, !first);     // db2 += dZ2^T 1
      gemm_km(tWork0, smem_u32(sH2), kHid, smem_u32(sW2), kHid, kHid, kHid, false);   // dH1 = dZ2 W2
      mma_commit(ctx.bar);
    }
    mma_wait(ctx);
    // dZ1 = (H1 > 0) * dH1, in place over H1
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      const int c0 = half * 64 + cc * 16;
      float v[16], h1[16];
      tmem_ld16(tWork0 + (static_cast<uint32_t>(q * 32) << 16) + c0, v);
      uint8_t* p0 = sH1 + tile_off(row, c0, kHid);
      uint8_t* p1 = sH1 + tile_off(row, c0 + 8, kHid);
      unpack8(*reinterpret_cast<const uint4*>(p0), h1);
      unpack8(*reinterpret_cast<const uint4*>(p1), h1 + 8);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = h1[j] > 0.f ? v[j] : 0.f;
      *reinterpret_cast<uint4*>(p0) = pack8(v);
      *reinterpret_cast<uint4*>(p1) = pack8(v + 8);
    }
    sync_for_mma();
    if (tid == 0) {
      gemm_mm(tdW1, smem_u32(sH1), kHid, smem_u32(sX), K1, K1, kTile, !first);        // dW1~ += dZ1^T X~
      gemm_km(tWork1, smem_u32(sH1), kHid, smem_u32(sW1), K1, 16, kHid, false);       // dX = dZ1 W1[:, :16]
      mma_commit(ctx.bar);
    }
    mma_wait(ctx);
    if (half == 0) {
      float v[16];
      tmem_ld16(tWork1 + (static_cast<uint32_t>(q * 32) << 16), v);
      tmem_ld_wait();
      if (valid) {
        float* __restrict__ o = d_feat + s * C;
        for (int c = 0; c < C; ++c) o[c] = v[c] * inv_scale;
      }
    }
    first = false;
    // buffers / TMEM work columns are re-used by the next tile after its first sync_for_mma()
    fence_before_sync();
    __syncthreads();
  }

  // flush the TMEM-resident weight-gradient accumulators (rows = lanes = output feature n)
  fence_after_sync();
  {
    const int n = row;
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      const int c0 = half * 64 + cc * 16;
      float v[16];
      tmem_ld16(tdW2 + (static_cast<uint32_t>(q * 32) << 16) + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) atomicAdd(g.W2 + n * kHid + c0 + j, v[j] * inv_scale);
    }
    if (half == 0) {
      for (int c0 = 0; c0 < K1; c0 += 16) {
        float v[16];
        tmem_ld16(tdW1 + (static_cast<uint32_t>(q * 32) << 16) + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int c = c0 + j;
          if (c < d_in) atomicAdd(g.W1 + n * d_in + c, v[j] * inv_scale);
          else if (c == d_in) atomicAdd(g.b1 + n, v[j] * inv_scale);
        }
      }
    } else {
      float v[16];
      tmem_ld16(tdW3 + (static_cast<uint32_t>(q * 32) << 16), v);  // dW3^T [j = n][c]
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 3; ++c) atomicAdd(g.W3 + c * kHid + n, v[c] * inv_scale);
      tmem_ld16(tdB2 + (static_cast<uint32_t>(q * 32) << 16), v);
      tmem_ld_wait();
      atomicAdd(g.b2 + n, v[0] * inv_scale);
    }
    if (half == 0) {
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
  return 1024 + ((tile_bytes(kHid, K1) + 1023) & ~1023u) + tile_bytes(kHid, kHid) + ((tile_bytes(kTile, K1) + 1023) & ~1023u) +
         tile_bytes(kTile, kHid) + (3 * kHid + kHid + 4 + 3 * kTile) * sizeof(float) + 64;
}
static inline size_t mlp_bwd_smem(int K1) {
  return 1024 + ((tile_bytes(kHid, K1) + 1023) & ~1023u) + tile_bytes(kHid, kHid) + ((tile_bytes(kTile, K1) + 1023) & ~1023u) +
         2 * tile_bytes(kTile, kHid) + 2 * ((tile_bytes(kTile, 16) + 1023) & ~1023u) + (3 * kHid + kHid + 4) * sizeof(float) + 64;
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
  if (tid == 0) {
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
  if (tid == 0) {
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

}  // namespace dvgo

using namespace dvgo;

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

static inline int mlp_k1(int d_in) { return ((d_in + 1 + 15) / 16) * 16; }

DVGO_API int dvgo_mlp_fwd(const float* feat, int C, const int32_t* s_ray, const float* pe, int P,
                          const int32_t* counters, int64_t surv_cap, const float* W1, const float* b1,
                          const float* W2, const float* b2, const float* W3, const float* b3, int width, float* rgb,
                          dvgo_stream_t stream) {
  if (width != kHid || C < 0 || P < 0 || C + P < 1 || C + P > 63 || surv_cap < 0) return DVGO_EINVAL;
  if (!feat || !s_ray || (P > 0 && !pe) || !counters || !W1 || !b1 || !W2 || !b2 || !W3 || !b3 || !rgb)
    return DVGO_EINVAL;
  if (surv_cap == 0) return 0;
  const int K1 = mlp_k1(C + P);
  const size_t bytes = mlp_fwd_smem(K1);
  cudaError_t e = cudaFuncSetAttribute(mlp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
  if (e != cudaSuccess) return static_cast<int>(e);
  const int64_t tiles = (surv_cap + kTile - 1) / kTile;
  const int grid = static_cast<int>(tiles < 2 * kNumSMs ? tiles : 2 * kNumSMs);
  MlpW w{W1, b1, W2, b2, W3, b3};
  mlp_fwd_kernel<<<grid, kMlpThreads, bytes, as_stream(stream)>>>(feat, C, s_ray, pe, P, counters, surv_cap, w, K1, rgb);
  return launch_status();
}

DVGO_API int dvgo_mlp_bwd(const float* feat, int C, const int32_t* s_ray, const float* pe, int P,
                          const int32_t* counters, int64_t surv_cap, const float* W1, const float* b1,
                          const float* W2, const float* b2, const float* W3, const float* b3, int width,
                          const float* rgb, const float* d_rgb, float grad_scale, float* d_feat, float* gW1,
                          float* gb1, float* gW2, float* gb2, float* gW3, float* gb3, dvgo_stream_t stream) {
  if (width != kHid || C < 1 || C > 16 || P < 0 || C + P > 63 || surv_cap < 0 || !(grad_scale > 0.f)) return DVGO_EINVAL;
  if (!feat || !s_ray || (P > 0 && !pe) || !counters || !W1 || !b1 || !W2 || !b2 || !W3 || !b3 || !rgb || !d_rgb ||
      !d_feat || !gW1 || !gb1 || !gW2 || !gb2 || !gW3 || !gb3)
    return DVGO_EINVAL;
  if (surv_cap == 0) return 0;
  const int K1 = mlp_k1(C + P);
  const size_t bytes = mlp_bwd_smem(K1);
  cudaError_t e = cudaFuncSetAttribute(mlp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
  if (e != cudaSuccess) return static_cast<int>(e);
  const int64_t tiles = (surv_cap + kTile - 1) / kTile;
  const int grid = static_cast<int>(tiles < kNumSMs ? tiles : kNumSMs);
  MlpW w{W1, b1, W2, b2, W3, b3};
  MlpG g{gW1, gb1, gW2, gb2, gW3, gb3};
  mlp_bwd_kernel<<<grid, kMlpThreads, bytes, as_stream(stream)>>>(feat, C, s_ray, pe, P, counters, surv_cap, w, K1, rgb,
                                                                 d_rgb, grad_scale, d_feat, g);
  return launch_status();
}
