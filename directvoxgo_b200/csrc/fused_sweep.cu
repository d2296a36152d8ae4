// fused_sweep.cu -- ONE pass over each grid per step: total-variation gradient + (masked) Adam +
// gradient re-zeroing, on the trainer's layouts (density [X,Y,Z], k0 channel-last [X,Y,Z,C]).
//
// Reference sequence (run.py:389-397): total_variation_add_grad (read p + 6 neighbours, RMW grad:
// 12 B/elem of HBM) -> masked_adam_upd (read g,p,m,v, write p,m,v: 28 B/elem) -> next step's
// zero_grad/backward re-materialise grad (4 B/elem) = 44 B/elem in 2 launches + allocator traffic.
// Here: read p,g,m,v (16 B) + write p',m,v,g=0 (16 B) = 32 B/elem in one launch; the six TV
// neighbours are L1/L2 hits (z +-1: same/adjacent line; y +-1: Z*C*4 B away; x +-1: Y*Z*C*4 B away,
// re-used within ~2.5 MB of streaming, far below the 126 MB L2).
//
// Stencil-then-write hazard: TV must read the OLD neighbours (total_variation_kernel.cu:27-32), so
// with TV on, new parameters go to a second buffer (ping-pong, no extra traffic); without TV the
// update is in place.  Semantics kept from the reference: weights /6, wx unused and the x axis
// uses wz (:31-32, :45-47), term order k-,k+,j-,j+,i-,i+, sparse TV gates on grad != 0 BEFORE the
// add (:21), masked Adam gates on grad != 0 AFTER it (adam_upd_kernel.cu:35).
#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/dvgo_b200_fused.h"

namespace dvgo {

__device__ __forceinline__ float clamp1f(float v) { return fminf(fmaxf(v, -1.f), 1.f); }
__device__ __forceinline__ float tvt(float w, float p, float pn) { return fmul(w, clamp1f(fsub(p, pn))); }

__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, float lrs,
                                          float step_size, float beta1, float beta2, float eps) {
  m = fma_(beta1, m, fmul(fsub(1.f, beta1), g));
  v = fma_(beta2, v, fmul(fmul(fsub(1.f, beta2), g), g));
  p = fsub(p, fdiv(fmul(fmul(step_size, lrs), m), fadd(sqrtf(v), eps)));
}
__device__ __forceinline__ void adam_elem_nolr(float& p, float g, float& m, float& v, float step_size,
                                               float beta1, float beta2, float eps) {
  m = fma_(beta1, m, fmul(fsub(1.f, beta1), g));
  v = fma_(beta2, v, fmul(fmul(fsub(1.f, beta2), g), g));
  p = fsub(p, fdiv(fmul(step_size, m), fadd(sqrtf(v), eps)));
}

// Ray-sharded data parallel over NVLink peer memory (dvgo_fused_sweep_peer): every rank owns an x-slab; for the
// elements of its slab it reads the gradient of ALL ranks straight from their buffers (peer loads), sums them in
// rank order, does TV + Adam, and stores the new parameters into ALL ranks' parameter buffers (peer stores).  The
// reduce-scatter, the sweep and the all-gather are ONE kernel: no staging pass over local HBM, and the NVLink
// transfers overlap the HBM streaming of p, m, v element by element.
struct SweepPeers {
  int n;                  // 0: local mode (single GPU, or gradients already reduced by NCCL)
  const float* grad[8];   // grad[r]: rank r's gradient accumulator (same layout everywhere)
  float* pout[8];         // pout[r]: rank r's output parameter buffer
  // NVLink SHARP (NVLS) multicast addresses of the same two buffers, or null: one multimem.ld_reduce returns the
  // gradient already summed over all ranks by the switch, one multimem.st writes all ranks' parameter buffers --
  // the per-GPU NVLink traffic drops from (n-1)/n of the grid each way to 1/n.
  const float* grad_mc;
  float* pout_mc;
  int prefetch;           // multicast path, opt-in (DVGO_PEER_PREFETCH=1): the summed gradient of the NEXT grid-stride
                          // element is requested one iteration ahead (two multimem.ld_reduce in flight per thread).
                          // Measured at 4 ranks on B200: no gain (sweep stage 0.517 vs 0.510 ms) -- the switch's
                          // reduction rate, not the number of requests in flight, bounds the peer sweep.
};

__device__ __forceinline__ float4 multimem_ld_reduce_add4(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ float multimem_ld_reduce_add1(const float* mc) {
  float v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f32 %0, [%1];" : "=f"(v) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st4(float* mc, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void multimem_st1(float* mc, float v) {
  asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(mc), "f"(v) : "memory");
}

// VEC = 4, kZ = false: C % 4 == 0, each thread owns one float4 = 4 channels of one voxel.
// VEC = 4, kZ = true : C == 1 and Z % 4 == 0 (density): one float4 = 4 consecutive z voxels; the z neighbours are the
//                      vector shifted by one lane plus one scalar load at each end.
// VEC = 1: generic (C == 3 or 9 grids, odd Z).
// kEager: every element will be updated (dense TV, or Adam without the zero-gradient mask), so p, g, m, v and the six
//         neighbour vectors are all loaded up front, independent of each other: ~10 loads in flight per thread instead
//         of three dependent phases (p,g -> neighbours -> m,v).  The lazy variant keeps the phases: on sparse scenes
//         (masked Adam, sparse TV) most elements stop after reading g (4 B/elem).
#ifndef DVGO_PEER_MINB
#define DVGO_PEER_MINB 4
#endif
template <int VEC, bool kTV, bool kEager, bool kZ, bool kPeer>
__global__ void __launch_bounds__(256, kPeer ? DVGO_PEER_MINB : 1) sweep_kernel(
    const float* __restrict__ pin, float* __restrict__ pout, float* __restrict__ grad,
    float* __restrict__ m_, float* __restrict__ v_, const float* __restrict__ perlr, int X, int Y,
    int Z, int C, int x_begin, int x_end, int tv_dense, float wy, float wz, int masked,
    float step_size, float beta1, float beta2, float eps, SweepPeers peers) {
  static_assert(!kZ || VEC == 4, "z-vectorised variant uses float4");
  // this launch owns the x-slab [x_begin, x_end) (the whole grid on one GPU, 1/n of it when the sweep is
  // sharded after a reduce-scatter); neighbours outside the slab are still read from the full buffers
  const int64_t e_begin = static_cast<int64_t>(x_begin) * Y * Z * C;
  const int64_t n_work = static_cast<int64_t>(x_end - x_begin) * Y * Z * C / VEC;
  const int64_t sz = C;                                // element stride of z +- 1
  const int64_t sy = static_cast<int64_t>(Z) * C;      // y +- 1
  const int64_t sx = static_cast<int64_t>(Y) * Z * C;  // x +- 1
  auto ld = [](const float* q, float* o) {
    if constexpr (VEC == 4) {
      const float4 a = *reinterpret_cast<const float4*>(q);
      o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
    } else {
      o[0] = *q;
    }
  };
  auto ldg = [](const float* q, float* o) {
    if constexpr (VEC == 4) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(q));
      o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
    } else {
      o[0] = __ldg(q);
    }
  };
  auto st = [](float* q, const float* o) {
    if constexpr (VEC == 4) *reinterpret_cast<float4*>(q) = make_float4(o[0], o[1], o[2], o[3]);
    else *q = o[0];
  };
  const int64_t q_first = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t q_stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const bool mc_prefetch = kPeer && peers.grad_mc && peers.prefetch;
  float gpre[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) gpre[k] = 0.f;
  auto ld_mc = [&](int64_t e, float* o) {
    if constexpr (VEC == 4) {
      const float4 a = multimem_ld_reduce_add4(peers.grad_mc + e);
      o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
    } else {
      o[0] = multimem_ld_reduce_add1(peers.grad_mc + e);
    }
  };
  if (mc_prefetch && q_first < n_work) ld_mc(e_begin + q_first * VEC, gpre);
  for (int64_t q = q_first; q < n_work; q += q_stride) {
    const int64_t e0 = e_begin + q * VEC;
    float p[VEC], g[VEC], m[VEC], v[VEC], l[VEC];
    ld(pin + e0, p);
    ld(grad + e0, g);
    if constexpr (kEager) {
      ld(m_ + e0, m);
      ld(v_ + e0, v);
      if (perlr) ldg(perlr + e0, l);
    }
    bool dirty = false;  // original (local) gradient non-zero somewhere -> must be re-zeroed
#pragma unroll
    for (int k = 0; k < VEC; ++k) dirty = dirty || (g[k] != 0.f);
    if (kPeer && peers.grad_mc) {   // summed by the switch
      if (mc_prefetch) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) g[k] = gpre[k];
        if (q + q_stride < n_work) ld_mc(e_begin + (q + q_stride) * VEC, gpre);
      } else {
        ld_mc(e0, g);
      }
    } else if constexpr (kPeer) {   // g = sum over ranks, in rank order (peers.grad[self] is the local buffer)
#pragma unroll
      for (int k = 0; k < VEC; ++k) g[k] = 0.f;
#pragma unroll
      for (int b = 0; b < 8; b += 4) {       // four peers' vectors in flight at a time
        if (b < peers.n) {
          float gr[4][VEC];
#pragma unroll
          for (int r = 0; r < 4; ++r)
            if (b + r < peers.n) ld(peers.grad[b + r] + e0, gr[r]);
#pragma unroll
          for (int r = 0; r < 4; ++r)
            if (b + r < peers.n) {
#pragma unroll
              for (int k = 0; k < VEC; ++k) g[k] = fadd(g[k], gr[r][k]);
            }
        }
      }
    }
    if (kTV) {
      bool any = kEager || tv_dense != 0;
#pragma unroll
      for (int k = 0; k < VEC; ++k) any = any || (g[k] != 0.f);
      if (any) {
        const int64_t vox = e0 / C;
        const int z = static_cast<int>(vox % Z);
        const int y = static_cast<int>((vox / Z) % Y);
        const int x = static_cast<int>(vox / (static_cast<int64_t>(Z) * Y));
        float add[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) add[k] = 0.f;
        // all neighbour loads first (independent), then the adds in the reference's term order k-,k+,j-,j+,i-,i+
        float nzm[VEC], nzp[VEC], nym[VEC], nyp[VEC], nxm[VEC], nxp[VEC];
        const bool oym = y > 0, oyp = y < Y - 1, oxm = x > 0, oxp = x < X - 1;
        bool ozm[VEC], ozp[VEC];
        if constexpr (kZ) {
          const float before = z > 0 ? __ldg(pin + e0 - 1) : 0.f;
          const float after = z + VEC < Z ? __ldg(pin + e0 + VEC) : 0.f;
#pragma unroll
          for (int k = 0; k < VEC; ++k) {
            nzm[k] = k == 0 ? before : p[k - 1];
            nzp[k] = k == VEC - 1 ? after : p[k + 1];
            ozm[k] = z + k > 0;
            ozp[k] = z + k < Z - 1;
          }
        } else {
          const bool a = z > 0, b2 = z < Z - 1;
          if (a) ldg(pin + e0 - sz, nzm);
          if (b2) ldg(pin + e0 + sz, nzp);
#pragma unroll
          for (int k = 0; k < VEC; ++k) { ozm[k] = a; ozp[k] = b2; }
        }
        if (oym) ldg(pin + e0 - sy, nym);
        if (oyp) ldg(pin + e0 + sy, nyp);
        if (oxm) ldg(pin + e0 - sx, nxm);
        if (oxp) ldg(pin + e0 + sx, nxp);
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          if (ozm[k]) add[k] = fadd(add[k], tvt(wz, p[k], nzm[k]));
          if (ozp[k]) add[k] = fadd(add[k], tvt(wz, p[k], nzp[k]));
          if (oym) add[k] = fadd(add[k], tvt(wy, p[k], nym[k]));
          if (oyp) add[k] = fadd(add[k], tvt(wy, p[k], nyp[k]));
          if (oxm) add[k] = fadd(add[k], tvt(wz, p[k], nxm[k]));      // reference quirk: the x axis uses wz
          if (oxp) add[k] = fadd(add[k], tvt(wz, p[k], nxp[k]));
        }
#pragma unroll
        for (int k = 0; k < VEC; ++k)
          if (tv_dense || g[k] != 0.f) g[k] = fadd(g[k], add[k]);
      }
    }
    bool upd = !masked;
#pragma unroll
    for (int k = 0; k < VEC; ++k) upd = upd || (g[k] != 0.f);
    if (upd) {
      if constexpr (!kEager) {
        ld(m_ + e0, m);
        ld(v_ + e0, v);
        if (perlr) ldg(perlr + e0, l);
      }
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        if (masked && g[k] == 0.f) continue;  // adam_upd_kernel.cu:35
        if (perlr) adam_elem(p[k], g[k], m[k], v[k], l[k], step_size, beta1, beta2, eps);
        else adam_elem_nolr(p[k], g[k], m[k], v[k], step_size, beta1, beta2, eps);
      }
      st(m_ + e0, m);
      st(v_ + e0, v);
    }
    // new parameters: always written when ping-ponging (pout != pin), only when changed in place
    if (upd || pout != pin) {
      if (kPeer && peers.pout_mc) {
        if constexpr (VEC == 4) multimem_st4(peers.pout_mc + e0, make_float4(p[0], p[1], p[2], p[3]));
        else multimem_st1(peers.pout_mc + e0, p[0]);
      } else if constexpr (kPeer) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
          if (r < peers.n) st(peers.pout[r] + e0, p);
      } else {
        st(pout + e0, p);
      }
    }
    // re-zero the gradient accumulator for the next step (only where it was non-zero)
    if (dirty) {
      float zero[VEC];
#pragma unroll
      for (int k = 0; k < VEC; ++k) zero[k] = 0.f;
      st(grad + e0, zero);
    }
  }
  if constexpr (kPeer) __threadfence_system();   // peer / multicast stores performed before the kernel retires
}

// ---- row-staged sweep: bulk copies (cp.async.bulk, the TMA engine's 1-D form) instead of per-thread loads -------------
// For a channel-last grid a z-row (x, y, 0..Z-1) is Z*C contiguous floats (7.7 KB at 160 x 12) and the six TV
// neighbours of its elements lie in that row (z -+ 1) and in four other whole rows ((x, y -+ 1), (x -+ 1, y)).  One
// elected thread per CTA pulls, per row, up to eight rows into a shared-memory stage -- p of the row and of its four
// neighbours, g, m, v -- with one bulk copy each, arms an mbarrier with the byte count, and runs kRowStages rows ahead;
// eight consumer warps compute TV + Adam out of shared memory and store p', m, v, g = 0 straight to global memory.
// What this buys over sweep_kernel<4, TV, EAGER>: the loads need no registers and no LSU requests (that kernel holds
// ten 16-byte loads per thread in 114 registers at 25 % occupancy and runs at 0.71 of the HBM rate against 0.96
// without TV), and 3 x 61 KB are in flight per SM.  Same arithmetic, same term order (bit-identical results).
// Used for every TV sweep (dense or sparse) of a vectorisable grid on one GPU (or an NCCL slab): measured on B200, k0
// 160^3 x 12: dense TV 0.332 -> 0.278 ms (5.67 TB/s), sparse TV (10 % of the cells touched) 0.48 -> 0.28 ms.  The peer /
// per-lr variants and the sweeps without TV (in place, lazy: most elements stop after reading g) stay with sweep_kernel.
constexpr int kRowStages = 3;
constexpr int kRowConsumers = 512;   // + one producer warp (measured: 256 -> 0.365 ms, 512 -> 0.308 ms, 768 -> 0.308 ms)

// A row longer than the stage budget is processed as `nseg` equal SEGMENTS (voxel-aligned); the centre copy of a
// segment carries one voxel of z-halo on each side (clipped at the row ends), the seven other copies are the segment.
__global__ void __launch_bounds__(kRowConsumers + 32, 1) sweep_rows_kernel(
    const float* __restrict__ pin, float* __restrict__ pout, float* __restrict__ grad, float* __restrict__ m_,
    float* __restrict__ v_, int X, int Y, int Z, int C, int x_begin, int x_end, int nseg, int tv_dense, float wy,
    float wz, int masked, float step_size, float beta1, float beta2, float eps) {
  extern __shared__ __align__(128) uint8_t srow[];
  __shared__ __align__(8) uint64_t bars[2 * kRowStages];   // full[s], empty[s]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row_f4 = Z * C / 4;                        // float4 per row
  const int seg_f4 = row_f4 / nseg;                    // float4 per segment (a multiple of G)
  const int G = C / 4;                                 // float4 per voxel
  const uint32_t seg_bytes = static_cast<uint32_t>(seg_f4) * 16u;
  const uint32_t ctr_bytes = (static_cast<uint32_t>(seg_f4 + 2 * G) * 16u + 127u) & ~127u;   // centre buffer: + halo
  const int64_t n_units = static_cast<int64_t>(x_end - x_begin) * Y * nseg;
  // units are dealt round-robin: at any time the CTAs work on ~gridDim consecutive segments, a compact window of each
  // array (one contiguous run per CTA -- 148 x 8 streams 1.3 MB apart -- measured slower: DRAM page locality)
  const int64_t u0 = blockIdx.x, u1 = n_units, u_step = gridDim.x;
  if (u0 >= u1) return;
  const int64_t sy = static_cast<int64_t>(Z) * C, sx = sy * Y;
  using namespace tc;
  if (tid == 0) {
    for (int s = 0; s < kRowStages; ++s) {
      mbar_init(smem_u32(&bars[s]), 1);
      mbar_init(smem_u32(&bars[kRowStages + s]), kRowConsumers / 32);
    }
    mbar_init_fence();
  }
  __syncthreads();
  // stage layout: [p with halo | p(y-1) | p(y+1) | p(x-1) | p(x+1) | g | m | v]
  const size_t stage_bytes = ctr_bytes + 7u * static_cast<size_t>(seg_bytes);
  auto stage = [&](int s, int which) {
    return srow + static_cast<size_t>(s) * stage_bytes + (which == 0 ? 0u : ctr_bytes + (which - 1) * static_cast<size_t>(seg_bytes));
  };

  if (warp == kRowConsumers / 32) {
    // ---- producer ----
    if (elect_one()) {
      uint32_t ephase = 0u;
      int64_t j = 0;
      for (int64_t u = u0; u < u1; u += u_step, ++j) {
        const int s = static_cast<int>(j % kRowStages);
        if (j >= kRowStages) {             // the consumers have finished with the unit that used this stage
          mbar_wait(smem_u32(&bars[kRowStages + s]), (ephase >> s) & 1u);
          ephase ^= 1u << s;
        }
        const int64_t r = u / nseg;
        const int seg = static_cast<int>(u - r * nseg);
        const int x = x_begin + static_cast<int>(r / Y), y = static_cast<int>(r % Y);
        const int f0 = seg * seg_f4;                                    // first float4 of the segment in its row
        const int64_t e0 = (static_cast<int64_t>(x) * Y + y) * sy + static_cast<int64_t>(f0) * 4;
        const int lo = f0 >= G ? G : 0, hi = f0 + seg_f4 + G <= row_f4 ? G : 0;   // halo float4 present on each side
        const uint32_t full = smem_u32(&bars[s]);
        const int n_other = 3 + (y > 0) + (y < Y - 1) + (x > 0) + (x < X - 1);
        mbar_expect_tx(full, static_cast<uint32_t>(seg_f4 + lo + hi) * 16u + n_other * seg_bytes);
        // element f0 + i of the row sits at float4 index G + i of the centre buffer
        bulk_g2s(smem_u32(stage(s, 0)) + static_cast<uint32_t>(G - lo) * 16u, pin + e0 - lo * 4,
                 static_cast<uint32_t>(seg_f4 + lo + hi) * 16u, full);
        if (y > 0) bulk_g2s(smem_u32(stage(s, 1)), pin + e0 - sy, seg_bytes, full);
        if (y < Y - 1) bulk_g2s(smem_u32(stage(s, 2)), pin + e0 + sy, seg_bytes, full);
        if (x > 0) bulk_g2s(smem_u32(stage(s, 3)), pin + e0 - sx, seg_bytes, full);
        if (x < X - 1) bulk_g2s(smem_u32(stage(s, 4)), pin + e0 + sx, seg_bytes, full);
        bulk_g2s(smem_u32(stage(s, 5)), grad + e0, seg_bytes, full);
        bulk_g2s(smem_u32(stage(s, 6)), m_ + e0, seg_bytes, full);
        bulk_g2s(smem_u32(stage(s, 7)), v_ + e0, seg_bytes, full);
      }
    }
    return;
  }

  // ---- consumers ----
  uint32_t fphase = 0u;
  int64_t j = 0;
  for (int64_t u = u0; u < u1; u += u_step, ++j) {
    const int s = static_cast<int>(j % kRowStages);
    mbar_wait(smem_u32(&bars[s]), (fphase >> s) & 1u);
    fphase ^= 1u << s;
    const int64_t r = u / nseg;
    const int seg = static_cast<int>(u - r * nseg);
    const int x = x_begin + static_cast<int>(r / Y), y = static_cast<int>(r % Y);
    const int f0 = seg * seg_f4;
    const int64_t e0 = (static_cast<int64_t>(x) * Y + y) * sy + static_cast<int64_t>(f0) * 4;
    const bool oym = y > 0, oyp = y < Y - 1, oxm = x > 0, oxp = x < X - 1;
    const float4* __restrict__ sp = reinterpret_cast<const float4*>(stage(s, 0)) + G;   // sp[i]: element f0 + i
    const float4* __restrict__ sym = reinterpret_cast<const float4*>(stage(s, 1));
    const float4* __restrict__ syp = reinterpret_cast<const float4*>(stage(s, 2));
    const float4* __restrict__ sxm = reinterpret_cast<const float4*>(stage(s, 3));
    const float4* __restrict__ sxp = reinterpret_cast<const float4*>(stage(s, 4));
    const float4* __restrict__ sg = reinterpret_cast<const float4*>(stage(s, 5));
    const float4* __restrict__ sm = reinterpret_cast<const float4*>(stage(s, 6));
    const float4* __restrict__ sv = reinterpret_cast<const float4*>(stage(s, 7));
    for (int i = tid; i < seg_f4; i += kRowConsumers) {
      const bool ozm = f0 + i >= G, ozp = f0 + i < row_f4 - G;   // z > 0, z < Z - 1 (no integer division by the run-time G)
      const float4 p4 = sp[i], g4 = sg[i], m4 = sm[i], v4 = sv[i];
      const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 nzm = ozm ? sp[i - G] : zero4, nzp = ozp ? sp[i + G] : zero4;
      const float4 nym = oym ? sym[i] : zero4, nyp = oyp ? syp[i] : zero4;
      const float4 nxm = oxm ? sxm[i] : zero4, nxp = oxp ? sxp[i] : zero4;
      float p[4] = {p4.x, p4.y, p4.z, p4.w}, g[4] = {g4.x, g4.y, g4.z, g4.w};
      float m[4] = {m4.x, m4.y, m4.z, m4.w}, v[4] = {v4.x, v4.y, v4.z, v4.w};
      const float azm[4] = {nzm.x, nzm.y, nzm.z, nzm.w}, azp[4] = {nzp.x, nzp.y, nzp.z, nzp.w};
      const float aym[4] = {nym.x, nym.y, nym.z, nym.w}, ayp[4] = {nyp.x, nyp.y, nyp.z, nyp.w};
      const float axm[4] = {nxm.x, nxm.y, nxm.z, nxm.w}, axp[4] = {nxp.x, nxp.y, nxp.z, nxp.w};
      bool dirty = false;
#pragma unroll
      for (int k = 0; k < 4; ++k) dirty = dirty || (g[k] != 0.f);
#pragma unroll
      for (int k = 0; k < 4; ++k) {      // term order of the reference: k-, k+, j-, j+, i-, i+ (x axis uses wz: its quirk)
        float add = 0.f;
        if (ozm) add = fadd(add, tvt(wz, p[k], azm[k]));
        if (ozp) add = fadd(add, tvt(wz, p[k], azp[k]));
        if (oym) add = fadd(add, tvt(wy, p[k], aym[k]));
        if (oyp) add = fadd(add, tvt(wy, p[k], ayp[k]));
        if (oxm) add = fadd(add, tvt(wz, p[k], axm[k]));
        if (oxp) add = fadd(add, tvt(wz, p[k], axp[k]));
        if (tv_dense || g[k] != 0.f) g[k] = fadd(g[k], add);
      }
      bool upd = !masked;
#pragma unroll
      for (int k = 0; k < 4; ++k) upd = upd || (g[k] != 0.f);
      const int64_t e = e0 + static_cast<int64_t>(i) * 4;
      if (upd) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (masked && g[k] == 0.f) continue;  // adam_upd_kernel.cu:35
          adam_elem_nolr(p[k], g[k], m[k], v[k], step_size, beta1, beta2, eps);
        }
        *reinterpret_cast<float4*>(m_ + e) = make_float4(m[0], m[1], m[2], m[3]);
        *reinterpret_cast<float4*>(v_ + e) = make_float4(v[0], v[1], v[2], v[3]);
      }
      *reinterpret_cast<float4*>(pout + e) = make_float4(p[0], p[1], p[2], p[3]);   // ping-pong: always written
      if (dirty) *reinterpret_cast<float4*>(grad + e) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&bars[kRowStages + s]));
  }
}

// ---- cross-GPU ordering and the small (rgbnet) gradient over peer memory --------------------------------------------
// Round 1 ordered the ranks around the peer sweep with two tiny NCCL all-reduces (~55 + ~70 us at 8 ranks, 6 % of the
// step).  Here each rank owns an array of n int32 flags in symmetric memory: a barrier is one 32-thread kernel in which
// lane r stores the new epoch into slot `self` of rank r's array (st.release.sys over NVLink) and then spins on slot r of
// its own array (ld.acquire.sys, local memory) until it shows the epoch.  Stream order makes every earlier kernel of
// this GPU complete (its writes performed) before the release stores; the acquire loads order the later kernels' peer
// reads after the other ranks' writes.  The spin is bounded: a missing rank traps instead of hanging the GPU.
struct PeerFlags {
  int32_t* flags[8];
  int n, self;
};
__global__ void __launch_bounds__(32) peer_barrier_kernel(PeerFlags pf, int epoch) {
  const int r = threadIdx.x;
  if (r >= pf.n) return;
  __threadfence_system();
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(pf.flags[r] + pf.self), "r"(epoch) : "memory");
  const int32_t* mine = pf.flags[pf.self] + r;
  const long long t0 = clock64();
  int v;
  do {
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
    if (v - epoch >= 0) break;
    __nanosleep(100);
    if (clock64() - t0 > 8000000000LL) __trap();   // ~4 s at 1.9 GHz: a rank is missing
  } while (true);
  __threadfence_system();
}

// Adam on a small replicated tensor whose gradient is the SUM over ranks, read straight from the peers' buffers (or,
// with NVLS, already summed by the switch): every rank computes the identical update, no collective call.
struct SmallPeers {
  const float* grad[8];
  const float* grad_mc;
  int n;
};
__global__ void __launch_bounds__(256) adam_peer_kernel(float* __restrict__ param, SmallPeers sp, float* __restrict__ m_,
                                                        float* __restrict__ v_, int64_t N, float step_size, float beta1,
                                                        float beta2, float eps) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < N;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float g = 0.f;
    if (sp.grad_mc) {
      g = multimem_ld_reduce_add1(sp.grad_mc + i);
    } else {
#pragma unroll
      for (int r = 0; r < 8; ++r)
        if (r < sp.n) g = fadd(g, sp.grad[r][i]);      // rank order: every rank gets bit-identical sums
    }
    float p = param[i], m = m_[i], v = v_[i];
    adam_elem_nolr(p, g, m, v, step_size, beta1, beta2, eps);
    param[i] = p; m_[i] = m; v_[i] = v;
  }
}

__global__ void __launch_bounds__(256) ncdhw_to_cl_kernel(const float* __restrict__ src,
                                                          float* __restrict__ dst, int C, int64_t G) {
  const int64_t n = G * C;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t vox = i / C;
    const int c = static_cast<int>(i - vox * C);
    dst[i] = src[c * G + vox];
  }
}

__global__ void __launch_bounds__(256) cl_to_ncdhw_kernel(const float* __restrict__ src,
                                                          float* __restrict__ dst, int C, int64_t G) {
  const int64_t n = G * C;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i / G);
    const int64_t vox = i - c * G;
    dst[i] = src[vox * C + c];
  }
}

static inline int sweep_grid(int64_t n, int threads) {
  const int64_t want = (n + threads - 1) / threads;
  // One resident wave (4 CTAs of 256 threads per SM at 43 registers) looping grid-stride: measured on B200,
  // 148*4 CTAs run the sweep 12 % faster than 148*16 (multiple waves with tails) -- profiles/r01_sweep_grid.txt
  const int64_t cap = static_cast<int64_t>(kNumSMs) * 4;
  return static_cast<int>(want < cap ? (want > 0 ? want : 1) : cap);
}

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace dvgo

using namespace dvgo;

static int sweep_launch(const float* param_in, float* param_out, float* grad, float* exp_avg,
                        float* exp_avg_sq, const float* perlr, int X, int Y, int Z, int C,
                        int x_begin, int x_end, int tv, int tv_dense, float wx, float wy, float wz,
                        int masked, int step, float beta1, float beta2, float lr, float eps,
                        const SweepPeers& peers, dvgo_stream_t stream) {
  (void)wx;
  if (X <= 0 || Y <= 0 || Z <= 0 || C <= 0 || step <= 0) return DVGO_EINVAL;
  if (x_end < 0) x_end = X;
  if (x_begin < 0 || x_begin > x_end || x_end > X) return DVGO_EINVAL;
  if (x_begin == x_end) return 0;
  if (!param_in || !param_out || !grad || !exp_avg || !exp_avg_sq) return DVGO_EINVAL;
  if (tv && param_in == param_out) return DVGO_EINVAL;  // TV needs the old neighbours
  const float step_size = lr * sqrtf(1.f - powf(beta2, static_cast<float>(step))) /
                          (1.f - powf(beta1, static_cast<float>(step)));  // adam_upd_kernel.cu:72
  wy /= 6;  // total_variation_kernel.cu:45-47
  wz /= 6;
  const int64_t n = static_cast<int64_t>(x_end - x_begin) * Y * Z * C;
  const bool vec = (C % 4 == 0) && al16(param_in) && al16(param_out) && al16(grad) && al16(exp_avg) &&
                   al16(exp_avg_sq) && (!perlr || al16(perlr));
  cudaStream_t s = as_stream(stream);
#define SWEEP_ARGS param_in, param_out, grad, exp_avg, exp_avg_sq, perlr, X, Y, Z, C, x_begin, x_end, tv_dense, wy, wz, \
                   masked, step_size, beta1, beta2, eps, peers
  const bool eager = !masked || (tv && tv_dense);   // every element is updated: issue all loads up front
  const bool zvec = !vec && C == 1 && Z % 4 == 0 && al16(param_in) && al16(param_out) && al16(grad) && al16(exp_avg) &&
                    al16(exp_avg_sq) && (!perlr || al16(perlr));
  // one resident wave looping grid-stride (profiles/r01_sweep_grid.txt): grid = SMs x CTAs that fit per SM
#define SWEEP_ONE2(V, TV, EAGER, KZ, PEER, N)                                                         \
  do {                                                                                                \
    int occ = 4;                                                                                      \
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sweep_kernel<V, TV, EAGER, KZ, PEER>, 256, 0); \
    const int64_t want = ((N) + 255) / 256, cap = static_cast<int64_t>(kNumSMs) * (occ > 0 ? occ : 1); \
    const int blocks = static_cast<int>(want < cap ? (want > 0 ? want : 1) : cap);                    \
    sweep_kernel<V, TV, EAGER, KZ, PEER><<<blocks, 256, 0, s>>>(SWEEP_ARGS);                          \
  } while (0)
#define SWEEP_ONE(V, TV, EAGER, KZ, N)                                                                \
  do {                                                                                                \
    if (peers.n > 0) SWEEP_ONE2(V, TV, EAGER, KZ, true, N);                                           \
    else SWEEP_ONE2(V, TV, EAGER, KZ, false, N);                                                      \
  } while (0)
#define SWEEP_LAUNCH(V, KZ, N)                                                                        \
  do {                                                                                                \
    if (tv && eager) SWEEP_ONE(V, true, true, KZ, N);                                                 \
    else if (tv) SWEEP_ONE(V, true, false, KZ, N);                                                    \
    else if (eager) SWEEP_ONE(V, false, true, KZ, N);                                                 \
    else SWEEP_ONE(V, false, false, KZ, N);                                                           \
  } while (0)
  static const int bulk_env = [] { const char* e = getenv("DVGO_SWEEP_BULK"); return e ? atoi(e) : 1; }();
  // row segments: the fewest equal, voxel-aligned pieces of a z-row whose three stages fit 220 KB of shared memory
  const int row_f4 = Z * C / 4, G4 = C / 4;
  int nseg = 0;
  size_t rows_smem = 0;
  if (bulk_env && vec && tv && peers.n == 0 && !perlr) {
    for (int k = 1; k <= 8 && !nseg; ++k) {
      if (row_f4 % k || (row_f4 / k) % G4) continue;
      const size_t ctr = ((static_cast<size_t>(row_f4 / k + 2 * G4) * 16) + 127) & ~static_cast<size_t>(127);
      const size_t need = kRowStages * (ctr + 7 * static_cast<size_t>(row_f4 / k) * 16);
      if (need <= 220 * 1024) { nseg = k; rows_smem = need; }
    }
  }
  if (nseg) {
    // row-staged sweep: one persistent CTA per SM, whole rows (or row segments) through shared memory by bulk copies
    cudaError_t e = cudaFuncSetAttribute(sweep_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(rows_smem));
    if (e != cudaSuccess) return static_cast<int>(e);
    const int64_t n_units = static_cast<int64_t>(x_end - x_begin) * Y * nseg;
    const int blocks = static_cast<int>(n_units < kNumSMs ? n_units : kNumSMs);
    sweep_rows_kernel<<<blocks, kRowConsumers + 32, rows_smem, s>>>(param_in, param_out, grad, exp_avg, exp_avg_sq, X, Y, Z,
                                                                    C, x_begin, x_end, nseg, tv_dense, wy, wz, masked,
                                                                    step_size, beta1, beta2, eps);
  } else if (vec) SWEEP_LAUNCH(4, false, n / 4);
  else if (zvec) SWEEP_LAUNCH(4, true, n / 4);
  else SWEEP_LAUNCH(1, false, n);
#undef SWEEP_LAUNCH
#undef SWEEP_ONE
#undef SWEEP_ONE2
#undef SWEEP_ARGS
  return launch_status();
}

DVGO_API int dvgo_fused_sweep(const float* param_in, float* param_out, float* grad, float* exp_avg,
                              float* exp_avg_sq, const float* perlr, int X, int Y, int Z, int C,
                              int x_begin, int x_end, int tv, int tv_dense, float wx, float wy, float wz,
                              int masked, int step, float beta1, float beta2, float lr, float eps,
                              dvgo_stream_t stream) {
  SweepPeers none;
  none.n = 0;
  none.grad_mc = nullptr;
  none.pout_mc = nullptr;
  none.prefetch = 0;
  return sweep_launch(param_in, param_out, grad, exp_avg, exp_avg_sq, perlr, X, Y, Z, C, x_begin, x_end, tv, tv_dense,
                      wx, wy, wz, masked, step, beta1, beta2, lr, eps, none, stream);
}

DVGO_API int dvgo_fused_sweep_peer(const float* param_in, float* const* param_out_peers_host,
                                   float* const* grad_peers_host, float* param_out_multicast,
                                   const float* grad_multicast, int n_peers, int self_rank, float* exp_avg,
                                   float* exp_avg_sq, const float* perlr, int X, int Y, int Z, int C, int x_begin,
                                   int x_end, int tv, int tv_dense, float wx, float wy, float wz, int masked, int step,
                                   float beta1, float beta2, float lr, float eps, dvgo_stream_t stream) {
  if (!param_out_peers_host || !grad_peers_host || n_peers < 1 || n_peers > 8 || self_rank < 0 || self_rank >= n_peers)
    return DVGO_EINVAL;
  SweepPeers peers;
  peers.n = n_peers;
  peers.grad_mc = grad_multicast;
  peers.pout_mc = param_out_multicast;
  static const int prefetch_env = [] { const char* e = getenv("DVGO_PEER_PREFETCH"); return e ? atoi(e) : 0; }();
  peers.prefetch = prefetch_env;
  for (int r = 0; r < 8; ++r) {
    peers.grad[r] = r < n_peers ? grad_peers_host[r] : nullptr;
    peers.pout[r] = r < n_peers ? param_out_peers_host[r] : nullptr;
    if (r < n_peers && (!peers.grad[r] || !peers.pout[r])) return DVGO_EINVAL;
  }
  return sweep_launch(param_in, peers.pout[self_rank], const_cast<float*>(peers.grad[self_rank]), exp_avg, exp_avg_sq,
                      perlr, X, Y, Z, C, x_begin, x_end, tv, tv_dense, wx, wy, wz, masked, step, beta1, beta2, lr, eps,
                      peers, stream);
}

DVGO_API int dvgo_peer_barrier(int32_t* const* flags_peers_host, int n_peers, int self_rank, int epoch,
                               dvgo_stream_t stream) {
  if (!flags_peers_host || n_peers < 1 || n_peers > 8 || self_rank < 0 || self_rank >= n_peers) return DVGO_EINVAL;
  PeerFlags pf;
  pf.n = n_peers;
  pf.self = self_rank;
  for (int r = 0; r < 8; ++r) {
    pf.flags[r] = r < n_peers ? flags_peers_host[r] : nullptr;
    if (r < n_peers && !pf.flags[r]) return DVGO_EINVAL;
  }
  peer_barrier_kernel<<<1, 32, 0, as_stream(stream)>>>(pf, epoch);
  return launch_status();
}

DVGO_API int dvgo_adam_upd_peer(float* param, const float* const* grad_peers_host, const float* grad_multicast,
                                int n_peers, float* exp_avg, float* exp_avg_sq, int64_t N, int step, float beta1,
                                float beta2, float lr, float eps, dvgo_stream_t stream) {
  if (!param || !grad_peers_host || !exp_avg || !exp_avg_sq || n_peers < 1 || n_peers > 8 || N < 0 || step <= 0)
    return DVGO_EINVAL;
  if (N == 0) return 0;
  SmallPeers sp;
  sp.n = n_peers;
  sp.grad_mc = grad_multicast;
  for (int r = 0; r < 8; ++r) {
    sp.grad[r] = r < n_peers ? grad_peers_host[r] : nullptr;
    if (r < n_peers && !sp.grad[r]) return DVGO_EINVAL;
  }
  const float step_size = lr * sqrtf(1.f - powf(beta2, static_cast<float>(step))) /
                          (1.f - powf(beta1, static_cast<float>(step)));  // adam_upd_kernel.cu:72
  const int blocks = static_cast<int>((N + 255) / 256 < kNumSMs ? (N + 255) / 256 : kNumSMs);
  adam_peer_kernel<<<blocks, 256, 0, as_stream(stream)>>>(param, sp, exp_avg, exp_avg_sq, N, step_size, beta1, beta2, eps);
  return launch_status();
}

DVGO_API int dvgo_grid_ncdhw_to_cl(const float* src, float* dst, int C, int64_t G,
                                   dvgo_stream_t stream) {
  if (C <= 0 || G < 0 || !src || !dst) return DVGO_EINVAL;
  if (G == 0) return 0;
  ncdhw_to_cl_kernel<<<sweep_grid(G * C, 256), 256, 0, as_stream(stream)>>>(src, dst, C, G);
  return launch_status();
}

DVGO_API int dvgo_grid_cl_to_ncdhw(const float* src, float* dst, int C, int64_t G,
                                   dvgo_stream_t stream) {
  if (C <= 0 || G < 0 || !src || !dst) return DVGO_EINVAL;
  if (G == 0) return 0;
  cl_to_ncdhw_kernel<<<sweep_grid(G * C, 256), 256, 0, as_stream(stream)>>>(src, dst, C, G);
  return launch_status();
}
