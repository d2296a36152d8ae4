// fused_march.cu -- the fused per-ray kernels: ray_setup, march_fwd, march_bwd.
//
// One WARP owns one ray and walks it in chunks of 32 consecutive steps (lane = step), so
//   * every per-sample array is read / written coalesced (the reference's per-ray kernels stride by
//     the ray length across a warp, render_utils_kernel.cu:447-455),
//   * the transmittance product is a 5-step shuffle scan instead of a serial loop,
//   * the four data-dependent masks of DirectVoxGO.forward (lib/dvgo.py:444-447, 469-473, 478-484,
//     488-494: 16 boolean-index compactions, each a host sync) become lane predicates, and
//   * 32 neighbouring samples of a ray touch a ~16-voxel-long tube of the grid: the 8 trilinear
//     corner fetches of adjacent lanes hit the same L1 lines.
// Grid layout: density [X,Y,Z]; k0 channel-last [X,Y,Z,C], so one corner is C contiguous floats
// (48 B for C=12 = three 16-byte vector loads, the two z-corners are adjacent in memory) instead of
// C loads from C planes 16 MB apart as in the reference's NCDHW layout.
//
// Arithmetic follows the reference expression trees exactly where integer / boolean results depend
// on it (common.cuh); see include/dvgo_b200_fused.h for the data layout.
#include <climits>

#include "fused_scene.cuh"
#include "k0_tiles.cuh"
#include "tc_common.cuh"

namespace dvgo {

// ---- ray_setup -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fused_count_kernel(const float* __restrict__ rays_o,
                                                          const float* __restrict__ rays_d,
                                                          SceneArgs a, int n_rays,
                                                          float* __restrict__ t_min,
                                                          int32_t* __restrict__ n_steps) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rays) return;
  if (a.ndc) { t_min[r] = 0.f; n_steps[r] = a.ndc_samples; return; }
  const TMinMax t = slab_test(rays_o[3 * r], rays_o[3 * r + 1], rays_o[3 * r + 2], rays_d[3 * r],
                              rays_d[3 * r + 1], rays_d[3 * r + 2], a.xyz_min, a.xyz_max, a.near,
                              a.far);
  t_min[r] = t.t_min;
  n_steps[r] = static_cast<int32_t>(n_samples_of(t.t_min, t.t_max, a.stepdist));
}

// Single-CTA exclusive scan of int32 counts into ray_off[0..n] (ray_off[n] = total).
__global__ void __launch_bounds__(1024) fused_scan_kernel(const int32_t* __restrict__ in, int n,
                                                          int32_t* __restrict__ out) {
  __shared__ int32_t warp_sums[32];
  __shared__ int32_t carry_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  constexpr int kItems = 4;
  for (int base = 0; base < n; base += 1024 * kItems) {
    int32_t v[kItems];
    int32_t local = 0;
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
      const int i = base + tid * kItems + k;
      v[k] = (i < n) ? in[i] : 0;
      local += v[k];
    }
    int32_t incl = local;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int32_t up = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += up;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      int32_t ws = warp_sums[lane];
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int32_t up = __shfl_up_sync(0xffffffffu, ws, off);
        if (lane >= off) ws += up;
      }
      warp_sums[lane] = ws;
    }
    __syncthreads();
    int32_t run = carry_s + (wid ? warp_sums[wid - 1] : 0) + (incl - local);  // exclusive prefix
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
      const int i = base + tid * kItems + k;
      if (i < n) out[i] = run;
      run += v[k];
    }
    __syncthreads();
    if (tid == 1023) carry_s = run;
    __syncthreads();
  }
  if (tid == 0) out[n] = carry_s;
}

// ---- march_fwd -------------------------------------------------------------------------------------
template <int C>
__device__ __forceinline__ void gather_k0(const float* __restrict__ k0, const Corner8& cn,
                                          float* __restrict__ out) {
  float acc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c] = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (!cn.ok(k)) continue;
    const float* __restrict__ v = k0 + static_cast<int64_t>(cn.off(k)) * C;
    const float wk = cn.w(k);
    if (C % 4 == 0) {
#pragma unroll
      for (int q = 0; q < C / 4; ++q) {
        const float4 f = __ldg(reinterpret_cast<const float4*>(v) + q);
        acc[4 * q + 0] = fma_(f.x, wk, acc[4 * q + 0]);
        acc[4 * q + 1] = fma_(f.y, wk, acc[4 * q + 1]);
        acc[4 * q + 2] = fma_(f.z, wk, acc[4 * q + 2]);
        acc[4 * q + 3] = fma_(f.w, wk, acc[4 * q + 3]);
      }
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) acc[c] = fma_(__ldg(v + c), wk, acc[c]);
    }
  }
  if (C % 4 == 0) {
#pragma unroll
    for (int q = 0; q < C / 4; ++q)
      reinterpret_cast<float4*>(out)[q] =
          make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
  } else {
#pragma unroll
    for (int c = 0; c < C; ++c) out[c] = acc[c];
  }
}

// kExactT = false (default): the transmittance of a 32-sample chunk is a shuffle product scan in double, rounded to
// float once per output (T / weights rel 5e-6 vs the reference, whose `float T_cum` is re-rounded after EVERY sample,
// render_utils_kernel.cu:447-451).  kExactT = true (scene.exact_transmittance): the warp replays that per-sample
// recurrence in lock-step over the live lanes of the chunk (all lanes hold the same T_cum; lane j keeps the value it
// saw at its own sample) -> T, weights, alphainv_last, the early-stop index and hence the survivor SET are bit-exact
// with the reference kernels, at one dependent cvt-DMUL-cvt step per live sample.
template <int C, bool kExactT>
__global__ void __launch_bounds__(256, 4) march_fwd_kernel(
    const float* __restrict__ rays_o, const float* __restrict__ rays_d, SceneArgs a,
    const float* __restrict__ density, const float* __restrict__ k0, int n_rays,
    const float* __restrict__ t_min, const int32_t* __restrict__ n_steps,
    const int32_t* __restrict__ ray_off, int64_t slot_cap, int64_t surv_cap,
    float* __restrict__ slot_alpha, float* __restrict__ slot_T, float* __restrict__ slot_expd,
    int32_t* __restrict__ slot_code, float* __restrict__ feat, int32_t* __restrict__ s_ray,
    int32_t* __restrict__ s_slot, float* __restrict__ s_weight, float* __restrict__ alphainv_last,
    int32_t* __restrict__ counters, float4* __restrict__ s_pos) {
  const SceneDev sc = load_scene(a);
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const bool use_thres = sc.thres > 0.f;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < n_rays; r += gridDim.x * wpb) {
    const int n = n_steps[r];
    const int64_t off = ray_off[r];
    const RayGeom g = ray_geom(sc, rays_o, rays_d, r, t_min[r]);
    double carry = 1.0;
    float last_T = 1.f;   // kExactT: the running float T_cum itself
    bool stopped = false;
    for (int base = 0; base < n; base += 32) {
      const int i = base + lane;
      const bool valid = i < n;
      const int64_t slot = off + i;
      float px, py, pz;
      sample_point(sc, g, i, px, py, pz);
      bool live = valid && !stopped && !out_of_bbox(px, py, pz, sc.lo, sc.hi);  // lib/dvgo.py:444-447
      live = live && occupancy(sc, px, py, pz);                                  // :469-473
      float alpha = 0.f, e = 0.f;
      Corner8 cn;
      float fx = 0.f, fy = 0.f, fz = 0.f;
      if (live) {
        voxel_coords(sc, px, py, pz, fx, fy, fz);
        cn = corner8_idx(sc, fx, fy, fz);
        float dens = 0.f;  // :476 trilinear density (ATen accumulation order)
        float dv[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) dv[k] = __ldg(density + (cn.ok(k) ? cn.off(k) : 0));   // unconditional: eight loads in flight
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (cn.ok(k)) dens = fma_(dv[k], cn.w(k), dens);
        e = expf(fadd(dens, sc.act_shift));                     // render_utils_kernel.cu:366
        alpha = fsub(1.f, powf(fadd(1.f, e), -sc.interval));    // :368
        if (use_thres) live = alpha > sc.thres;                 // lib/dvgo.py:478-484
      }
      // transmittance: exclusive product of (1 - alpha + 1e-10) over the samples still alive
      const double f = live ? ((1.0 - static_cast<double>(alpha)) + 1e-10) : 1.0;
      float T_before, T_after = 0.f;
      unsigned hits;
      int first;
      [[maybe_unused]] double incl = 1.0;
      if constexpr (kExactT) {
        T_before = last_T;
        first = 32;
        for (unsigned m = __ballot_sync(0xffffffffu, live); m; m &= m - 1) {   // live lanes, near to far (warp-uniform)
          const int j = __ffs(m) - 1;
          const double fj = __shfl_sync(0xffffffffu, f, j);
          if (lane == j) T_before = last_T;                                            // render_utils_kernel.cu:448
          last_T = __double2float_rn(__dmul_rn(static_cast<double>(last_T), fj));      // :450
          if (static_cast<double>(last_T) < 1e-3) { first = j; break; }                // :451
        }
        hits = first < 32 ? 1u : 0u;
      } else {
        incl = f;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const double up = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl *= up;
        }
        double excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = 1.0;
        T_before = static_cast<float>(carry * excl);
        T_after = static_cast<float>(carry * incl);
        const bool hit = live && (static_cast<double>(T_after) < 1e-3);  // render_utils_kernel.cu:451
        hits = __ballot_sync(0xffffffffu, hit);
        first = hits ? (__ffs(hits) - 1) : 32;
      }
      const bool in_scan = live && lane <= first;            // inside [i_start, i_end)
      const float w = fmul(T_before, alpha);                  // :449
      const bool surv = in_scan && (!use_thres || w > sc.thres);  // lib/dvgo.py:488-494
      const unsigned smask = __ballot_sync(0xffffffffu, surv);
      int idx4 = -1;
      if (smask) {
        int base4 = 0;
        if (lane == 0) base4 = atomicAdd(counters, __popc(smask));
        base4 = __shfl_sync(0xffffffffu, base4, 0);
        if (surv) idx4 = base4 + __popc(smask & ((1u << lane) - 1u));
      }
      if (!slot_code) {
        // forward-only call (rendering): no backward pass will read the per-slot record
      } else if (valid && slot < slot_cap) {
        slot_code[slot] = surv ? idx4 : (in_scan ? -1 : -2);
        if (in_scan) { slot_alpha[slot] = alpha; slot_T[slot] = T_before; slot_expd[slot] = e; }
      } else if (valid) {
        counters[1] = 1;  // capacity overflow (cannot happen with dvgo_fused_max_steps sizing)
      }
      if (surv) {
        if (idx4 < surv_cap) {
          s_ray[idx4] = r;
          s_slot[idx4] = static_cast<int32_t>(slot);
          s_weight[idx4] = w;
          if (s_pos) s_pos[idx4] = make_float4(fx, fy, fz, __int_as_float(r));   // the k0 kernels' per-survivor record
          if (k0) gather_k0<C>(k0, cn, feat + static_cast<int64_t>(idx4) * C);  // lib/dvgo.py:509
        } else {
          counters[1] = 1;
        }
      }
      if constexpr (kExactT) {
        if (hits) stopped = true;  // later chunks have no live lane: last_T stays the reference's final T_cum
      } else if (hits) {
        stopped = true;  // later chunks only mark their slots as culled
        last_T = __shfl_sync(0xffffffffu, T_after, first);
      } else if (!stopped) {
        carry = carry * __shfl_sync(0xffffffffu, incl, 31);
        last_T = static_cast<float>(carry);
      }
    }
    if (lane == 0) alphainv_last[r] = last_T;  // render_utils_kernel.cu:457
  }
}

// ---- march_bwd -------------------------------------------------------------------------------------
template <int C>
__device__ __forceinline__ void scatter_k0(float* __restrict__ gk0, const Corner8& cn,
                                           const float* __restrict__ df) {
  float d[C];
  if (C % 4 == 0) {
#pragma unroll
    for (int q = 0; q < C / 4; ++q) {
      const float4 f = __ldg(reinterpret_cast<const float4*>(df) + q);
      d[4 * q] = f.x; d[4 * q + 1] = f.y; d[4 * q + 2] = f.z; d[4 * q + 3] = f.w;
    }
  } else {
#pragma unroll
    for (int c = 0; c < C; ++c) d[c] = __ldg(df + c);
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (!cn.ok(k)) continue;
    float* __restrict__ dst = gk0 + static_cast<int64_t>(cn.off(k)) * C;
    const float w = cn.w(k);
    if (C % 4 == 0) {
#pragma unroll
      for (int q = 0; q < C / 4; ++q) {
        // one 16-byte vector reduction (red.global.add.v4.f32, sm_90+) instead of four scalar atomics
        atomicAdd(reinterpret_cast<float4*>(dst) + q,
                  make_float4(fmul(w, d[4 * q]), fmul(w, d[4 * q + 1]), fmul(w, d[4 * q + 2]),
                              fmul(w, d[4 * q + 3])));
      }
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) atomicAdd(dst + c, fmul(w, d[c]));
    }
  }
}

template <int C>
__global__ void __launch_bounds__(256, 3) march_bwd_kernel(
    const float* __restrict__ rays_o, const float* __restrict__ rays_d, SceneArgs a, int n_rays,
    const float* __restrict__ t_min, const int32_t* __restrict__ n_steps,
    const int32_t* __restrict__ ray_off, const float* __restrict__ slot_alpha,
    const float* __restrict__ slot_T, const float* __restrict__ slot_expd,
    const int32_t* __restrict__ slot_code, const float* __restrict__ d_feat,
    const float* __restrict__ d_w, const float* __restrict__ alphainv_last,
    const float* __restrict__ g_last, float* __restrict__ grad_density,
    float* __restrict__ grad_k0) {
  const SceneDev sc = load_scene(a);
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < n_rays; r += gridDim.x * wpb) {
    const int n = n_steps[r];
    const int64_t off = ray_off[r];
    const RayGeom g = ray_geom(sc, rays_o, rays_d, r, t_min[r]);
    float back = fmul(g_last[r], alphainv_last[r]);  // render_utils_kernel.cu:525
    // far to near; lane 0 = farthest sample of the chunk
    for (int hi = n; hi > 0; hi -= 32) {
      const int i = hi - 1 - lane;
      const bool valid = i >= 0;
      const int64_t slot = off + i;
      const int code = valid ? slot_code[slot] : -2;
      const bool in_scan = code >= -1;
      float alpha = 0.f, T = 0.f, e = 0.f, gw = 0.f;
      if (in_scan) {
        alpha = slot_alpha[slot]; T = slot_T[slot]; e = slot_expd[slot];
        if (code >= 0) gw = d_w[code];  // samples dropped by the weight mask carry dL/dw = 0
      }
      const float term = in_scan ? fmul(gw, fmul(T, alpha)) : 0.f;  // gw * weight, :528
      float incl = term;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
      }
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = 0.f;
      const float back_i = back + excl;
      back += __shfl_sync(0xffffffffu, incl, 31);
      if (!in_scan) continue;
      // alpha2weight backward (:527) then raw2alpha backward (:404)
      const float g_alpha = static_cast<float>(
          static_cast<double>(fmul(gw, T)) -
          static_cast<double>(back_i) / (static_cast<double>(fsub(1.f, alpha)) + 1e-10));
      const double mm = fmin(static_cast<double>(e), 1e10);
      const double pw = static_cast<double>(powf(fadd(1.f, e), fsub(-sc.interval, 1.f)));
      const float g_dens = static_cast<float>(mm * pw * static_cast<double>(sc.interval) *
                                              static_cast<double>(g_alpha));
      float px, py, pz;
      sample_point(sc, g, i, px, py, pz);
      const Corner8 cn = corner8(sc, px, py, pz);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (cn.ok(k)) atomicAdd(grad_density + cn.off(k), fmul(cn.w(k), g_dens));
      if (code >= 0 && grad_k0) scatter_k0<C>(grad_k0, cn, d_feat + static_cast<int64_t>(code) * C);
    }
  }
}


// ---- k0 gather / scatter over the compacted survivor stream -------------------------------------------
// The transmittance scan makes march_fwd / march_bwd walk a ray chunk by chunk; the k0 traffic (8 corners x
// C floats per survivor -- 6x the density traffic) does not depend on the scan at all.  These two kernels do
// it with one thread per (survivor, 16-byte channel group): every load / reduction of a thread is
// independent, so many more are in flight per SM than inside the per-ray kernels.
template <int C>
__global__ void __launch_bounds__(256) k0_gather_kernel(
    const float* __restrict__ rays_o, const float* __restrict__ rays_d, SceneArgs a,
    const float* __restrict__ k0, const float* __restrict__ t_min, const int32_t* __restrict__ ray_off,
    const int32_t* __restrict__ s_ray, const int32_t* __restrict__ s_slot,
    const int32_t* __restrict__ counters, int64_t surv_cap, float* __restrict__ feat) {
  constexpr int G = (C % 4 == 0) ? C / 4 : 1;   // channel groups per sample
  constexpr int W = (C % 4 == 0) ? 4 : C;       // floats per group
  const SceneDev sc = load_scene(a);
  int64_t n = counters[0];
  if (n > surv_cap) n = surv_cap;
  const int64_t total = n * G;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t p = i / G;
    const int q = static_cast<int>(i - p * G);
    const int r = s_ray[p];
    const int step = s_slot[p] - ray_off[r];
    const RayGeom g = ray_geom(sc, rays_o, rays_d, r, t_min[r]);
    float px, py, pz;
    sample_point(sc, g, step, px, py, pz);
    const Corner8 cn = corner8(sc, px, py, pz);
    float acc[W];
#pragma unroll
    for (int c = 0; c < W; ++c) acc[c] = 0.f;
    float4 f8[(C % 4 == 0) ? 8 : 1];
    if (C % 4 == 0) {
#pragma unroll
      for (int k = 0; k < 8; ++k)      // unconditional (see k0_gather_tiles_kernel): eight loads in flight
        f8[k] = __ldg(reinterpret_cast<const float4*>(k0 + static_cast<int64_t>(cn.ok(k) ? cn.off(k) : 0) * C + q * W));
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (!cn.ok(k)) continue;
      const float* __restrict__ v = k0 + static_cast<int64_t>(cn.off(k)) * C + q * W;
      const float wk = cn.w(k);
      if (C % 4 == 0) {
        const float4 f = f8[(C % 4 == 0) ? k : 0];
        acc[0] = fma_(f.x, wk, acc[0]); acc[1] = fma_(f.y, wk, acc[1]);
        acc[2 % W] = fma_(f.z, wk, acc[2 % W]); acc[3 % W] = fma_(f.w, wk, acc[3 % W]);
      } else {
#pragma unroll
        for (int c = 0; c < W; ++c) acc[c] = fma_(__ldg(v + c), wk, acc[c]);
      }
    }
    float* __restrict__ o = feat + p * C + q * W;
    if (C % 4 == 0) *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2 % W], acc[3 % W]);
    else {
#pragma unroll
      for (int c = 0; c < W; ++c) o[c] = acc[c];
    }
  }
}

// k0_gather with the rgbnet's X~ tiles as output (tc_common.cuh: tiles_for; include/dvgo_b200_fused.h): row p of
// the tile stream = [k0 features (C) | pe[s_ray[p]] (pe_stride) | 0 ...] as saturated fp16 in the tensor core's operand
// layout, K1 columns (the row itself: k0_tiles.cuh).  Rows from the survivor count to the next multiple of 256 are
// written as zeros (the backward kernel consumes tile pairs).  A warp holds 32/G whole survivors (30 active lanes for
// G = 3).  The fused step / renderer use mlp_fwd_gather_kernel (fused_mlp.cu) instead, which builds the same rows
// straight into the shared-memory tile; this kernel serves callers that want the tiles in global memory.
template <int C>
__global__ void __launch_bounds__(256, 4) k0_gather_tiles_kernel(
    const float* __restrict__ rays_o, const float* __restrict__ rays_d, SceneArgs a,
    const float* __restrict__ k0, const float* __restrict__ t_min, const int32_t* __restrict__ ray_off,
    const int32_t* __restrict__ s_ray, const int32_t* __restrict__ s_slot,
    const int32_t* __restrict__ counters, int64_t surv_cap, const float4* __restrict__ s_pos,
    const uint8_t* __restrict__ pe16, int pe_stride, int K1, uint8_t* __restrict__ xt) {
  constexpr int G = K0TileShape<C>::G;           // threads per survivor
  constexpr int SPW = K0TileShape<C>::SPW;       // survivors per warp
  const SceneDev sc = load_scene(a);
  int64_t n = counters[0];
  if (n > surv_cap) n = surv_cap;
  const int64_t rows = (n + 255) / 256 * 256;
  const uint32_t xb = tc::tile_bytes(128, K1);
  const int used_chunks = (C + pe_stride + 7) >> 3;    // chunks with any non-padding column
  const int lane = threadIdx.x & 31;
  const int q = lane % G, sub = lane / G;
  const int64_t warp0 = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  // the per-survivor record of the NEXT grid-stride iteration is requested one iteration ahead: record -> corner
  // addresses -> corner loads is a chain of two dependent memory round trips per iteration otherwise
  float4 rec_next = make_float4(0.f, 0.f, 0.f, 0.f);
  if (s_pos && sub < SPW && warp0 * SPW + sub < n) rec_next = __ldg(s_pos + warp0 * SPW + sub);
  for (int64_t base = warp0 * SPW; base < rows; base += n_warps * SPW) {   // warp-uniform trip count
    const int64_t p = base + sub;
    const bool active = sub < SPW && p < rows;
    const bool live = active && p < n;
    uint8_t* __restrict__ trow = xt + (p >> 7) * xb + tc::tile_off(static_cast<int>(p & 127), 0, K1);
    const float4 rec = rec_next;
    const int64_t p_next = p + n_warps * SPW;
    if (s_pos && sub < SPW && p_next < n) rec_next = __ldg(s_pos + p_next);
    int r = 0;
    Corner8 cn;
    cn.valid = 0u;
    if (live) {
      if (s_pos) {       // march_fwd's record: continuous voxel coordinates + ray index, one 16-byte load
        r = __float_as_int(rec.w);
        cn = corner8_idx(sc, rec.x, rec.y, rec.z);
      } else {
        r = s_ray[p];
        const int step = s_slot[p] - ray_off[r];
        const RayGeom g = ray_geom(sc, rays_o, rays_d, r, t_min[r]);
        float px, py, pz;
        sample_point(sc, g, step, px, py, pz);
        cn = corner8(sc, px, py, pz);
      }
    }
    k0_tile_row<C, false>(k0, cn, r, active, live, pe16, K1, used_chunks, trow, q);
  }
}

// Warp-aggregated reductions.  A thread owns (survivor p, 16-byte channel group q); lane + G is survivor p + 1, the
// next sample of the same ray most of the time, and at half a voxel per step it lies in the SAME base cell about 40 %
// of the time -- all eight of its reductions then go to the same eight addresses.  Such runs of equal base cell are
// merged in registers before anything leaves the SM: the issuing lane pulls its (up to two) predecessors' inputs
// (voxel coordinates + gradient group: 7 shuffles each), recomputes their corner weights and adds the products, and
// the predecessors issue nothing.  Runs are cut into triples from their start (a run is at most 4 samples long at this
// step size); lanes of different warps never merge.  The kernel is bound by the L2 reduction rate (one 32-byte sector
// per slice per clock: profiles/r01_sweep_grid.txt), so fewer red.global.add.v4.f32 is what buys time.
template <int C>
__global__ void __launch_bounds__(256) k0_scatter_kernel(
    const float* __restrict__ rays_o, const float* __restrict__ rays_d, SceneArgs a,
    const float* __restrict__ t_min, const int32_t* __restrict__ ray_off,
    const int32_t* __restrict__ s_ray, const int32_t* __restrict__ s_slot,
    const int32_t* __restrict__ counters, int64_t surv_cap, const float4* __restrict__ s_pos,
    const float* __restrict__ d_feat, float* __restrict__ grad_k0) {
  constexpr int G = (C % 4 == 0) ? C / 4 : 1;
  constexpr int W = (C % 4 == 0) ? 4 : C;
  const SceneDev sc = load_scene(a);
  int64_t n = counters[0];
  if (n > surv_cap) n = surv_cap;
  const int64_t total = n * G;
  const int lane = threadIdx.x & 31;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i - lane < total;   // warp-uniform
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const bool valid = i < total;
    const int64_t p = valid ? i / G : 0;
    const int q = static_cast<int>(i - (i / G) * G);
    float fx = 0.f, fy = 0.f, fz = 0.f;
    if (valid) {
      if (s_pos) {         // march_fwd's per-survivor record (continuous voxel coordinates)
        const float4 rec = __ldg(s_pos + p);
        fx = rec.x; fy = rec.y; fz = rec.z;
      } else {
        const int r = s_ray[p];
        const int step = s_slot[p] - ray_off[r];
        const RayGeom g = ray_geom(sc, rays_o, rays_d, r, t_min[r]);
        float px, py, pz;
        sample_point(sc, g, step, px, py, pz);
        voxel_coords(sc, px, py, pz, fx, fy, fz);
      }
    }
    Corner8 cn = corner8_idx(sc, fx, fy, fz);
    if (!valid) cn.valid = 0u;
    float d[W];
    if (C % 4 == 0) {
      float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid) f = __ldg(reinterpret_cast<const float4*>(d_feat + p * C + q * W));
      d[0] = f.x; d[1] = f.y; d[2 % W] = f.z; d[3 % W] = f.w;
      // ---- runs of equal base cell among the lanes of this channel group ----
      constexpr unsigned kFull = 0xffffffffu;
      const int base = valid ? cn.base : INT_MIN;       // INT_MIN never equals a real (even padded) base index
      const int pb = __shfl_up_sync(kFull, base, G);
      const int nb = __shfl_down_sync(kFull, base, G);
      const bool head = !(lane >= G && pb == base);
      const bool last = !(lane + G < 32 && nb == base);
      const unsigned heads = __ballot_sync(kFull, head);
      unsigned mine = 0u;                                // lanes of my channel group at or below me
#pragma unroll
      for (int l = 0; l < 32; ++l)
        if (l % G == 0) mine |= 1u << l;
      mine = (mine << (lane % G)) & (0xffffffffu >> (31 - lane));
      const int head_lane = 31 - __clz(heads & mine);   // my own lane if I am a head
      const int pos = (lane - head_lane) / G;
      const int npull = pos % 3;                         // predecessors an issuing lane absorbs
      const bool issue = valid && (last || npull == 2);
      float acc[8][4];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float wk = cn.w(k);
        acc[k][0] = fmul(wk, d[0]); acc[k][1] = fmul(wk, d[1]); acc[k][2] = fmul(wk, d[2 % W]); acc[k][3] = fmul(wk, d[3 % W]);
      }
#pragma unroll
      for (int back = 1; back <= 2; ++back) {            // every lane shuffles; only issuing lanes with enough run use it
        const float qx = __shfl_up_sync(kFull, fx, back * G), qy = __shfl_up_sync(kFull, fy, back * G),
                    qz = __shfl_up_sync(kFull, fz, back * G);
        float e[4];
        e[0] = __shfl_up_sync(kFull, d[0], back * G); e[1] = __shfl_up_sync(kFull, d[1], back * G);
        e[2] = __shfl_up_sync(kFull, d[2 % W], back * G); e[3] = __shfl_up_sync(kFull, d[3 % W], back * G);
        if (issue && npull >= back) {
          Corner8 pc = cn;                               // same base cell, same validity: only the weights differ
          const Tri t = tri_from_index(qx, qy, qz);
          pc.wx0 = t.wx0; pc.wx1 = t.wx1; pc.wy0 = t.wy0; pc.wy1 = t.wy1; pc.wz0 = t.wz0; pc.wz1 = t.wz1;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float wk = pc.w(k);
            acc[k][0] = fma_(wk, e[0], acc[k][0]); acc[k][1] = fma_(wk, e[1], acc[k][1]);
            acc[k][2] = fma_(wk, e[2], acc[k][2]); acc[k][3] = fma_(wk, e[3], acc[k][3]);
          }
        }
      }
      if (issue) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (!cn.ok(k)) continue;
          float* __restrict__ dst = grad_k0 + static_cast<int64_t>(cn.off(k)) * C + q * W;
          atomicAdd(reinterpret_cast<float4*>(dst), make_float4(acc[k][0], acc[k][1], acc[k][2], acc[k][3]));
        }
      }
    } else if (valid) {
      const float* __restrict__ src = d_feat + p * C + q * W;
#pragma unroll
      for (int c = 0; c < W; ++c) d[c] = __ldg(src + c);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (!cn.ok(k)) continue;
        float* __restrict__ dst = grad_k0 + static_cast<int64_t>(cn.off(k)) * C + q * W;
        const float wk = cn.w(k);
#pragma unroll
        for (int c = 0; c < W; ++c) atomicAdd(dst + c, fmul(wk, d[c]));
      }
    }
  }
}

static inline int ray_blocks(int n_rays, int wpb) {
  const int64_t want = (static_cast<int64_t>(n_rays) + wpb - 1) / wpb;
  const int64_t cap = static_cast<int64_t>(kNumSMs) * 32;
  return static_cast<int>(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace dvgo

using namespace dvgo;

DVGO_API int dvgo_fused_max_steps(const dvgo_scene_t* s) {
  if (!s) return DVGO_EINVAL;
  if (s->ndc) return s->ndc_samples;
  const float c = ceilf((s->far - s->near) / s->stepdist);
  return c > 1.f ? static_cast<int>(c) : 1;
}

DVGO_API int dvgo_fused_ray_setup(const float* rays_o, const float* rays_d, const dvgo_scene_t* scene,
                                  int n_rays, float* t_min, int32_t* n_steps, int32_t* ray_off,
                                  dvgo_stream_t stream) {
  if (n_rays < 0 || !scene || !ray_off) return DVGO_EINVAL;
  if (n_rays > 0 && (!rays_o || !rays_d || !t_min || !n_steps)) return DVGO_EINVAL;
  cudaStream_t s = as_stream(stream);
  if (n_rays > 0)
    fused_count_kernel<<<blocks_for(n_rays, 256), 256, 0, s>>>(rays_o, rays_d, to_args(scene), n_rays,
                                                               t_min, n_steps);
  fused_scan_kernel<<<1, 1024, 0, s>>>(n_steps, n_rays, ray_off);
  return launch_status(n_rays > 0 ? 2 : 1);
}

DVGO_API int dvgo_fused_march_fwd(const float* rays_o, const float* rays_d, const dvgo_scene_t* scene,
                                  const float* density, const float* k0_cl, int n_rays,
                                  const float* t_min, const int32_t* n_steps, const int32_t* ray_off,
                                  int64_t slot_cap, int64_t surv_cap, float* slot_alpha, float* slot_T,
                                  float* slot_expd, int32_t* slot_code, float* feat, int32_t* s_ray,
                                  int32_t* s_slot, float* s_weight, float* alphainv_last,
                                  int32_t* counters, float* s_pos, dvgo_stream_t stream) {
  if (n_rays < 0 || !scene) return DVGO_EINVAL;
  if (n_rays == 0) return 0;
  if (!rays_o || !rays_d || !density || !t_min || !n_steps || !ray_off || !s_ray || !s_slot || !s_weight ||
      !alphainv_last || !counters || (k0_cl && !feat))
    return DVGO_EINVAL;
  // the four slot arrays come together or not at all (all NULL = forward-only: nothing is kept for march_bwd)
  const int n_slot_arrays = (slot_alpha != nullptr) + (slot_T != nullptr) + (slot_expd != nullptr) + (slot_code != nullptr);
  if (n_slot_arrays != 0 && n_slot_arrays != 4) return DVGO_EINVAL;
  // 2 rays (warps) per CTA: a CTA holds its slot until its longest ray ends, so small CTAs pack the SMs better
  // (measured on B200, 8192 rays: 8 warps/CTA 0.205 + 0.243 ms for the two march stages, 2 warps/CTA 0.198 + 0.231 ms)
  const int wpb = 2;
  if (scene->exact_transmittance) {
    DVGO_DISPATCH_C(scene->C, (march_fwd_kernel<kC, true><<<ray_blocks(n_rays, wpb), wpb * 32, 0,
                                                            as_stream(stream)>>>(
        rays_o, rays_d, to_args(scene), density, k0_cl, n_rays, t_min, n_steps, ray_off, slot_cap,
        surv_cap, slot_alpha, slot_T, slot_expd, slot_code, feat, s_ray, s_slot, s_weight,
        alphainv_last, counters, reinterpret_cast<float4*>(s_pos))));
    return launch_status();
  }
  DVGO_DISPATCH_C(scene->C, (march_fwd_kernel<kC, false><<<ray_blocks(n_rays, wpb), wpb * 32, 0,
                                                             as_stream(stream)>>>(
      rays_o, rays_d, to_args(scene), density, k0_cl, n_rays, t_min, n_steps, ray_off, slot_cap,
      surv_cap, slot_alpha, slot_T, slot_expd, slot_code, feat, s_ray, s_slot, s_weight,
      alphainv_last, counters, reinterpret_cast<float4*>(s_pos))));
  return launch_status();
}

DVGO_API int dvgo_fused_march_bwd(const float* rays_o, const float* rays_d, const dvgo_scene_t* scene,
                                  int n_rays, const float* t_min, const int32_t* n_steps,
                                  const int32_t* ray_off, const float* slot_alpha, const float* slot_T,
                                  const float* slot_expd, const int32_t* slot_code, const float* d_feat,
                                  const float* d_w, const float* alphainv_last, const float* g_last,
                                  float* grad_density, float* grad_k0_cl, dvgo_stream_t stream) {
  if (n_rays < 0 || !scene) return DVGO_EINVAL;
  if (n_rays == 0) return 0;
  if (!rays_o || !rays_d || !t_min || !n_steps || !ray_off || !slot_alpha || !slot_T || !slot_expd ||
      !slot_code || !d_w || !alphainv_last || !g_last || !grad_density || (grad_k0_cl && !d_feat))
    return DVGO_EINVAL;
  const int wpb = 2;
  DVGO_DISPATCH_C(scene->C, (march_bwd_kernel<kC><<<ray_blocks(n_rays, wpb), wpb * 32, 0,
                                                    as_stream(stream)>>>(
      rays_o, rays_d, to_args(scene), n_rays, t_min, n_steps, ray_off, slot_alpha, slot_T, slot_expd,
      slot_code, d_feat, d_w, alphainv_last, g_last, grad_density, grad_k0_cl)));
  return launch_status();
}

DVGO_API int dvgo_fused_k0_gather(const float* rays_o, const float* rays_d, const dvgo_scene_t* scene,
                                  const float* k0_cl, const float* t_min, const int32_t* ray_off,
                                  const int32_t* s_ray, const int32_t* s_slot, const int32_t* counters,
                                  int64_t surv_cap, float* feat, dvgo_stream_t stream) {
  if (!scene || surv_cap < 0) return DVGO_EINVAL;
  if (surv_cap == 0) return 0;
  if (!rays_o || !rays_d || !k0_cl || !t_min || !ray_off || !s_ray || !s_slot || !counters || !feat)
    return DVGO_EINVAL;
  const int groups = (scene->C % 4 == 0) ? scene->C / 4 : 1;
  const int64_t want = (surv_cap * groups + 255) / 256;
  const int blocks = static_cast<int>(want < kNumSMs * 32 ? want : kNumSMs * 32);
  DVGO_DISPATCH_C(scene->C, (k0_gather_kernel<kC><<<blocks, 256, 0, as_stream(stream)>>>(
      rays_o, rays_d, to_args(scene), k0_cl, t_min, ray_off, s_ray, s_slot, counters, surv_cap, feat)));
  return launch_status();
}

DVGO_API int dvgo_fused_k0_gather_tiles(const float* rays_o, const float* rays_d, const dvgo_scene_t* scene,
                                        const float* k0_cl, const float* t_min, const int32_t* ray_off,
                                        const int32_t* s_ray, const int32_t* s_slot, const int32_t* counters,
                                        int64_t surv_cap, const float* s_pos, const void* pe_rows16,
                                        int pe_stride, void* xt, dvgo_stream_t stream) {
  if (!scene || surv_cap < 0 || pe_stride < 1 || scene->C + pe_stride > 64) return DVGO_EINVAL;
  if (surv_cap == 0) return 0;
  if (!rays_o || !rays_d || !k0_cl || !t_min || !ray_off || !s_ray || !s_slot || !counters || !pe_rows16 || !xt)
    return DVGO_EINVAL;
  const int K1 = ((scene->C + pe_stride + 15) / 16) * 16;
  const int groups = (scene->C % 4 == 0) ? scene->C / 4 : 1;
  const int64_t rows = (surv_cap + 255) / 256 * 256;
  const int64_t want = (rows / (32 / groups) + 1 + 7) / 8;      // 8 warps per CTA, 32/G survivors per warp
  const int blocks = static_cast<int>(want < kNumSMs * 32 ? want : kNumSMs * 32);
  DVGO_DISPATCH_C(scene->C, (k0_gather_tiles_kernel<kC><<<blocks, 256, 0, as_stream(stream)>>>(
      rays_o, rays_d, to_args(scene), k0_cl, t_min, ray_off, s_ray, s_slot, counters, surv_cap,
      reinterpret_cast<const float4*>(s_pos), static_cast<const uint8_t*>(pe_rows16), pe_stride, K1,
      static_cast<uint8_t*>(xt))));
  return launch_status();
}

DVGO_API int dvgo_fused_k0_scatter(const float* rays_o, const float* rays_d, const dvgo_scene_t* scene,
                                   const float* t_min, const int32_t* ray_off, const int32_t* s_ray,
                                   const int32_t* s_slot, const int32_t* counters, int64_t surv_cap,
                                   const float* s_pos, const float* d_feat, float* grad_k0_cl,
                                   dvgo_stream_t stream) {
  if (!scene || surv_cap < 0) return DVGO_EINVAL;
  if (surv_cap == 0) return 0;
  if (!rays_o || !rays_d || !t_min || !ray_off || !s_ray || !s_slot || !counters || !d_feat || !grad_k0_cl)
    return DVGO_EINVAL;
  const int groups = (scene->C % 4 == 0) ? scene->C / 4 : 1;
  const int64_t want = (surv_cap * groups + 255) / 256;
  const int blocks = static_cast<int>(want < kNumSMs * 32 ? want : kNumSMs * 32);
  DVGO_DISPATCH_C(scene->C, (k0_scatter_kernel<kC><<<blocks, 256, 0, as_stream(stream)>>>(
      rays_o, rays_d, to_args(scene), t_min, ray_off, s_ray, s_slot, counters, surv_cap,
      reinterpret_cast<const float4*>(s_pos), d_feat, grad_k0_cl)));
  return launch_status();
}
