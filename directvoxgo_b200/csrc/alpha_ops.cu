// alpha_ops.cu -- density -> alpha and alpha -> transmittance weights (rows a8, a9).
// Reference: lib/cuda/render_utils_kernel.cu:358-561.
//
// The reference's alpha2weight runs ONE THREAD PER RAY through a serial loop of dependent
// float*double multiplies (:447-455), with stride-M uncoalesced accesses across the warp.
// Here one WARP owns a ray: 32 consecutive samples are loaded coalesced and the warp replays the
// reference's per-sample recurrence (float T_cum re-rounded after every double multiply) in
// lock-step from shuffled factors, so T, the weights and the early-stop index are bit-exact.
// (The fused trainer's march kernels keep a double product scan instead: tolerance class)
#include "common.cuh"

namespace dvgo {

// ---- a8: raw2alpha -------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) raw2alpha_kernel(const float* __restrict__ density,
                                                        float shift, float interval, int64_t n,
                                                        float* __restrict__ exp_d,
                                                        float* __restrict__ alpha) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float e = expf(fadd(density[i], shift));  // :366, may be inf
    exp_d[i] = e;
    alpha[i] = fsub(1.f, powf(fadd(1.f, e), -interval));  // :368
  }
}

__global__ void __launch_bounds__(256) raw2alpha_backward_kernel(const float* __restrict__ exp_d,
                                                                 const float* __restrict__ grad_back,
                                                                 float interval, int64_t n,
                                                                 float* __restrict__ grad) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    // :404 -- min(e, 1e10) against a double literal promotes the product chain to double.
    const float e = exp_d[i];
    const double m = fmin(static_cast<double>(e), 1e10);
    const double pw = static_cast<double>(powf(fadd(1.f, e), fsub(-interval, 1.f)));
    grad[i] = static_cast<float>(m * pw * static_cast<double>(interval) *
                                 static_cast<double>(grad_back[i]));
  }
}

// ---- a9: alpha2weight ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256) a2w_init_kernel(int n_rays, float* __restrict__ alphainv_last,
                                                       int64_t* __restrict__ i_start,
                                                       int64_t* __restrict__ i_end) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n_rays) { alphainv_last[r] = 1.f; i_start[r] = 0; i_end[r] = 0; }  // :480-482
}

// Segment bounds from the sorted ray_id (:461-471 and the host index_put at :489).
__global__ void __launch_bounds__(256) a2w_bounds_kernel(const int64_t* __restrict__ ray_id,
                                                         int64_t n_pts,
                                                         int64_t* __restrict__ i_start,
                                                         int64_t* __restrict__ i_end) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n_pts) return;
  const int64_t r = ray_id[i];
  if (i > 0) {
    const int64_t rp = ray_id[i - 1];
    if (r != rp) { i_start[r] = i; i_end[rp] = i; }
  }
  if (i == n_pts - 1) i_end[r] = n_pts;
}

// One warp per ray.  Writes weight/T for EVERY sample of the ray's segment (fills 0 / 1 after the
// early stop, :478-479), the stop index into i_end (:456) and alphainv_last (:457).
//
// The reference re-rounds its running transmittance to float after every sample (float T_cum, :447:
// T_cum = (float)((double)T_cum * (1. - alpha + 1e-10))), so the recurrence is inherently serial and
// a product scan cannot reproduce its roundings.  Here 32 consecutive alphas are loaded coalesced
// and all lanes then replay the 32 dependent updates TOGETHER from shuffled factors: every lane
// holds the same T_cum and lane j keeps the value it saw at its own sample.  Bit-exact with the
// reference (T, weights, alphainv_last and the stop index), coalesced loads / stores, and the
// serial chain is 32 steps per chunk instead of one thread walking the whole ray.
__global__ void __launch_bounds__(256) alpha2weight_kernel(const float* __restrict__ alpha,
                                                           int n_rays, float* __restrict__ weight,
                                                           float* __restrict__ T,
                                                           float* __restrict__ alphainv_last,
                                                           const int64_t* __restrict__ i_start,
                                                           int64_t* __restrict__ i_end) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < n_rays; r += gridDim.x * wpb) {
    const int64_t i_s = i_start[r];
    const int64_t i_e_max = i_end[r];
    if (i_e_max <= i_s) continue;  // no samples: keeps alphainv_last = 1, i_end = i_start (= 0)
    float T_cum = 1.f;             // :447
    int64_t stop = -1;             // index of the sample whose update drove T below 1e-3
    for (int64_t base = i_s; base < i_e_max && stop < 0; base += 32) {
      const int64_t i = base + lane;
      const bool valid = i < i_e_max;
      const float a = valid ? alpha[i] : 0.f;
      // :450  (1. - alpha + 1e-10) in double
      const double f = __dadd_rn(__dsub_rn(1.0, static_cast<double>(a)), 1e-10);
      const int n_valid = static_cast<int>(min(static_cast<int64_t>(32), i_e_max - base));
      float T_mine = 1.f;
      int first = 32;  // lane whose update drove T below 1e-3
      for (int j = 0; j < n_valid; ++j) {
        const double fj = __shfl_sync(0xffffffffu, f, j);
        if (lane == j) T_mine = T_cum;
        T_cum = __double2float_rn(__dmul_rn(static_cast<double>(T_cum), fj));  // :450
        if (static_cast<double>(T_cum) < 1e-3) { first = j; break; }           // :451 (uniform across the warp)
      }
      if (valid) {
        if (lane <= first) {
          T[i] = T_mine;                // :448
          weight[i] = fmul(T_mine, a);  // :449
        } else {
          T[i] = 1.f;                   // fills of :478-479
          weight[i] = 0.f;
        }
      }
      if (first < 32) {
        stop = base + first;
        // every later sample of the segment keeps the fills
        for (int64_t j = base + 32 + lane; j < i_e_max; j += 32) { T[j] = 1.f; weight[j] = 0.f; }
      }
    }
    if (lane == 0) {
      i_end[r] = (stop >= 0) ? stop + 1 : i_e_max;  // :452-456
      alphainv_last[r] = T_cum;                     // :457
    }
  }
}

// ---- a9 backward ---------------------------------------------------------------------------------
// grad[i] = gw[i]*T[i] - back_i / (1 - alpha[i] + 1e-10), back_i = float accumulator started at
// g_last*alphainv_last and advanced by back = fma(gw[j], w[j], back) from the far end (:522-529; the
// FFMA is what the reference's SASS shows).  Same replay scheme as the forward: one warp per ray walks
// [i_start, i_end) from the far end in chunks of 32; bit-exact with the reference.
__global__ void __launch_bounds__(256) alpha2weight_backward_kernel(
    const float* __restrict__ alpha, const float* __restrict__ weight, const float* __restrict__ T,
    const float* __restrict__ alphainv_last, const int64_t* __restrict__ i_start,
    const int64_t* __restrict__ i_end, int n_rays, const float* __restrict__ grad_weights,
    const float* __restrict__ grad_last, float* __restrict__ grad) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < n_rays; r += gridDim.x * wpb) {
    const int64_t i_s = i_start[r];
    const int64_t i_e = i_end[r];
    if (i_e <= i_s) continue;
    float back = fmul(grad_last[r], alphainv_last[r]);  // :522
    for (int64_t hi = i_e; hi > i_s; hi -= 32) {
      const int64_t i = hi - 1 - lane;  // lane 0 = the farthest sample of the chunk
      const bool valid = i >= i_s;
      const float gw = valid ? grad_weights[i] : 0.f;
      const float w = valid ? weight[i] : 0.f;
      const int n_valid = static_cast<int>(min(static_cast<int64_t>(32), hi - i_s));
      float back_mine = 0.f;
      for (int j = 0; j < n_valid; ++j) {
        const float gwj = __shfl_sync(0xffffffffu, gw, j);
        const float wj = __shfl_sync(0xffffffffu, w, j);
        if (lane == j) back_mine = back;
        back = fma_(gwj, wj, back);  // :525
      }
      if (valid) {
        const float gwT = fmul(gw, T[i]);
        const float one_m_a = fsub(1.f, alpha[i]);
        grad[i] = __double2float_rn(__dsub_rn(static_cast<double>(gwT),
                                              __ddiv_rn(static_cast<double>(back_mine),
                                                        __dadd_rn(static_cast<double>(one_m_a), 1e-10))));  // :524
      }
    }
  }
}

static inline int grid_for(int64_t n, int threads) {
  const int64_t want = (n + threads - 1) / threads;
  const int64_t cap = static_cast<int64_t>(kNumSMs) * 32;
  return static_cast<int>(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace dvgo

using namespace dvgo;

DVGO_API int dvgo_raw2alpha(const float* density, float shift, float interval, int64_t n_pts,
                            float* exp_d, float* alpha, dvgo_stream_t stream) {
  if (n_pts < 0) return DVGO_EINVAL;
  if (n_pts == 0) return 0;  // :377-379
  if (!density || !exp_d || !alpha) return DVGO_EINVAL;
  raw2alpha_kernel<<<grid_for(n_pts, 256), 256, 0, as_stream(stream)>>>(density, shift, interval,
                                                                        n_pts, exp_d, alpha);
  return launch_status();
}

DVGO_API int dvgo_raw2alpha_backward(const float* exp_d, const float* grad_back, float interval,
                                     int64_t n_pts, float* grad, dvgo_stream_t stream) {
  if (n_pts < 0) return DVGO_EINVAL;
  if (n_pts == 0) return 0;
  if (!exp_d || !grad_back || !grad) return DVGO_EINVAL;
  raw2alpha_backward_kernel<<<grid_for(n_pts, 256), 256, 0, as_stream(stream)>>>(
      exp_d, grad_back, interval, n_pts, grad);
  return launch_status();
}

DVGO_API int dvgo_alpha2weight(const float* alpha, const int64_t* ray_id, int n_rays, int64_t n_pts,
                               float* weight, float* T, float* alphainv_last, int64_t* i_start,
                               int64_t* i_end, dvgo_stream_t stream) {
  if (n_rays < 0 || n_pts < 0) return DVGO_EINVAL;
  if (n_rays == 0) return 0;
  if (!alphainv_last || !i_start || !i_end) return DVGO_EINVAL;
  cudaStream_t s = as_stream(stream);
  a2w_init_kernel<<<blocks_for(n_rays, 256), 256, 0, s>>>(n_rays, alphainv_last, i_start, i_end);
  if (n_pts == 0) return launch_status(1);  // :483-485
  if (!alpha || !ray_id || !weight || !T) return DVGO_EINVAL;
  a2w_bounds_kernel<<<blocks_for(n_pts, 256), 256, 0, s>>>(ray_id, n_pts, i_start, i_end);
  const int wpb = 8;
  const int64_t want = (static_cast<int64_t>(n_rays) + wpb - 1) / wpb;
  const int blocks = static_cast<int>(want < kNumSMs * 16 ? want : kNumSMs * 16);
  alpha2weight_kernel<<<blocks, wpb * 32, 0, s>>>(alpha, n_rays, weight, T, alphainv_last, i_start,
                                                  i_end);
  return launch_status(3);
}

DVGO_API int dvgo_alpha2weight_backward(const float* alpha, const float* weight, const float* T,
                                        const float* alphainv_last, const int64_t* i_start,
                                        const int64_t* i_end, int n_rays, int64_t n_pts,
                                        const float* grad_weights, const float* grad_last,
                                        float* grad, dvgo_stream_t stream) {
  if (n_rays < 0 || n_pts < 0) return DVGO_EINVAL;
  if (n_pts == 0) return 0;
  if (!grad) return DVGO_EINVAL;
  cudaStream_t s = as_stream(stream);
  int err = static_cast<int>(cudaMemsetAsync(grad, 0, sizeof(float) * n_pts, s));  // :538
  if (err) return err;
  if (n_rays == 0) return 0;  // :539-541
  if (!alpha || !weight || !T || !alphainv_last || !i_start || !i_end || !grad_weights || !grad_last)
    return DVGO_EINVAL;
  const int wpb = 8;
  const int64_t want = (static_cast<int64_t>(n_rays) + wpb - 1) / wpb;
  const int blocks = static_cast<int>(want < kNumSMs * 16 ? want : kNumSMs * 16);
  alpha2weight_backward_kernel<<<blocks, wpb * 32, 0, s>>>(alpha, weight, T, alphainv_last, i_start,
                                                           i_end, n_rays, grad_weights, grad_last,
                                                           grad);
  return launch_status();
}
