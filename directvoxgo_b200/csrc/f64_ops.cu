// f64_ops.cu -- the float64 instantiations behind include/dvgo_b200_f64.h: what the reference's
// AT_DISPATCH_FLOATING_TYPES sites run when the tensors are double (lib/cuda/render_utils_kernel.cu,
// lib/cuda/total_variation_kernel.cu, lib/cuda/adam_upd_kernel.cu).
//
// Not on the training hot path (the reference's models are fp32); used by gradcheck and by callers
// that hold rays in double.  Same launch structure as the float32 files (2 launches for
// sample_pts_on_rays, a warp per ray for the per-ray recurrences, grid-stride sweeps), but the
// arithmetic is NOT "everything in double": the reference's templates keep many temporaries in
// `float` whatever scalar_t is, and a double tensor then carries float-precision values.  Every
// such rounding is reproduced here (f32() marks them) so the outputs equal the reference's double
// kernels.  Where nvcc's default -fmad=true had a choice, the contraction follows the SASS of the
// reference's double kernels compiled for sm_100a from its unmodified sources:
//   infer_ray_start_dir  DMUL(dy,dy); DFMA(dx,dx,.); DFMA(dz,dz,.); start = DFMA(d, t_min, o)
//   sample_pts / ndc     p = F2F.F32(DFMA(dir, dist, start))
//   maskcache            DFMA(x, scale, shift) then round-half-away, int conversion
//   alpha2weight         T*alpha DMUL; T *= (1 - a) + 1e-10 DADD,DADD,DMUL; bwd: DMUL(gw,T) - quotient, DFMA(gw,w,back)
//   total_variation      DMUL(w, clamp) then DADD into the float accumulator
//   adam                 m = DFMA(b1, m, (1-b1)*g); v = DFMA(b2, v, g*(g*(1-b2))); p -= (m*step) / (sqrt(v)+eps)
// Intrinsics (__dmul_rn, __fma_rn, ...) are never re-contracted by the compiler.
#include "common.cuh"
#include "scan.cuh"

#include "../../include/dvgo_b200_f64.h"

namespace dvgo {
namespace f64 {

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dfma(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }
// A value the reference holds in a `float` variable inside its double instantiation.
__device__ __forceinline__ float f32(double v) { return __double2float_rn(v); }

// ---- a1: ray / AABB clip --------------------------------------------------------------------------
// :23-33 with scalar_t = double: the operands are double, every named temporary is float.
struct Clip { float t_min, t_max; };
__device__ __forceinline__ Clip clip_ray(const double* __restrict__ o, const double* __restrict__ d,
                                         const double* __restrict__ lo, const double* __restrict__ hi,
                                         float near, float far) {
  float a[3], b[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = f32(d[c] == 0.0 ? 1e-6 : d[c]);                           // :23-25
    a[c] = f32(ddiv(dsub(hi[c], o[c]), static_cast<double>(v)));              // :26-28
    b[c] = f32(ddiv(dsub(lo[c], o[c]), static_cast<double>(v)));              // :29-31
  }
  Clip r;
  r.t_min = fmaxf(fminf(fmaxf(fmaxf(fminf(a[0], b[0]), fminf(a[1], b[1])), fminf(a[2], b[2])), far), near);
  r.t_max = fmaxf(fminf(fminf(fminf(fmaxf(a[0], b[0]), fmaxf(a[1], b[1])), fmaxf(a[2], b[2])), far), near);
  return r;
}

// :47  max(ceil((t_max - t_min) / stepdist), 1.) -- subtraction, division and ceil in double.
__device__ __forceinline__ int64_t count_steps(double t_min, double t_max, float stepdist) {
  const double c = ceil(ddiv(dsub(t_max, t_min), static_cast<double>(stepdist)));
  return static_cast<int64_t>(fmax(c, 1.));
}

// :62-71
struct StartDir64 { double s[3], u[3]; };
__device__ __forceinline__ StartDir64 start_dir(const double* __restrict__ o,
                                                const double* __restrict__ d, double t_min) {
  const float rnorm = f32(sqrt(dfma(d[2], d[2], dfma(d[0], d[0], dmul(d[1], d[1])))));  // float rnorm, :62
  StartDir64 r;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    r.s[c] = dfma(d[c], t_min, o[c]);
    r.u[c] = ddiv(d[c], static_cast<double>(rnorm));
  }
  return r;
}

__device__ __forceinline__ bool outside(float px, float py, float pz, const double* __restrict__ lo,
                                        const double* __restrict__ hi) {
  const double x = px, y = py, z = pz;  // :185-186 compares the double bounds with the float point
  return (lo[0] > x) | (lo[1] > y) | (lo[2] > z) | (hi[0] < x) | (hi[1] < y) | (hi[2] < z);
}

__global__ void __launch_bounds__(256) t_minmax_kernel(
    const double* __restrict__ rays_o, const double* __restrict__ rays_d,
    const double* __restrict__ xyz_min, const double* __restrict__ xyz_max, float near, float far,
    float stepdist, int n_rays, double* __restrict__ t_min, double* __restrict__ t_max,
    int64_t* __restrict__ N_steps) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rays) return;
  const Clip t = clip_ray(rays_o + 3 * r, rays_d + 3 * r, xyz_min, xyz_max, near, far);
  t_min[r] = t.t_min;
  t_max[r] = t.t_max;
  if (N_steps) N_steps[r] = count_steps(t.t_min, t.t_max, stepdist);
}

__global__ void __launch_bounds__(256) n_samples_kernel(const double* __restrict__ t_min,
                                                        const double* __restrict__ t_max,
                                                        float stepdist, int n_rays,
                                                        int64_t* __restrict__ n_samples) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n_rays) n_samples[r] = count_steps(t_min[r], t_max[r], stepdist);
}

__global__ void __launch_bounds__(256) start_dir_kernel(const double* __restrict__ rays_o,
                                                        const double* __restrict__ rays_d,
                                                        const double* __restrict__ t_min, int n_rays,
                                                        double* __restrict__ rays_start,
                                                        double* __restrict__ rays_dir) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rays) return;
  const StartDir64 s = start_dir(rays_o + 3 * r, rays_d + 3 * r, t_min[r]);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    rays_start[3 * r + c] = s.s[c];
    rays_dir[3 * r + c] = s.u[c];
  }
}

// a4 phase 2: a warp writes the contiguous run of samples of one ray (coalesced stores).
__global__ void __launch_bounds__(256) fill_kernel(
    const double* __restrict__ rays_o, const double* __restrict__ rays_d,
    const double* __restrict__ xyz_min, const double* __restrict__ xyz_max,
    const double* __restrict__ t_min, const int64_t* __restrict__ cumsum, float stepdist, int n_rays,
    double* __restrict__ rays_pts, uint8_t* __restrict__ mask_outbbox, int64_t* __restrict__ ray_id,
    int64_t* __restrict__ step_id) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < n_rays; r += gridDim.x * wpb) {
    const int64_t end = cumsum[r];
    const int64_t begin = r ? cumsum[r - 1] : 0;
    const int n = static_cast<int>(end - begin);
    const StartDir64 s = start_dir(rays_o + 3 * r, rays_d + 3 * r, t_min[r]);
    for (int i = lane; i < n; i += 32) {
      const double dist = static_cast<double>(fmul(stepdist, static_cast<float>(i)));  // float dist, :178
      const float px = f32(dfma(s.u[0], dist, s.s[0]));                                // float px, :179-181
      const float py = f32(dfma(s.u[1], dist, s.s[1]));
      const float pz = f32(dfma(s.u[2], dist, s.s[2]));
      const int64_t idx = begin + i;
      rays_pts[3 * idx] = px;
      rays_pts[3 * idx + 1] = py;
      rays_pts[3 * idx + 2] = pz;
      mask_outbbox[idx] = outside(px, py, pz, xyz_min, xyz_max);
      ray_id[idx] = r;
      step_id[idx] = i;
    }
  }
}

__global__ void __launch_bounds__(256) ndc_kernel(const double* __restrict__ rays_o,
                                                  const double* __restrict__ rays_d,
                                                  const double* __restrict__ xyz_min,
                                                  const double* __restrict__ xyz_max, int N_samples,
                                                  int64_t total, double* __restrict__ rays_pts,
                                                  uint8_t* __restrict__ mask_outbbox) {
  const float denom = static_cast<float>(N_samples - 1);
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(idx / N_samples);
    const int s = static_cast<int>(idx - static_cast<int64_t>(r) * N_samples);
    const double dist = static_cast<double>(fdiv(static_cast<float>(s), denom));  // float dist, :254
    const float px = f32(dfma(rays_d[3 * r], dist, rays_o[3 * r]));
    const float py = f32(dfma(rays_d[3 * r + 1], dist, rays_o[3 * r + 1]));
    const float pz = f32(dfma(rays_d[3 * r + 2], dist, rays_o[3 * r + 2]));
    rays_pts[3 * idx] = px;
    rays_pts[3 * idx + 1] = py;
    rays_pts[3 * idx + 2] = pz;
    mask_outbbox[idx] = outside(px, py, pz, xyz_min, xyz_max);
  }
}

// ---- a6 ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) maskcache_kernel(
    const uint8_t* __restrict__ world, const double* __restrict__ xyz,
    const double* __restrict__ scale, const double* __restrict__ shift, int sz_i, int sz_j, int sz_k,
    int64_t n_pts, uint8_t* __restrict__ out) {
  for (int64_t p = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; p < n_pts;
       p += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    // :312-314: DFMA, round() half away from zero, conversion to int (truncating; the value is integral)
    const int i = static_cast<int>(round(dfma(xyz[3 * p], scale[0], shift[0])));
    const int j = static_cast<int>(round(dfma(xyz[3 * p + 1], scale[1], shift[1])));
    const int k = static_cast<int>(round(dfma(xyz[3 * p + 2], scale[2], shift[2])));
    bool v = false;
    if ((0 <= i) & (i < sz_i) & (0 <= j) & (j < sz_j) & (0 <= k) & (k < sz_k))
      v = world[(static_cast<int64_t>(i) * sz_j + j) * sz_k + k] != 0;
    out[p] = v;
  }
}

// ---- a8 ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) raw2alpha_kernel(const double* __restrict__ density,
                                                        float shift, float interval, int64_t n,
                                                        double* __restrict__ exp_d,
                                                        double* __restrict__ alpha) {
  const double sh = shift;
  const double neg_interval = static_cast<double>(-interval);  // negated in float, then widened (:368)
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const double e = exp(dadd(density[i], sh));  // :366, may be inf
    exp_d[i] = e;
    alpha[i] = dsub(1.0, pow(dadd(1.0, e), neg_interval));  // :368
  }
}

__global__ void __launch_bounds__(256) raw2alpha_backward_kernel(const double* __restrict__ exp_d,
                                                                 const double* __restrict__ grad_back,
                                                                 float interval, int64_t n,
                                                                 double* __restrict__ grad) {
  const double p = static_cast<double>(fsub(-interval, 1.f));  // -interval-1 in float (:404)
  const double iv = interval;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const double e = exp_d[i];
    // :404  min(e, 1e10) * pow(1 + e, -interval - 1) * interval * grad_back, left to right
    grad[i] = dmul(dmul(dmul(fmin(e, 1e10), pow(dadd(1.0, e), p)), iv), grad_back[i]);
  }
}

// ---- a9 ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) a2w_init_kernel(int n_rays, double* __restrict__ alphainv_last,
                                                       int64_t* __restrict__ i_start,
                                                       int64_t* __restrict__ i_end) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n_rays) { alphainv_last[r] = 1.0; i_start[r] = 0; i_end[r] = 0; }  // :480-482
}

__global__ void __launch_bounds__(256) a2w_bounds_kernel(const int64_t* __restrict__ ray_id,
                                                         int64_t n_pts, int64_t* __restrict__ i_start,
                                                         int64_t* __restrict__ i_end) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n_pts) return;
  const int64_t r = ray_id[i];
  if (i > 0) {
    const int64_t rp = ray_id[i - 1];
    if (r != rp) { i_start[r] = i; i_end[rp] = i; }  // :461-471
  }
  if (i == n_pts - 1) i_end[r] = n_pts;              // :489
}

// The reference re-rounds its running transmittance to float after every sample (float T_cum, :447),
// so the recurrence is inherently serial.  One warp per ray: 32 consecutive alphas are loaded
// coalesced, then all lanes replay the 32 dependent updates together from shuffled factors (every
// lane holds the same T_cum; lane j keeps the value it saw at its own sample).  Exact, and the
// loads / stores stay coalesced (the reference's thread-per-ray loop strides by the ray length).
__global__ void __launch_bounds__(256) alpha2weight_kernel(const double* __restrict__ alpha,
                                                           int n_rays, double* __restrict__ weight,
                                                           double* __restrict__ T,
                                                           double* __restrict__ alphainv_last,
                                                           const int64_t* __restrict__ i_start,
                                                           int64_t* __restrict__ i_end) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < n_rays; r += gridDim.x * wpb) {
    const int64_t i_s = i_start[r];
    const int64_t i_e_max = i_end[r];
    if (i_e_max <= i_s) continue;  // keeps alphainv_last = 1, i_end = i_start
    float T_cum = 1.f;             // :447
    int64_t stop = -1;
    for (int64_t base = i_s; base < i_e_max && stop < 0; base += 32) {
      const int64_t i = base + lane;
      const bool valid = i < i_e_max;
      const double a = valid ? alpha[i] : 0.0;
      const double f = dadd(dsub(1.0, a), 1e-10);  // (1. - alpha + 1e-10), :450
      const int n_valid = static_cast<int>(min(static_cast<int64_t>(32), i_e_max - base));
      float T_mine = 1.f;
      int first = 32;  // lane whose update drove T below 1e-3
      for (int j = 0; j < n_valid; ++j) {
        const double fj = __shfl_sync(0xffffffffu, f, j);
        if (lane == j) T_mine = T_cum;
        T_cum = f32(dmul(static_cast<double>(T_cum), fj));           // :450
        if (static_cast<double>(T_cum) < 1e-3) { first = j; break; }  // :451 (uniform across the warp)
      }
      if (valid) {
        if (lane <= first) {
          T[i] = T_mine;                                     // :448
          weight[i] = dmul(static_cast<double>(T_mine), a);  // :449
        } else {
          T[i] = 1.0;  // fills of :478-479
          weight[i] = 0.0;
        }
      }
      if (first < 32) {
        stop = base + first;
        for (int64_t j = base + 32 + lane; j < i_e_max; j += 32) { T[j] = 1.0; weight[j] = 0.0; }
      }
    }
    if (lane == 0) {
      i_end[r] = (stop >= 0) ? stop + 1 : i_e_max;  // :452-456
      alphainv_last[r] = T_cum;                     // :457
    }
  }
}

// Backward: float back_cum (:522), walked from the far end; same replay scheme.
__global__ void __launch_bounds__(256) alpha2weight_backward_kernel(
    const double* __restrict__ alpha, const double* __restrict__ weight, const double* __restrict__ T,
    const double* __restrict__ alphainv_last, const int64_t* __restrict__ i_start,
    const int64_t* __restrict__ i_end, int n_rays, const double* __restrict__ grad_weights,
    const double* __restrict__ grad_last, double* __restrict__ grad) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < n_rays; r += gridDim.x * wpb) {
    const int64_t i_s = i_start[r];
    const int64_t i_e = i_end[r];
    if (i_e <= i_s) continue;
    float back = f32(dmul(grad_last[r], alphainv_last[r]));  // :522
    for (int64_t hi = i_e; hi > i_s; hi -= 32) {
      const int64_t i = hi - 1 - lane;  // lane 0 = farthest sample of the chunk
      const bool valid = i >= i_s;
      const double gw = valid ? grad_weights[i] : 0.0;
      const double w = valid ? weight[i] : 0.0;
      const int n_valid = static_cast<int>(min(static_cast<int64_t>(32), hi - i_s));
      float back_mine = 0.f;
      for (int j = 0; j < n_valid; ++j) {
        const double gwj = __shfl_sync(0xffffffffu, gw, j);
        const double wj = __shfl_sync(0xffffffffu, w, j);
        if (lane == j) back_mine = back;
        back = f32(dfma(gwj, wj, static_cast<double>(back)));  // :525
      }
      if (valid)  // :524
        grad[i] = dsub(dmul(gw, T[i]),
                       ddiv(static_cast<double>(back_mine), dadd(dsub(1.0, alpha[i]), 1e-10)));
    }
  }
}

// ---- a12 -----------------------------------------------------------------------------------------
__device__ __forceinline__ double clamp1(double v) { return fmin(fmax(v, -1.0), 1.0); }

template <bool kDense>
__global__ void __launch_bounds__(256) tv_kernel(const double* __restrict__ param,
                                                 double* __restrict__ grad, float wy, float wz,
                                                 int64_t sz_i, int64_t sz_j, int64_t sz_k, int64_t N) {
  const int64_t sjk = sz_j * sz_k;
  const double dwy = wy, dwz = wz;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < N;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const double g0 = grad[idx];
    if (!kDense && g0 == 0.0) continue;  // :21
    const int64_t k = idx % sz_k;
    const int64_t j = idx / sz_k % sz_j;
    const int64_t i = idx / sjk % sz_i;
    const double p = param[idx];
    float acc = 0.f;  // float grad_to_add, :25; each += rounds (double)acc + w*clamp back to float
    if (k != 0) acc = f32(dadd(acc, dmul(dwz, clamp1(dsub(p, param[idx - 1])))));
    if (k != sz_k - 1) acc = f32(dadd(acc, dmul(dwz, clamp1(dsub(p, param[idx + 1])))));
    if (j != 0) acc = f32(dadd(acc, dmul(dwy, clamp1(dsub(p, param[idx - sz_k])))));
    if (j != sz_j - 1) acc = f32(dadd(acc, dmul(dwy, clamp1(dsub(p, param[idx + sz_k])))));
    if (i != 0) acc = f32(dadd(acc, dmul(dwz, clamp1(dsub(p, param[idx - sjk])))));      // i-axis uses wz, :31-32
    if (i != sz_i - 1) acc = f32(dadd(acc, dmul(dwz, clamp1(dsub(p, param[idx + sjk])))));
    grad[idx] = dadd(g0, static_cast<double>(acc));  // :33
  }
}

// ---- a13 -----------------------------------------------------------------------------------------
// mode 0: adam_upd (:8-23), 1: masked_adam_upd (:25-40), 2: adam_upd_with_perlr (:42-58).
template <int kMode>
__global__ void __launch_bounds__(256) adam_kernel(double* __restrict__ param,
                                                   const double* __restrict__ grad,
                                                   double* __restrict__ exp_avg,
                                                   double* __restrict__ exp_avg_sq,
                                                   const double* __restrict__ perlr, int64_t N,
                                                   float step_size, float beta1, float beta2,
                                                   float eps) {
  const double b1 = beta1, b2 = beta2, ss = step_size, de = eps;
  const double omb1 = static_cast<double>(fsub(1.f, beta1));  // (1 - beta) in float, then widened
  const double omb2 = static_cast<double>(fsub(1.f, beta2));
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < N;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const double g = grad[i];
    if (kMode == 1 && g == 0.0) continue;  // :35
    const double m = dfma(b1, exp_avg[i], dmul(omb1, g));
    const double v = dfma(b2, exp_avg_sq[i], dmul(g, dmul(g, omb2)));
    exp_avg[i] = m;
    exp_avg_sq[i] = v;
    const double num = (kMode == 2) ? dmul(dmul(ss, perlr[i]), m) : dmul(m, ss);
    param[i] = dsub(param[i], ddiv(num, dadd(sqrt(v), de)));
  }
}

static inline int grid_for(int64_t n, int threads) {
  const int64_t want = (n + threads - 1) / threads;
  const int64_t cap = static_cast<int64_t>(kNumSMs) * 16;
  return static_cast<int>(want < cap ? (want > 0 ? want : 1) : cap);
}

static inline int ray_grid(int n_rays, int wpb) {
  const int64_t want = (static_cast<int64_t>(n_rays) + wpb - 1) / wpb;
  return static_cast<int>(want < kNumSMs * 16 ? want : kNumSMs * 16);
}

// Host-side bias correction, adam_upd_kernel.cu:72: a float variable whatever the tensors are.
static inline float adam_step_size(int step, float beta1, float beta2, float lr) {
  return lr * sqrtf(1.f - powf(beta2, static_cast<float>(step))) /
         (1.f - powf(beta1, static_cast<float>(step)));
}

template <int kMode>
static int adam_launch(double* param, const double* grad, double* exp_avg, double* exp_avg_sq,
                       const double* perlr, int64_t N, int step, float beta1, float beta2, float lr,
                       float eps, dvgo_stream_t stream) {
  if (N < 0) return DVGO_EINVAL;
  if (N == 0) return 0;
  if (!param || !grad || !exp_avg || !exp_avg_sq || (kMode == 2 && !perlr)) return DVGO_EINVAL;
  adam_kernel<kMode><<<grid_for(N, 256), 256, 0, as_stream(stream)>>>(
      param, grad, exp_avg, exp_avg_sq, perlr, N, adam_step_size(step, beta1, beta2, lr), beta1,
      beta2, eps);
  return launch_status();
}

}  // namespace f64
}  // namespace dvgo

using namespace dvgo;

DVGO_API int dvgo_infer_t_minmax_f64(const double* rays_o, const double* rays_d,
                                     const double* xyz_min, const double* xyz_max, float near,
                                     float far, int n_rays, double* t_min, double* t_max,
                                     dvgo_stream_t stream) {
  if (n_rays < 0) return DVGO_EINVAL;
  if (n_rays == 0) return 0;
  if (!rays_o || !rays_d || !xyz_min || !xyz_max || !t_min || !t_max) return DVGO_EINVAL;
  f64::t_minmax_kernel<<<blocks_for(n_rays, 256), 256, 0, as_stream(stream)>>>(
      rays_o, rays_d, xyz_min, xyz_max, near, far, 0.f, n_rays, t_min, t_max, nullptr);
  return launch_status();
}

DVGO_API int dvgo_infer_n_samples_f64(const double* t_min, const double* t_max, float stepdist,
                                      int n_rays, int64_t* n_samples, dvgo_stream_t stream) {
  if (n_rays < 0) return DVGO_EINVAL;
  if (n_rays == 0) return 0;
  if (!t_min || !t_max || !n_samples) return DVGO_EINVAL;
  f64::n_samples_kernel<<<blocks_for(n_rays, 256), 256, 0, as_stream(stream)>>>(t_min, t_max, stepdist,
                                                                               n_rays, n_samples);
  return launch_status();
}

DVGO_API int dvgo_infer_ray_start_dir_f64(const double* rays_o, const double* rays_d,
                                          const double* t_min, int n_rays, double* rays_start,
                                          double* rays_dir, dvgo_stream_t stream) {
  if (n_rays < 0) return DVGO_EINVAL;
  if (n_rays == 0) return 0;
  if (!rays_o || !rays_d || !t_min || !rays_start || !rays_dir) return DVGO_EINVAL;
  f64::start_dir_kernel<<<blocks_for(n_rays, 256), 256, 0, as_stream(stream)>>>(
      rays_o, rays_d, t_min, n_rays, rays_start, rays_dir);
  return launch_status();
}

DVGO_API int dvgo_sample_pts_count_f64(const double* rays_o, const double* rays_d,
                                       const double* xyz_min, const double* xyz_max, float near,
                                       float far, float stepdist, int n_rays, double* t_min,
                                       double* t_max, int64_t* N_steps, int64_t* N_steps_cumsum,
                                       int64_t* total_host, dvgo_stream_t stream) {
  if (n_rays < 0 || !total_host) return DVGO_EINVAL;
  *total_host = 0;
  if (n_rays == 0) return 0;
  if (!rays_o || !rays_d || !xyz_min || !xyz_max || !t_min || !t_max || !N_steps || !N_steps_cumsum)
    return DVGO_EINVAL;
  cudaStream_t s = as_stream(stream);
  f64::t_minmax_kernel<<<blocks_for(n_rays, 256), 256, 0, s>>>(rays_o, rays_d, xyz_min, xyz_max, near,
                                                               far, stepdist, n_rays, t_min, t_max,
                                                               N_steps);
  inclusive_scan_i64_kernel<<<1, kScanThreads, 0, s>>>(N_steps, n_rays, N_steps_cumsum);
  int err = launch_status(2);
  if (err) return err;
  // the reference's one host sync (render_utils_kernel.cu:206)
  err = static_cast<int>(cudaMemcpyAsync(total_host, N_steps_cumsum + (n_rays - 1), sizeof(int64_t),
                                         cudaMemcpyDeviceToHost, s));
  if (err) return err;
  return static_cast<int>(cudaStreamSynchronize(s));
}

DVGO_API int dvgo_sample_pts_fill_f64(const double* rays_o, const double* rays_d,
                                      const double* xyz_min, const double* xyz_max,
                                      const double* t_min, const int64_t* N_steps_cumsum,
                                      float stepdist, int n_rays, int64_t total, double* rays_pts,
                                      uint8_t* mask_outbbox, int64_t* ray_id, int64_t* step_id,
                                      dvgo_stream_t stream) {
  if (n_rays < 0 || total < 0 || total >= (int64_t(1) << 31)) return DVGO_EINVAL;
  if (n_rays == 0 || total == 0) return 0;
  if (!rays_o || !rays_d || !xyz_min || !xyz_max || !t_min || !N_steps_cumsum || !rays_pts ||
      !mask_outbbox || !ray_id || !step_id)
    return DVGO_EINVAL;
  f64::fill_kernel<<<f64::ray_grid(n_rays, 8), 256, 0, as_stream(stream)>>>(
      rays_o, rays_d, xyz_min, xyz_max, t_min, N_steps_cumsum, stepdist, n_rays, rays_pts,
      mask_outbbox, ray_id, step_id);
  return launch_status();
}

DVGO_API int dvgo_sample_ndc_pts_on_rays_f64(const double* rays_o, const double* rays_d,
                                             const double* xyz_min, const double* xyz_max,
                                             int N_samples, int n_rays, double* rays_pts,
                                             uint8_t* mask_outbbox, dvgo_stream_t stream) {
  if (n_rays < 0 || N_samples < 0) return DVGO_EINVAL;
  const int64_t total = static_cast<int64_t>(n_rays) * N_samples;
  if (total == 0) return 0;
  if (!rays_o || !rays_d || !xyz_min || !xyz_max || !rays_pts || !mask_outbbox) return DVGO_EINVAL;
  f64::ndc_kernel<<<f64::grid_for(total, 256), 256, 0, as_stream(stream)>>>(
      rays_o, rays_d, xyz_min, xyz_max, N_samples, total, rays_pts, mask_outbbox);
  return launch_status();
}

DVGO_API int dvgo_maskcache_lookup_f64(const uint8_t* world, const double* xyz,
                                       const double* xyz2ijk_scale, const double* xyz2ijk_shift,
                                       int sz_i, int sz_j, int sz_k, int64_t n_pts, uint8_t* out,
                                       dvgo_stream_t stream) {
  if (n_pts < 0 || sz_i < 0 || sz_j < 0 || sz_k < 0) return DVGO_EINVAL;
  if (n_pts == 0) return 0;  // :333-335
  if (!world || !xyz || !xyz2ijk_scale || !xyz2ijk_shift || !out) return DVGO_EINVAL;
  f64::maskcache_kernel<<<f64::grid_for(n_pts, 256), 256, 0, as_stream(stream)>>>(
      world, xyz, xyz2ijk_scale, xyz2ijk_shift, sz_i, sz_j, sz_k, n_pts, out);
  return launch_status();
}

DVGO_API int dvgo_raw2alpha_f64(const double* density, float shift, float interval, int64_t n_pts,
                                double* exp_d, double* alpha, dvgo_stream_t stream) {
  if (n_pts < 0) return DVGO_EINVAL;
  if (n_pts == 0) return 0;  // :377-379
  if (!density || !exp_d || !alpha) return DVGO_EINVAL;
  f64::raw2alpha_kernel<<<f64::grid_for(n_pts, 256), 256, 0, as_stream(stream)>>>(
      density, shift, interval, n_pts, exp_d, alpha);
  return launch_status();
}

DVGO_API int dvgo_raw2alpha_backward_f64(const double* exp_d, const double* grad_back, float interval,
                                         int64_t n_pts, double* grad, dvgo_stream_t stream) {
  if (n_pts < 0) return DVGO_EINVAL;
  if (n_pts == 0) return 0;
  if (!exp_d || !grad_back || !grad) return DVGO_EINVAL;
  f64::raw2alpha_backward_kernel<<<f64::grid_for(n_pts, 256), 256, 0, as_stream(stream)>>>(
      exp_d, grad_back, interval, n_pts, grad);
  return launch_status();
}

DVGO_API int dvgo_alpha2weight_f64(const double* alpha, const int64_t* ray_id, int n_rays,
                                   int64_t n_pts, double* weight, double* T, double* alphainv_last,
                                   int64_t* i_start, int64_t* i_end, dvgo_stream_t stream) {
  if (n_rays < 0 || n_pts < 0) return DVGO_EINVAL;
  if (n_rays == 0) return 0;
  if (!alphainv_last || !i_start || !i_end) return DVGO_EINVAL;
  cudaStream_t s = as_stream(stream);
  f64::a2w_init_kernel<<<blocks_for(n_rays, 256), 256, 0, s>>>(n_rays, alphainv_last, i_start, i_end);
  if (n_pts == 0) return launch_status(1);  // :483-485
  if (!alpha || !ray_id || !weight || !T) return DVGO_EINVAL;
  f64::a2w_bounds_kernel<<<blocks_for(n_pts, 256), 256, 0, s>>>(ray_id, n_pts, i_start, i_end);
  f64::alpha2weight_kernel<<<f64::ray_grid(n_rays, 8), 256, 0, s>>>(alpha, n_rays, weight, T,
                                                                    alphainv_last, i_start, i_end);
  return launch_status(3);
}

DVGO_API int dvgo_alpha2weight_backward_f64(const double* alpha, const double* weight,
                                            const double* T, const double* alphainv_last,
                                            const int64_t* i_start, const int64_t* i_end, int n_rays,
                                            int64_t n_pts, const double* grad_weights,
                                            const double* grad_last, double* grad,
                                            dvgo_stream_t stream) {
  if (n_rays < 0 || n_pts < 0) return DVGO_EINVAL;
  if (n_pts == 0) return 0;
  if (!grad) return DVGO_EINVAL;
  cudaStream_t s = as_stream(stream);
  int err = static_cast<int>(cudaMemsetAsync(grad, 0, sizeof(double) * n_pts, s));  // :538
  if (err) return err;
  if (n_rays == 0) return 0;  // :539-541
  if (!alpha || !weight || !T || !alphainv_last || !i_start || !i_end || !grad_weights || !grad_last)
    return DVGO_EINVAL;
  f64::alpha2weight_backward_kernel<<<f64::ray_grid(n_rays, 8), 256, 0, s>>>(
      alpha, weight, T, alphainv_last, i_start, i_end, n_rays, grad_weights, grad_last, grad);
  return launch_status();
}

DVGO_API int dvgo_total_variation_add_grad_f64(const double* param, double* grad, float wx, float wy,
                                               float wz, int dense_mode, int64_t N, int64_t sz_i,
                                               int64_t sz_j, int64_t sz_k, dvgo_stream_t stream) {
  (void)wx;  // unused by the reference as well (total_variation_kernel.cu:31-32)
  if (N < 0 || sz_i <= 0 || sz_j <= 0 || sz_k <= 0) return DVGO_EINVAL;
  if (N == 0) return 0;
  if (!param || !grad) return DVGO_EINVAL;
  wy /= 6;  // float divisions on the host, :45-47
  wz /= 6;
  const int blocks = f64::grid_for(N, 256);
  if (dense_mode)
    f64::tv_kernel<true><<<blocks, 256, 0, as_stream(stream)>>>(param, grad, wy, wz, sz_i, sz_j, sz_k, N);
  else
    f64::tv_kernel<false><<<blocks, 256, 0, as_stream(stream)>>>(param, grad, wy, wz, sz_i, sz_j, sz_k, N);
  return launch_status();
}

DVGO_API int dvgo_adam_upd_f64(double* param, const double* grad, double* exp_avg, double* exp_avg_sq,
                               int64_t N, int step, float beta1, float beta2, float lr, float eps,
                               dvgo_stream_t stream) {
  return f64::adam_launch<0>(param, grad, exp_avg, exp_avg_sq, nullptr, N, step, beta1, beta2, lr, eps,
                             stream);
}

DVGO_API int dvgo_masked_adam_upd_f64(double* param, const double* grad, double* exp_avg,
                                      double* exp_avg_sq, int64_t N, int step, float beta1,
                                      float beta2, float lr, float eps, dvgo_stream_t stream) {
  return f64::adam_launch<1>(param, grad, exp_avg, exp_avg_sq, nullptr, N, step, beta1, beta2, lr, eps,
                             stream);
}

DVGO_API int dvgo_adam_upd_with_perlr_f64(double* param, const double* grad, double* exp_avg,
                                          double* exp_avg_sq, const double* perlr, int64_t N,
                                          int step, float beta1, float beta2, float lr, float eps,
                                          dvgo_stream_t stream) {
  return f64::adam_launch<2>(param, grad, exp_avg, exp_avg_sq, perlr, N, step, beta1, beta2, lr, eps,
                             stream);
}
