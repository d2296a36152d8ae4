/*
 * dvgo_oracle_f64.c -- TEST INFRASTRUCTURE ONLY (compiled into oracle/libdvgo_oracle.so).
 *
 * Plain-C, single-threaded CPU restatement of the DOUBLE instantiation of the reference's kernels
 * (hbell99/DirectVoxGO, lib/cuda/render_utils_kernel.cu, total_variation_kernel.cu,
 * adam_upd_kernel.cu under AT_DISPATCH_FLOATING_TYPES with scalar_t = double).  It checks
 * directvoxgo_b200/csrc/f64_ops.cu; nothing in the product imports, links or executes it.
 *
 * The reference's templates keep many temporaries in `float` whatever scalar_t is; those roundings
 * are written out here with (float) casts, following the C++ usual arithmetic conversions of each
 * cited line.  Where nvcc (-fmad=true) contracts a*b+c into one DFMA -- read off the SASS of
 * oracle/_ref's double kernels -- fma() is written; gcc is run with -ffp-contract=off so that it
 * fuses nothing on its own.
 *
 * Pinning: tests/golden/ref_gpu_ops_f64.npz = outputs of the reference's own double kernels
 * (oracle/_ref) recorded on a B200 by `python -m oracle.make_golden_gpu <out> f64`.  exp()/pow()
 * come from glibc here and from CUDA's libdevice there (both < 1 ulp, not bit-identical), so
 * raw2alpha is pinned to a few double ulps; everything else is pinned bit for bit.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define EXPORT __attribute__((visibility("default")))

static inline float fminf_(float a, float b) { return a < b ? a : b; }
static inline float fmaxf_(float a, float b) { return a > b ? a : b; }

/* render_utils_kernel.cu:12-35 */
EXPORT void orc64_infer_t_minmax(const double* rays_o, const double* rays_d, const double* xyz_min,
                                 const double* xyz_max, float near, float far, int n_rays,
                                 double* t_min, double* t_max) {
  for (int r = 0; r < n_rays; ++r) {
    const double* o = rays_o + 3 * r;
    const double* d = rays_d + 3 * r;
    float a[3], b[3];
    for (int c = 0; c < 3; ++c) {
      float v = (float)((d[c] == 0) ? 1e-6 : d[c]);  /* float vx, :23-25 */
      a[c] = (float)((xyz_max[c] - o[c]) / (double)v); /* float ax, :26-28 */
      b[c] = (float)((xyz_min[c] - o[c]) / (double)v);
    }
    float lo = fmaxf_(fmaxf_(fminf_(a[0], b[0]), fminf_(a[1], b[1])), fminf_(a[2], b[2]));
    float hi = fminf_(fminf_(fmaxf_(a[0], b[0]), fmaxf_(a[1], b[1])), fmaxf_(a[2], b[2]));
    t_min[r] = fmaxf_(fminf_(lo, far), near);
    t_max[r] = fmaxf_(fminf_(hi, far), near);
  }
}

/* :38-49 */
EXPORT void orc64_infer_n_samples(const double* t_min, const double* t_max, float stepdist, int n_rays,
                                  int64_t* n_samples) {
  for (int r = 0; r < n_rays; ++r) {
    double c = ceil((t_max[r] - t_min[r]) / (double)stepdist);
    n_samples[r] = (int64_t)(c > 1. ? c : 1.);
  }
}

/* :52-73; SASS: DMUL(dy,dy) DFMA(dx,dx,.) DFMA(dz,dz,.), start = DFMA(d, t_min, o) */
EXPORT void orc64_infer_ray_start_dir(const double* rays_o, const double* rays_d, const double* t_min,
                                      int n_rays, double* rays_start, double* rays_dir) {
  for (int r = 0; r < n_rays; ++r) {
    const double* o = rays_o + 3 * r;
    const double* d = rays_d + 3 * r;
    const float rnorm = (float)sqrt(fma(d[2], d[2], fma(d[0], d[0], d[1] * d[1])));
    for (int c = 0; c < 3; ++c) {
      rays_start[3 * r + c] = fma(d[c], t_min[r], o[c]);
      rays_dir[3 * r + c] = d[c] / (double)rnorm;
    }
  }
}

/* :138-236.  Returns the total; N_steps per ray. */
EXPORT int64_t orc64_sample_pts_count(const double* rays_o, const double* rays_d, const double* xyz_min,
                                      const double* xyz_max, float near, float far, float stepdist,
                                      int n_rays, double* t_min, double* t_max, int64_t* N_steps) {
  orc64_infer_t_minmax(rays_o, rays_d, xyz_min, xyz_max, near, far, n_rays, t_min, t_max);
  orc64_infer_n_samples(t_min, t_max, stepdist, n_rays, N_steps);
  int64_t total = 0;
  for (int r = 0; r < n_rays; ++r) total += N_steps[r];
  return total;
}

static inline uint8_t outside64(float px, float py, float pz, const double* lo, const double* hi) {
  return (lo[0] > px) | (lo[1] > py) | (lo[2] > pz) | (hi[0] < px) | (hi[1] < py) | (hi[2] < pz);
}

EXPORT void orc64_sample_pts_fill(const double* rays_o, const double* rays_d, const double* xyz_min,
                                  const double* xyz_max, const double* t_min, const int64_t* N_steps,
                                  float stepdist, int n_rays, double* rays_pts, uint8_t* mask_outbbox,
                                  int64_t* ray_id, int64_t* step_id) {
  int64_t idx = 0;
  for (int r = 0; r < n_rays; ++r) {
    double s[3], u[3];
    orc64_infer_ray_start_dir(rays_o + 3 * r, rays_d + 3 * r, t_min + r, 1, s, u);
    for (int64_t i = 0; i < N_steps[r]; ++i, ++idx) {
      const float dist = stepdist * (float)(int)i;             /* float dist, :178 */
      const float px = (float)fma(u[0], (double)dist, s[0]);  /* float px, :179-181 */
      const float py = (float)fma(u[1], (double)dist, s[1]);
      const float pz = (float)fma(u[2], (double)dist, s[2]);
      rays_pts[3 * idx] = px;
      rays_pts[3 * idx + 1] = py;
      rays_pts[3 * idx + 2] = pz;
      mask_outbbox[idx] = outside64(px, py, pz, xyz_min, xyz_max);
      ray_id[idx] = r;
      step_id[idx] = i;
    }
  }
}

/* :239-287 */
EXPORT void orc64_sample_ndc_pts_on_rays(const double* rays_o, const double* rays_d,
                                         const double* xyz_min, const double* xyz_max, int N_samples,
                                         int n_rays, double* rays_pts, uint8_t* mask_outbbox) {
  for (int r = 0; r < n_rays; ++r)
    for (int s = 0; s < N_samples; ++s) {
      const int64_t idx = (int64_t)r * N_samples + s;
      const float dist = ((float)s) / (float)(N_samples - 1); /* :254 */
      const float px = (float)fma(rays_d[3 * r], (double)dist, rays_o[3 * r]);
      const float py = (float)fma(rays_d[3 * r + 1], (double)dist, rays_o[3 * r + 1]);
      const float pz = (float)fma(rays_d[3 * r + 2], (double)dist, rays_o[3 * r + 2]);
      rays_pts[3 * idx] = px;
      rays_pts[3 * idx + 1] = py;
      rays_pts[3 * idx + 2] = pz;
      mask_outbbox[idx] = outside64(px, py, pz, xyz_min, xyz_max);
    }
}

/* :294-351; SASS: DFMA(x, scale, shift), round half away from zero */
EXPORT void orc64_maskcache_lookup(const uint8_t* world, const double* xyz, const double* scale,
                                   const double* shift, int sz_i, int sz_j, int sz_k, int64_t n_pts,
                                   uint8_t* out) {
  for (int64_t p = 0; p < n_pts; ++p) {
    const double fi = round(fma(xyz[3 * p], scale[0], shift[0]));
    const double fj = round(fma(xyz[3 * p + 1], scale[1], shift[1]));
    const double fk = round(fma(xyz[3 * p + 2], scale[2], shift[2]));
    uint8_t v = 0;
    if (fi >= 0 && fi < sz_i && fj >= 0 && fj < sz_j && fk >= 0 && fk < sz_k)
      v = world[((int64_t)fi * sz_j + (int64_t)fj) * sz_k + (int64_t)fk] != 0;
    out[p] = v;
  }
}

/* :358-428 */
EXPORT void orc64_raw2alpha(const double* density, float shift, float interval, int64_t n,
                            double* exp_d, double* alpha) {
  const double neg_interval = (double)(-interval);
  for (int64_t i = 0; i < n; ++i) {
    const double e = exp(density[i] + (double)shift);
    exp_d[i] = e;
    alpha[i] = 1 - pow(1 + e, neg_interval);
  }
}

EXPORT void orc64_raw2alpha_backward(const double* exp_d, const double* grad_back, float interval,
                                     int64_t n, double* grad) {
  const double p = (double)(-interval - 1.f); /* float arithmetic, then widened, :404 */
  for (int64_t i = 0; i < n; ++i) {
    const double e = exp_d[i];
    grad[i] = (e < 1e10 ? e : 1e10) * pow(1 + e, p) * (double)interval * grad_back[i];
  }
}

/* :431-505 */
EXPORT void orc64_alpha2weight(const double* alpha, const int64_t* ray_id, int n_rays, int64_t n_pts,
                               double* weight, double* T, double* alphainv_last, int64_t* i_start,
                               int64_t* i_end) {
  for (int64_t i = 0; i < n_pts; ++i) { weight[i] = 0; T[i] = 1; }
  for (int r = 0; r < n_rays; ++r) { alphainv_last[r] = 1; i_start[r] = 0; i_end[r] = 0; }
  if (n_pts == 0) return;
  for (int64_t i = 1; i < n_pts; ++i)
    if (ray_id[i] != ray_id[i - 1]) { i_start[ray_id[i]] = i; i_end[ray_id[i - 1]] = i; }
  i_end[ray_id[n_pts - 1]] = n_pts;
  for (int r = 0; r < n_rays; ++r) {
    float T_cum = 1.f; /* float whatever scalar_t is, :447 */
    int64_t i;
    for (i = i_start[r]; i < i_end[r]; ++i) {
      T[i] = T_cum;
      weight[i] = (double)T_cum * alpha[i];
      T_cum = (float)((double)T_cum * (1. - alpha[i] + 1e-10));
      if ((double)T_cum < 1e-3) { i += 1; break; }
    }
    i_end[r] = i;
    alphainv_last[r] = T_cum;
  }
}

/* :507-561; SASS: DMUL(gw,T) - quotient, back = (float)DFMA(gw, w, back) */
EXPORT void orc64_alpha2weight_backward(const double* alpha, const double* weight, const double* T,
                                        const double* alphainv_last, const int64_t* i_start,
                                        const int64_t* i_end, int n_rays, int64_t n_pts,
                                        const double* grad_weights, const double* grad_last,
                                        double* grad) {
  memset(grad, 0, sizeof(double) * (size_t)n_pts);
  for (int r = 0; r < n_rays; ++r) {
    float back = (float)(grad_last[r] * alphainv_last[r]); /* float back_cum, :522 */
    for (int64_t i = i_end[r] - 1; i >= i_start[r]; --i) {
      grad[i] = grad_weights[i] * T[i] - (double)back / (1 - alpha[i] + 1e-10);
      back = (float)fma(grad_weights[i], weight[i], (double)back);
    }
  }
}

/* total_variation_kernel.cu:13-67 */
static inline double clamp1d(double v) { return fmin(fmax(v, -1.0), 1.0); }

EXPORT void orc64_total_variation_add_grad(const double* param, double* grad, float wx, float wy,
                                           float wz, int dense_mode, int64_t N, int64_t sz_i,
                                           int64_t sz_j, int64_t sz_k) {
  (void)wx;
  wy /= 6; /* float, :45-47 */
  wz /= 6;
  const int64_t sjk = sz_j * sz_k;
  for (int64_t idx = 0; idx < N; ++idx) {
    if (!dense_mode && grad[idx] == 0) continue;
    const int64_t k = idx % sz_k, j = idx / sz_k % sz_j, i = idx / sjk % sz_i;
    const double p = param[idx];
    float acc = 0; /* float grad_to_add, :25 */
    if (k != 0) acc = (float)((double)acc + (double)wz * clamp1d(p - param[idx - 1]));
    if (k != sz_k - 1) acc = (float)((double)acc + (double)wz * clamp1d(p - param[idx + 1]));
    if (j != 0) acc = (float)((double)acc + (double)wy * clamp1d(p - param[idx - sz_k]));
    if (j != sz_j - 1) acc = (float)((double)acc + (double)wy * clamp1d(p - param[idx + sz_k]));
    if (i != 0) acc = (float)((double)acc + (double)wz * clamp1d(p - param[idx - sjk]));
    if (i != sz_i - 1) acc = (float)((double)acc + (double)wz * clamp1d(p - param[idx + sjk]));
    grad[idx] += (double)acc;
  }
}

/* adam_upd_kernel.cu:8-132; SASS: m = DFMA(b1, m, (1-b1)*g); v = DFMA(b2, v, g*(g*(1-b2))) */
EXPORT void orc64_adam_upd(double* param, const double* grad, double* exp_avg, double* exp_avg_sq,
                           const double* perlr, int64_t N, int step, float beta1, float beta2,
                           float lr, float eps, int mode) {
  const float step_size = lr * sqrtf(1.f - powf(beta2, (float)step)) / (1.f - powf(beta1, (float)step));
  const double omb1 = (double)(1.f - beta1), omb2 = (double)(1.f - beta2);
  for (int64_t i = 0; i < N; ++i) {
    const double g = grad[i];
    if (mode == 1 && g == 0) continue;
    const double m = fma((double)beta1, exp_avg[i], omb1 * g);
    const double v = fma((double)beta2, exp_avg_sq[i], g * (g * omb2));
    exp_avg[i] = m;
    exp_avg_sq[i] = v;
    const double num = (mode == 2) ? ((double)step_size * perlr[i]) * m : m * (double)step_size;
    param[i] -= num / (sqrt(v) + (double)eps);
  }
}
