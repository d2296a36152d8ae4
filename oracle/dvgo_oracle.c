/*
 * dvgo_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-threaded CPU restatement of the reference's per-ray volume-rendering and
 * grid-optimisation kernels (hbell99/DirectVoxGO, lib/cuda/*.cu).  It is the checker for the
 * CUDA product in directvoxgo_b200/csrc; nothing in the product imports, links or executes it.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
 *
 * Every function cites the reference file:line it follows.  Where the reference is compiled by
 * nvcc with -fmad=true, the contraction the compiler performs (verified in
 * the SASS of oracle/_ref) is written out with fmaf() so that the integer / boolean outputs
 * (N_steps, ray_id, step_id, mask_outbbox, maskcache) are reproduced bit-for-bit on the CPU.
 * Build: gcc -O2 -ffp-contract=off (see oracle/Makefile) -- the compiler must not fuse on its own.
 *
 * Pinning: the reference ships no tests / golden vectors.  This file is
 * pinned against outputs of the reference's own CUDA kernels (oracle/_ref, built from the
 * unmodified sources) recorded on a B200 into tests/golden/ref_gpu_*.npz by
 * oracle/make_golden_gpu.py, and against the GPU reference live in tests/test_gpu_vs_ref.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define EXPORT __attribute__((visibility("default")))

static inline float fminf_(float a, float b) { return a < b ? a : b; }
static inline float fmaxf_(float a, float b) { return a > b ? a : b; }

/* ---- K1: infer_t_minmax  (lib/cuda/render_utils_kernel.cu:12-35) ---------------------------- */
EXPORT void orc_infer_t_minmax(const float* rays_o, const float* rays_d, const float* xyz_min,
                               const float* xyz_max, float near, float far, int n_rays,
                               float* t_min, float* t_max) {
  for (int r = 0; r < n_rays; ++r) {
    const float* o = rays_o + 3 * r;
    const float* d = rays_d + 3 * r;
    float a[3], b[3];
    for (int c = 0; c < 3; ++c) {
      /* :23-25  (d==0) ? 1e-6 : d  -- a double conditional narrowed to float */
      float v = (float)((d[c] == 0) ? 1e-6 : (double)d[c]);
      a[c] = (xyz_max[c] - o[c]) / v; /* :26-28 */
      b[c] = (xyz_min[c] - o[c]) / v; /* :29-31 */
    }
    /* :32-33 */
    float lo = fmaxf_(fmaxf_(fminf_(a[0], b[0]), fminf_(a[1], b[1])), fminf_(a[2], b[2]));
    float hi = fminf_(fminf_(fmaxf_(a[0], b[0]), fmaxf_(a[1], b[1])), fmaxf_(a[2], b[2]));
    t_min[r] = fmaxf_(fminf_(lo, far), near);
    t_max[r] = fmaxf_(fminf_(hi, far), near);
  }
}

/* ---- K2: infer_n_samples  (render_utils_kernel.cu:38-49) ------------------------------------ */
EXPORT void orc_infer_n_samples(const float* t_min, const float* t_max, float stepdist, int n_rays,
                                int64_t* n_samples) {
  for (int r = 0; r < n_rays; ++r) {
    double c = (double)ceilf((t_max[r] - t_min[r]) / stepdist); /* :47 float math, then double max */
    n_samples[r] = (int64_t)(c > 1. ? c : 1.);
  }
}

/* ---- K3: infer_ray_start_dir  (render_utils_kernel.cu:52-73) -------------------------------- */
EXPORT void orc_infer_ray_start_dir(const float* rays_o, const float* rays_d, const float* t_min,
                                    int n_rays, float* rays_start, float* rays_dir) {
  for (int r = 0; r < n_rays; ++r) {
    const float* o = rays_o + 3 * r;
    const float* d = rays_d + 3 * r;
    /* :62-65  dx*dx + dy*dy + dz*dz: the SASS of oracle/_ref shows FMUL(dy,dy), FFMA(dx,dx,.),
     * FFMA(dz,dz,.) -- nvcc fuses the LEFT product of a*a + b*b and keeps the right one plain. */
    float ss = fmaf(d[2], d[2], fmaf(d[0], d[0], d[1] * d[1]));
    float rnorm = sqrtf(ss);
    for (int c = 0; c < 3; ++c) {
      rays_start[3 * r + c] = fmaf(d[c], t_min[r], o[c]); /* :66-68 contracted */
      rays_dir[3 * r + c] = d[c] / rnorm;                 /* :69-71 */
    }
  }
}

/* ---- K4-K6: sample_pts_on_rays  (render_utils_kernel.cu:138-236) ---------------------------- */
/* Phase 1: t_min/t_max/N_steps and the total (the reference's host sync, :206). */
EXPORT int64_t orc_sample_pts_count(const float* rays_o, const float* rays_d, const float* xyz_min,
                                    const float* xyz_max, float near, float far, float stepdist,
                                    int n_rays, float* t_min, float* t_max, int64_t* N_steps) {
  orc_infer_t_minmax(rays_o, rays_d, xyz_min, xyz_max, near, far, n_rays, t_min, t_max);
  orc_infer_n_samples(t_min, t_max, stepdist, n_rays, N_steps);
  int64_t total = 0;
  for (int r = 0; r < n_rays; ++r) total += N_steps[r];
  return total;
}

/* Phase 2: ray_id / step_id (:138-158, 207-213) and the points + out-of-bbox mask (:161-188). */
EXPORT void orc_sample_pts_fill(const float* rays_o, const float* rays_d, const float* xyz_min,
                                const float* xyz_max, const float* t_min, const int64_t* N_steps,
                                float stepdist, int n_rays, float* rays_pts, uint8_t* mask_outbbox,
                                int64_t* ray_id, int64_t* step_id) {
  float* start = (float*)malloc(sizeof(float) * 3 * (size_t)n_rays);
  float* dir = (float*)malloc(sizeof(float) * 3 * (size_t)n_rays);
  orc_infer_ray_start_dir(rays_o, rays_d, t_min, n_rays, start, dir);
  int64_t idx = 0;
  for (int r = 0; r < n_rays; ++r) {
    for (int64_t s = 0; s < N_steps[r]; ++s, ++idx) {
      ray_id[idx] = r;
      step_id[idx] = s;
      const float dist = stepdist * (float)(int)s; /* :178 int -> float, FMUL */
      float p[3];
      for (int c = 0; c < 3; ++c) {
        p[c] = fmaf(dir[3 * r + c], dist, start[3 * r + c]); /* :179-181 contracted */
        rays_pts[3 * idx + c] = p[c];
      }
      mask_outbbox[idx] = (xyz_min[0] > p[0]) | (xyz_min[1] > p[1]) | (xyz_min[2] > p[2]) |
                          (xyz_max[0] < p[0]) | (xyz_max[1] < p[1]) | (xyz_max[2] < p[2]); /* :185 */
    }
  }
  free(start);
  free(dir);
}

/* ---- K7: sample_ndc_pts_on_rays  (render_utils_kernel.cu:239-264) --------------------------- */
EXPORT void orc_sample_ndc_pts_on_rays(const float* rays_o, const float* rays_d,
                                       const float* xyz_min, const float* xyz_max, int N_samples,
                                       int n_rays, float* rays_pts, uint8_t* mask_outbbox) {
  for (int r = 0; r < n_rays; ++r) {
    for (int s = 0; s < N_samples; ++s) {
      const int64_t idx = (int64_t)r * N_samples + s;
      const float dist = ((float)s) / (float)(N_samples - 1); /* :254 */
      float p[3];
      for (int c = 0; c < 3; ++c) {
        p[c] = fmaf(rays_d[3 * r + c], dist, rays_o[3 * r + c]); /* :255-257 contracted */
        rays_pts[3 * idx + c] = p[c];
      }
      mask_outbbox[idx] = (xyz_min[0] > p[0]) | (xyz_min[1] > p[1]) | (xyz_min[2] > p[2]) |
                          (xyz_max[0] < p[0]) | (xyz_max[1] < p[1]) | (xyz_max[2] < p[2]);
    }
  }
}

/* ---- K8: maskcache_lookup  (render_utils_kernel.cu:294-351) --------------------------------- */
EXPORT void orc_maskcache_lookup(const uint8_t* world, const float* xyz, const float* scale,
                                 const float* shift, int sz_i, int sz_j, int sz_k, int64_t n_pts,
                                 uint8_t* out) {
  for (int64_t p = 0; p < n_pts; ++p) {
    /* :312-314 x*scale+shift contracted to FFMA, then round() = half away from zero */
    const int i = (int)roundf(fmaf(xyz[3 * p + 0], scale[0], shift[0]));
    const int j = (int)roundf(fmaf(xyz[3 * p + 1], scale[1], shift[1]));
    const int k = (int)roundf(fmaf(xyz[3 * p + 2], scale[2], shift[2]));
    out[p] = 0; /* zero-initialised output, :332 */
    if (0 <= i && i < sz_i && 0 <= j && j < sz_j && 0 <= k && k < sz_k)
      out[p] = world[(int64_t)i * sz_j * sz_k + (int64_t)j * sz_k + k];
  }
}

/* ---- K9/K10: raw2alpha and its backward  (render_utils_kernel.cu:358-428) ------------------- */
EXPORT void orc_raw2alpha(const float* density, float shift, float interval, int64_t n,
                          float* exp_d, float* alpha) {
  for (int64_t i = 0; i < n; ++i) {
    const float e = expf(density[i] + shift); /* :366, may be inf */
    exp_d[i] = e;
    alpha[i] = 1.f - powf(1.f + e, -interval); /* :368 */
  }
}

EXPORT void orc_raw2alpha_backward(const float* exp_d, const float* grad_back, float interval,
                                   int64_t n, float* grad) {
  for (int64_t i = 0; i < n; ++i) {
    /* :404  min(e,1e10) promotes the product chain to double; pow stays float */
    const double m = (double)exp_d[i] < 1e10 ? (double)exp_d[i] : 1e10;
    const double pw = (double)powf(1.f + exp_d[i], -interval - 1.f);
    grad[i] = (float)(m * pw * (double)interval * (double)grad_back[i]);
  }
}

/* ---- K11/K12: alpha2weight  (render_utils_kernel.cu:431-505) -------------------------------- */
EXPORT void orc_alpha2weight(const float* alpha, const int64_t* ray_id, int n_rays, int64_t n_pts,
                             float* weight, float* T, float* alphainv_last, int64_t* i_start,
                             int64_t* i_end) {
  for (int64_t i = 0; i < n_pts; ++i) { weight[i] = 0.f; T[i] = 1.f; } /* :478-479 */
  for (int r = 0; r < n_rays; ++r) { alphainv_last[r] = 1.f; i_start[r] = 0; i_end[r] = 0; }
  if (n_pts == 0) return; /* :483 */
  for (int64_t i = 1; i < n_pts; ++i) /* :461-471 */
    if (ray_id[i] != ray_id[i - 1]) { i_start[ray_id[i]] = i; i_end[ray_id[i - 1]] = i; }
  i_end[ray_id[n_pts - 1]] = n_pts; /* :489 */
  for (int r = 0; r < n_rays; ++r) { /* :440-458 */
    const int i_s = (int)i_start[r], i_e_max = (int)i_end[r];
    float T_cum = 1.f;
    int i;
    for (i = i_s; i < i_e_max; ++i) {
      T[i] = T_cum;
      weight[i] = T_cum * alpha[i];
      T_cum = (float)((double)T_cum * ((1. - (double)alpha[i]) + 1e-10)); /* :450 double math */
      if ((double)T_cum < 1e-3) { i += 1; break; }
    }
    i_end[r] = i;
    alphainv_last[r] = T_cum;
  }
}

/* ---- K13: alpha2weight_backward  (render_utils_kernel.cu:508-561) --------------------------- */
EXPORT void orc_alpha2weight_backward(const float* alpha, const float* weight, const float* T,
                                      const float* alphainv_last, const int64_t* i_start,
                                      const int64_t* i_end, int n_rays, int64_t n_pts,
                                      const float* grad_weights, const float* grad_last,
                                      float* grad) {
  for (int64_t i = 0; i < n_pts; ++i) grad[i] = 0.f; /* :538 */
  for (int r = 0; r < n_rays; ++r) {
    const int i_s = (int)i_start[r], i_e = (int)i_end[r];
    float back_cum = grad_last[r] * alphainv_last[r]; /* :525 */
    for (int i = i_e - 1; i >= i_s; --i) {
      /* :527  (1-alpha) is a float subtract, then promoted; divide and subtract in double */
      const float gwT = grad_weights[i] * T[i];
      const float one_m_a = 1.f - alpha[i];
      grad[i] = (float)((double)gwT - (double)back_cum / ((double)one_m_a + 1e-10));
      back_cum = fmaf(grad_weights[i], weight[i], back_cum); /* :528 contracted */
    }
  }
}

/* ---- T1: trilinear DenseGrid sampling ------------------------------------------------------- */
/* lib/dvgo.py:312-328 -> F.grid_sample(grid[1,C,X,Y,Z], ind_norm, 'bilinear', align_corners=True,
 * zero padding).  ind_norm = ((xyz-min)/(max-min)).flip(-1)*2-1 (dvgo.py:316) and ATen maps it
 * back with ((c+1)/2)*(size-1)  (ATen/native/GridSampler.h grid_sampler_unnormalize).
 * The flip makes xyz[0] index the X (=D) axis of the grid, xyz[2] the Z (=W, contiguous) axis. */
static inline float unnorm_coord(float x, float lo, float hi, int size) {
  const float u = (x - lo) / (hi - lo); /* dvgo.py:316, separate torch ops */
  const float n = u * 2.f - 1.f;        /* two ops (mul, sub): computed unfused in torch */
  return ((n + 1.f) / 2.f) * (float)(size - 1);
}

/* Corner geometry exactly as ATen's grid_sampler_3d kernel computes it (ATen (ix,iy,iz) = DVGO
 * (z,y,x) because of the flip): integer corner = floor(coord); the weight of the low corner along an
 * axis is (float)(i0+1) - f, of the high corner f - (float)i0; the 3-factor product is formed as
 * (wW * wH) * wD = (wz * wy) * wx; corners are accumulated in the order tnw,tne,tsw,tse,bnw,...
 * = (x0y0z0, x0y0z1, x0y1z0, x0y1z1, x1y0z0, ...), each as out += v * w (an FFMA on the GPU). */
typedef struct { int x0, y0, z0; float wx[2], wy[2], wz[2]; } tri_t;

static inline tri_t tri_setup(const float* p, const float* lo, const float* hi, int X, int Y, int Z) {
  tri_t t;
  const float fx = unnorm_coord(p[0], lo[0], hi[0], X);
  const float fy = unnorm_coord(p[1], lo[1], hi[1], Y);
  const float fz = unnorm_coord(p[2], lo[2], hi[2], Z);
  t.x0 = (int)floorf(fx); t.y0 = (int)floorf(fy); t.z0 = (int)floorf(fz);
  t.wx[0] = (float)(t.x0 + 1) - fx; t.wx[1] = fx - (float)t.x0;
  t.wy[0] = (float)(t.y0 + 1) - fy; t.wy[1] = fy - (float)t.y0;
  t.wz[0] = (float)(t.z0 + 1) - fz; t.wz[1] = fz - (float)t.z0;
  return t;
}

EXPORT void orc_grid_sample_3d(const float* grid, int C, int X, int Y, int Z, const float* xyz,
                               const float* xyz_min, const float* xyz_max, int64_t n_pts,
                               float* out /* [n_pts, C] */) {
  const int64_t plane = (int64_t)X * Y * Z;
  for (int64_t p = 0; p < n_pts; ++p) {
    const tri_t t = tri_setup(xyz + 3 * p, xyz_min, xyz_max, X, Y, Z);
    for (int c = 0; c < C; ++c) out[p * C + c] = 0.f;
    for (int corner = 0; corner < 8; ++corner) {
      const int dx = corner >> 2, dy = (corner >> 1) & 1, dz = corner & 1;
      const int xi = t.x0 + dx, yi = t.y0 + dy, zi = t.z0 + dz;
      if (xi < 0 || xi >= X || yi < 0 || yi >= Y || zi < 0 || zi >= Z) continue; /* zero pad */
      const float w = (t.wz[dz] * t.wy[dy]) * t.wx[dx];
      const int64_t off = ((int64_t)xi * Y + yi) * Z + zi;
      for (int c = 0; c < C; ++c) out[p * C + c] = fmaf(grid[c * plane + off], w, out[p * C + c]);
    }
  }
}

/* Backward w.r.t. the grid: grad_grid[1,C,X,Y,Z] += w * grad_out[p,c]  (ATen grid_sampler_3d_backward,
 * the only grad DirectVoxGO needs: xyz carries no grad). */
EXPORT void orc_grid_sample_3d_backward(const float* grad_out /* [n_pts,C] */, int C, int X, int Y,
                                        int Z, const float* xyz, const float* xyz_min,
                                        const float* xyz_max, int64_t n_pts, float* grad_grid) {
  const int64_t plane = (int64_t)X * Y * Z;
  for (int64_t p = 0; p < n_pts; ++p) {
    const tri_t t = tri_setup(xyz + 3 * p, xyz_min, xyz_max, X, Y, Z);
    for (int corner = 0; corner < 8; ++corner) {
      const int dx = corner >> 2, dy = (corner >> 1) & 1, dz = corner & 1;
      const int xi = t.x0 + dx, yi = t.y0 + dy, zi = t.z0 + dz;
      if (xi < 0 || xi >= X || yi < 0 || yi >= Y || zi < 0 || zi >= Z) continue;
      const float w = (t.wz[dz] * t.wy[dy]) * t.wx[dx];
      const int64_t off = ((int64_t)xi * Y + yi) * Z + zi;
      for (int c = 0; c < C; ++c) grad_grid[c * plane + off] += w * grad_out[p * C + c];
    }
  }
}

/* ---- T2: tri-plane bilinear sampling = ATen grid_sampler_2d as lib/tri_dvgo.py:456-464 calls it ----------------
 * plane [C,H,W]; axis_w / axis_h: world axis whose normalised coordinate indexes W / H (the reference passes
 * ind_norm[..., [i,j]] with ind_norm = flipped (z,y,x)).  Corner order and weights as ATen: nw, ne, sw, se. */
static inline float unnorm1(float x, float lo, float hi, int size) {
  const float u = (x - lo) / (hi - lo);
  const float n = u * 2.f - 1.f;
  return ((n + 1.f) * 0.5f) * (float)(size - 1);
}
EXPORT void orc_grid_sample_2d(const float* plane, int C, int H, int W, const float* xyz, const float* xyz_min,
                               const float* xyz_max, int axis_w, int axis_h, int64_t n_pts, float* out) {
  const int64_t hw = (int64_t)H * W;
  for (int64_t p = 0; p < n_pts; ++p) {
    const float ix = unnorm1(xyz[3 * p + axis_w], xyz_min[axis_w], xyz_max[axis_w], W);
    const float iy = unnorm1(xyz[3 * p + axis_h], xyz_min[axis_h], xyz_max[axis_h], H);
    const float x0 = floorf(ix), y0 = floorf(iy);
    const float ex = (x0 + 1.f) - ix, wx = ix - x0, ey = (y0 + 1.f) - iy, wy = iy - y0;
    const float w[4] = {ex * ey, wx * ey, ex * wy, wx * wy};
    for (int c = 0; c < C; ++c) out[p * C + c] = 0.f;
    for (int k = 0; k < 4; ++k) {
      const int x = (int)x0 + (k & 1), y = (int)y0 + (k >> 1);
      if (x < 0 || x >= W || y < 0 || y >= H) continue;
      for (int c = 0; c < C; ++c) out[p * C + c] = fmaf(plane[c * hw + (int64_t)y * W + x], w[k], out[p * C + c]);
    }
  }
}
EXPORT void orc_grid_sample_2d_backward(const float* grad_out, int C, int H, int W, const float* xyz,
                                        const float* xyz_min, const float* xyz_max, int axis_w, int axis_h,
                                        int64_t n_pts, float* grad_plane) {
  const int64_t hw = (int64_t)H * W;
  for (int64_t p = 0; p < n_pts; ++p) {
    const float ix = unnorm1(xyz[3 * p + axis_w], xyz_min[axis_w], xyz_max[axis_w], W);
    const float iy = unnorm1(xyz[3 * p + axis_h], xyz_min[axis_h], xyz_max[axis_h], H);
    const float x0 = floorf(ix), y0 = floorf(iy);
    const float ex = (x0 + 1.f) - ix, wx = ix - x0, ey = (y0 + 1.f) - iy, wy = iy - y0;
    const float w[4] = {ex * ey, wx * ey, ex * wy, wx * wy};
    for (int k = 0; k < 4; ++k) {
      const int x = (int)x0 + (k & 1), y = (int)y0 + (k >> 1);
      if (x < 0 || x >= W || y < 0 || y >= H) continue;
      for (int c = 0; c < C; ++c) grad_plane[c * hw + (int64_t)y * W + x] += w[k] * grad_out[p * C + c];
    }
  }
}

/* ---- T3: torch_scatter.segment_coo(src, index, out, reduce='sum')  (lib/dvgo.py:554-558) ----- */
EXPORT void orc_segment_coo_sum(const float* src, const int64_t* index, int64_t n_pts, int D,
                                float* out /* [n_seg, D], accumulated into */) {
  for (int64_t p = 0; p < n_pts; ++p)
    for (int d = 0; d < D; ++d) out[index[p] * D + d] += src[p * D + d];
}

/* ---- K14: total_variation_add_grad  (lib/cuda/total_variation_kernel.cu:13-67) -------------- */
static inline float clamp1(float v) { return fminf_(fmaxf_(v, -1.f), 1.f); }

EXPORT void orc_total_variation_add_grad(const float* param, float* grad, float wx, float wy,
                                         float wz, int dense_mode, int64_t N, int64_t sz_i,
                                         int64_t sz_j, int64_t sz_k) {
  (void)wx;
  wx /= 6; wy /= 6; wz /= 6; /* :45-47 */
  /* TV reads param only, so the in-place grad update has no ordering hazard. */
  for (int64_t idx = 0; idx < N; ++idx) {
    if (!(dense_mode || grad[idx] != 0)) continue; /* :21 */
    const int64_t k = idx % sz_k, j = idx / sz_k % sz_j, i = idx / sz_k / sz_j % sz_i;
    /* :26-32.  SASS of oracle/_ref: each term is a predicated FMUL (w*clamp, or 0 at the border)
     * and the six terms are summed left to right with FADD -- no FMA contraction here. */
    float g = 0;
    g += (k == 0 ? 0.f : wz * clamp1(param[idx] - param[idx - 1]));
    g += (k == sz_k - 1 ? 0.f : wz * clamp1(param[idx] - param[idx + 1]));
    g += (j == 0 ? 0.f : wy * clamp1(param[idx] - param[idx - sz_k]));
    g += (j == sz_j - 1 ? 0.f : wy * clamp1(param[idx] - param[idx + sz_k]));
    g += (i == 0 ? 0.f : wz * clamp1(param[idx] - param[idx - sz_k * sz_j])); /* wz: sic */
    g += (i == sz_i - 1 ? 0.f : wz * clamp1(param[idx] - param[idx + sz_k * sz_j]));
    grad[idx] += g; /* :33 */
  }
}

/* ---- K15-K17: the three Adam updates  (lib/cuda/adam_upd_kernel.cu:8-132) ------------------- */
static inline float adam_step_size(int step, float beta1, float beta2, float lr) {
  /* :72  lr * sqrt(1 - pow(beta2,(float)step)) / (1 - pow(beta1,(float)step)); host float math */
  return lr * sqrtf(1.f - powf(beta2, (float)step)) / (1.f - powf(beta1, (float)step));
}

/* mode 0 = adam_upd, 1 = masked_adam_upd (skip grad==0), 2 = adam_upd_with_perlr */
EXPORT void orc_adam_upd(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                         const float* perlr, int64_t N, int step, float beta1, float beta2,
                         float lr, float eps, int mode) {
  const float step_size = adam_step_size(step, beta1, beta2, lr);
  for (int64_t i = 0; i < N; ++i) {
    if (mode == 1 && grad[i] == 0) continue; /* :35 */
    /* :19-21 (and :36-38, :54-56).  nvcc contracts  b*m + (1-b)*g  into  fma(b, m, (1-b)*g). */
    const float g = grad[i];
    const float m = fmaf(beta1, exp_avg[i], (1.f - beta1) * g);
    const float v = fmaf(beta2, exp_avg_sq[i], ((1.f - beta2) * g) * g);
    exp_avg[i] = m;
    exp_avg_sq[i] = v;
    float num = step_size * m;
    if (mode == 2) num = (step_size * perlr[i]) * m; /* :56 step_size * perlr * m, left to right */
    param[i] -= num / (sqrtf(v) + eps);
  }
}

EXPORT float orc_adam_step_size(int step, float beta1, float beta2, float lr) {
  return adam_step_size(step, beta1, beta2, lr);
}
