// fused_prep.cu -- the "next" rows of the scope table (include/dvgo_b200_prep.h): ray generation,
// training-ray preparation against the occupancy grid, and the whole-grid / all-ray sweeps around the
// per-iteration path (voxel_count_views, occupancy refresh, progressive-growing resize).
//
// Every kernel re-uses the sampling / occupancy / trilinear device code of the fused march kernels
// (fused_scene.cuh, common.cuh), so "which samples exist and which cells they touch" is the same
// bit-exact arithmetic as in the per-iteration path.
#include "fused_scene.cuh"
#include "../../include/dvgo_b200_prep.h"

namespace dvgo {

// ---- N2: ray of one pixel ----------------------------------------------------------------------------
// lib/ray_utils.py:9-47 with torch's elementwise fp32 semantics: every operation is rounded separately
// (no FMA), the 3-term dot product is summed left to right.
struct PixelRay { float ox, oy, oz, dx, dy, dz, vx, vy, vz; };

__device__ __forceinline__ PixelRay pixel_ray(const dvgo_view_t& v, int64_t pix) {
  const int row = static_cast<int>(pix / v.W), col = static_cast<int>(pix % v.W);
  float i = static_cast<float>(v.flip_x ? v.W - 1 - col : col);   // linspace(0, W-1, W) (+ flip)
  float j = static_cast<float>(v.flip_y ? v.H - 1 - row : row);
  if (v.mode == 1) { i = fadd(i, 0.5f); j = fadd(j, 0.5f); }
  const float a = fdiv(fsub(i, v.cx), v.fx);
  const float b0 = fdiv(fsub(j, v.cy), v.fy);
  const float b = v.inverse_y ? b0 : -b0;
  const float c = v.inverse_y ? 1.f : -1.f;
  PixelRay r;
  r.dx = fadd(fadd(fmul(a, v.c2w[0]), fmul(b, v.c2w[1])), fmul(c, v.c2w[2]));
  r.dy = fadd(fadd(fmul(a, v.c2w[4]), fmul(b, v.c2w[5])), fmul(c, v.c2w[6]));
  r.dz = fadd(fadd(fmul(a, v.c2w[8]), fmul(b, v.c2w[9])), fmul(c, v.c2w[10]));
  r.ox = v.c2w[3]; r.oy = v.c2w[7]; r.oz = v.c2w[11];
  const float nrm = sqrtf(fadd(fadd(fmul(r.dx, r.dx), fmul(r.dy, r.dy)), fmul(r.dz, r.dz)));  // :82
  r.vx = fdiv(r.dx, nrm); r.vy = fdiv(r.dy, nrm); r.vz = fdiv(r.dz, nrm);
  if (v.ndc) {  // lib/ray_utils.py:62-79 with near = 1.
    const float t = fdiv(-fadd(1.f, r.oz), r.dz);
    const float ox = fadd(r.ox, fmul(t, r.dx)), oy = fadd(r.oy, fmul(t, r.dy)), oz = fadd(r.oz, fmul(t, r.dz));
    const float o0 = fdiv(fmul(v.ndc_sx, ox), oz);
    const float o1 = fdiv(fmul(v.ndc_sy, oy), oz);
    const float o2 = fadd(1.f, fdiv(2.f, oz));
    const float d0 = fmul(v.ndc_sx, fsub(fdiv(r.dx, r.dz), fdiv(ox, oz)));
    const float d1 = fmul(v.ndc_sy, fsub(fdiv(r.dy, r.dz), fdiv(oy, oz)));
    const float d2 = fdiv(-2.f, oz);
    r.ox = o0; r.oy = o1; r.oz = o2; r.dx = d0; r.dy = d1; r.dz = d2;
  }
  return r;
}

__global__ void __launch_bounds__(256) rays_of_view_kernel(dvgo_view_t v, int64_t pix_begin, int64_t n_pix,
                                                           float* __restrict__ rays_o, float* __restrict__ rays_d,
                                                           float* __restrict__ viewdirs) {
  const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= n_pix) return;
  const PixelRay r = pixel_ray(v, pix_begin + q);
  if (rays_o) { rays_o[3 * q] = r.ox; rays_o[3 * q + 1] = r.oy; rays_o[3 * q + 2] = r.oz; }
  if (rays_d) { rays_d[3 * q] = r.dx; rays_d[3 * q + 1] = r.dy; rays_d[3 * q + 2] = r.dz; }
  if (viewdirs) { viewdirs[3 * q] = r.vx; viewdirs[3 * q + 1] = r.vy; viewdirs[3 * q + 2] = r.vz; }
}

// ---- N1: does a ray touch occupied space? ---------------------------------------------------------------
// One warp per ray, 32 consecutive steps per trip, ballot early exit (lib/dvgo.py:412-423 materialises every
// sample of every ray, looks all of them up, and scatters 1s).
__device__ __forceinline__ bool warp_ray_hits(const SceneDev& sc, float ox, float oy, float oz, float dx,
                                              float dy, float dz, int lane) {
  const TMinMax t = slab_test(ox, oy, oz, dx, dy, dz, sc.lo, sc.hi, sc.near, sc.far);
  const int n = static_cast<int>(n_samples_of(t.t_min, t.t_max, sc.stepdist));
  const RayGeom g = ray_geom_v(sc, ox, oy, oz, dx, dy, dz, t.t_min);
  for (int base = 0; base < n; base += kWarp) {
    const int i = base + lane;
    bool h = false;
    if (i < n) {
      float px, py, pz;
      sample_point(sc, g, i, px, py, pz);
      if (!out_of_bbox(px, py, pz, sc.lo, sc.hi)) h = occupancy(sc, px, py, pz);
    }
    if (__any_sync(0xffffffffu, h)) return true;
  }
  return false;
}

__global__ void __launch_bounds__(256) rays_hit_kernel(const float* __restrict__ rays_o,
                                                       const float* __restrict__ rays_d, SceneArgs a,
                                                       int64_t n_rays, uint8_t* __restrict__ hit) {
  const SceneDev sc = load_scene(a);
  const int lane = threadIdx.x & 31;
  const int64_t warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t r = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; r < n_rays; r += warps) {
    const bool h = warp_ray_hits(sc, rays_o[3 * r], rays_o[3 * r + 1], rays_o[3 * r + 2], rays_d[3 * r],
                                 rays_d[3 * r + 1], rays_d[3 * r + 2], lane);
    if (lane == 0) hit[r] = h ? 1 : 0;
  }
}

__global__ void __launch_bounds__(256) view_hit_kernel(dvgo_view_t v, SceneArgs a, uint8_t* __restrict__ hit) {
  const SceneDev sc = load_scene(a);
  const int lane = threadIdx.x & 31;
  const int64_t n = static_cast<int64_t>(v.H) * v.W;
  const int64_t warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t p = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; p < n; p += warps) {
    const PixelRay r = pixel_ray(v, p);
    const bool h = warp_ray_hits(sc, r.ox, r.oy, r.oz, r.dx, r.dy, r.dz, lane);
    if (lane == 0) hit[p] = h ? 1 : 0;
  }
}

__global__ void __launch_bounds__(256) view_gather_kernel(dvgo_view_t v, const uint8_t* __restrict__ hit,
                                                          const int64_t* __restrict__ pos_incl,
                                                          const int64_t* __restrict__ top,
                                                          const float* __restrict__ img, float* __restrict__ rgb_tr,
                                                          float* __restrict__ ro, float* __restrict__ rd,
                                                          float* __restrict__ vd) {
  const int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= static_cast<int64_t>(v.H) * v.W || !hit[p]) return;
  const int64_t q = *top + pos_incl[p] - 1;
  const PixelRay r = pixel_ray(v, p);
  if (img && rgb_tr) { rgb_tr[3 * q] = img[3 * p]; rgb_tr[3 * q + 1] = img[3 * p + 1]; rgb_tr[3 * q + 2] = img[3 * p + 2]; }
  if (ro) { ro[3 * q] = r.ox; ro[3 * q + 1] = r.oy; ro[3 * q + 2] = r.oz; }
  if (rd) { rd[3 * q] = r.dx; rd[3 * q + 1] = r.dy; rd[3 * q + 2] = r.dz; }
  if (vd) { vd[3 * q] = r.vx; vd[3 * q + 1] = r.vy; vd[3 * q + 2] = r.vz; }
}

// ---- N3: voxel_count_views ---------------------------------------------------------------------------------
// lib/dvgo.py:276-291 in torch's separately-rounded fp32 arithmetic; one warp per ray, lane = sample.
__global__ void __launch_bounds__(256) voxel_count_scatter_kernel(
    const float* __restrict__ rays_o, const float* __restrict__ rays_d, int64_t n_rays,
    const float* __restrict__ xyz_min, const float* __restrict__ xyz_max, int X, int Y, int Z, float near, float far,
    float stepdist, int n_samples, float* __restrict__ acc) {
  float lo[3], hi[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) { lo[k] = __ldg(xyz_min + k); hi[k] = __ldg(xyz_max + k); }
  const int lane = threadIdx.x & 31;
  const int64_t warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t r = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; r < n_rays; r += warps) {
    const float ox = rays_o[3 * r], oy = rays_o[3 * r + 1], oz = rays_o[3 * r + 2];
    const float dx = rays_d[3 * r], dy = rays_d[3 * r + 1], dz = rays_d[3 * r + 2];
    const float vx = dx == 0.f ? 1e-6f : dx, vy = dy == 0.f ? 1e-6f : dy, vz = dz == 0.f ? 1e-6f : dz;
    const float ax = fdiv(fsub(hi[0], ox), vx), bx = fdiv(fsub(lo[0], ox), vx);
    const float ay = fdiv(fsub(hi[1], oy), vy), by = fdiv(fsub(lo[1], oy), vy);
    const float az = fdiv(fsub(hi[2], oz), vz), bz = fdiv(fsub(lo[2], oz), vz);
    float t_min = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
    t_min = fminf(fmaxf(t_min, near), far);                                      // .clamp(min=near, max=far)
    const float nrm = sqrtf(fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz)));
    for (int i = lane; i < n_samples; i += kWarp) {
      const float interp = fadd(t_min, fdiv(fmul(stepdist, static_cast<float>(i)), nrm));
      const float px = fadd(ox, fmul(dx, interp)), py = fadd(oy, fmul(dy, interp)), pz = fadd(oz, fmul(dz, interp));
      const Tri t = tri_setup(px, py, pz, lo, hi, X, Y, Z);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int x = t.x0 + (k >> 2), y = t.y0 + ((k >> 1) & 1), z = t.z0 + (k & 1);
        if (x < 0 || x >= X || y < 0 || y >= Y || z < 0 || z >= Z) continue;
        const float w = fmul(fmul((k & 1) ? t.wz1 : t.wz0, ((k >> 1) & 1) ? t.wy1 : t.wy0), (k >> 2) ? t.wx1 : t.wx0);
        atomicAdd(acc + (static_cast<int64_t>(x) * Y + y) * Z + z, w);
      }
    }
  }
}

__global__ void __launch_bounds__(256) voxel_count_commit_kernel(float* __restrict__ acc, float* __restrict__ count,
                                                                 int64_t n) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    if (acc[i] > 1.f) count[i] += 1.f;
    acc[i] = 0.f;
  }
}

// ---- N3: occupancy refresh -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) alpha_grid_kernel(const float* __restrict__ density, int64_t n, float shift,
                                                         float interval, float* __restrict__ alpha) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float e = expf(fadd(density[i], shift));          // render_utils_kernel.cu:366
    alpha[i] = fsub(1.f, powf(fadd(1.f, e), -interval));    // :368
  }
}

__global__ void __launch_bounds__(256) maxpool_mask_kernel(const float* __restrict__ alpha, int X, int Y, int Z,
                                                           float thres, const uint8_t* __restrict__ mask_in,
                                                           uint8_t* __restrict__ mask_out) {
  const int64_t n = static_cast<int64_t>(X) * Y * Z;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int z = static_cast<int>(i % Z), y = static_cast<int>((i / Z) % Y), x = static_cast<int>(i / (static_cast<int64_t>(Z) * Y));
    float m = -INFINITY;                                    // max_pool3d pads with -inf
    for (int a = max(x - 1, 0); a <= min(x + 1, X - 1); ++a)
      for (int b = max(y - 1, 0); b <= min(y + 1, Y - 1); ++b)
        for (int c = max(z - 1, 0); c <= min(z + 1, Z - 1); ++c)
          m = fmaxf(m, alpha[(static_cast<int64_t>(a) * Y + b) * Z + c]);
    const bool keep = (m > thres) && (mask_in == nullptr || mask_in[i]);
    mask_out[i] = keep ? 1 : 0;
  }
}

// ---- N3: trilinear resize (ATen upsample_trilinear3d, align_corners=True) -----------------------------------
__global__ void __launch_bounds__(256) resize_trilinear_kernel(const float* __restrict__ src, int C, int X, int Y,
                                                               int Z, float* __restrict__ dst, int X2, int Y2, int Z2) {
  const float rx = X2 > 1 ? fdiv(static_cast<float>(X - 1), static_cast<float>(X2 - 1)) : 0.f;
  const float ry = Y2 > 1 ? fdiv(static_cast<float>(Y - 1), static_cast<float>(Y2 - 1)) : 0.f;
  const float rz = Z2 > 1 ? fdiv(static_cast<float>(Z - 1), static_cast<float>(Z2 - 1)) : 0.f;
  const int64_t n2 = static_cast<int64_t>(X2) * Y2 * Z2, n1 = static_cast<int64_t>(X) * Y * Z;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n2;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int z2 = static_cast<int>(i % Z2), y2 = static_cast<int>((i / Z2) % Y2), x2 = static_cast<int>(i / (static_cast<int64_t>(Z2) * Y2));
    const float fx = fmul(rx, static_cast<float>(x2)), fy = fmul(ry, static_cast<float>(y2)), fz = fmul(rz, static_cast<float>(z2));
    const int x1 = static_cast<int>(fx), y1 = static_cast<int>(fy), z1 = static_cast<int>(fz);
    const int xp = x1 < X - 1 ? 1 : 0, yp = y1 < Y - 1 ? 1 : 0, zp = z1 < Z - 1 ? 1 : 0;
    const float lx1 = fsub(fx, static_cast<float>(x1)), lx0 = fsub(1.f, lx1);
    const float ly1 = fsub(fy, static_cast<float>(y1)), ly0 = fsub(1.f, ly1);
    const float lz1 = fsub(fz, static_cast<float>(z1)), lz0 = fsub(1.f, lz1);
    const int64_t o00 = (static_cast<int64_t>(x1) * Y + y1) * Z + z1;
    const int64_t sy = static_cast<int64_t>(yp) * Z, sx = static_cast<int64_t>(xp) * Y * Z;
    for (int c = 0; c < C; ++c) {
      const float* s = src + c * n1 + o00;
      const float v =
          lx0 * (ly0 * (lz0 * s[0] + lz1 * s[zp]) + ly1 * (lz0 * s[sy] + lz1 * s[sy + zp])) +
          lx1 * (ly0 * (lz0 * s[sx] + lz1 * s[sx + zp]) + ly1 * (lz0 * s[sx + sy] + lz1 * s[sx + sy + zp]));
      dst[c * n2 + i] = v;
    }
  }
}

static inline int grid_for(int64_t n, int threads, int per_sm) {
  const int64_t want = (n + threads - 1) / threads;
  const int64_t cap = static_cast<int64_t>(kNumSMs) * per_sm;
  return static_cast<int>(want < 1 ? 1 : (want > cap ? cap : want));
}

}  // namespace dvgo

using namespace dvgo;

DVGO_API int dvgo_rays_of_view(const dvgo_view_t* v, int64_t pix_begin, int64_t n_pix, float* rays_o, float* rays_d,
                               float* viewdirs, dvgo_stream_t stream) {
  if (!v || pix_begin < 0 || n_pix < 0 || pix_begin + n_pix > static_cast<int64_t>(v->H) * v->W) return DVGO_EINVAL;
  if (n_pix == 0) return 0;
  rays_of_view_kernel<<<blocks_for(n_pix, 256), 256, 0, as_stream(stream)>>>(*v, pix_begin, n_pix, rays_o, rays_d, viewdirs);
  return launch_status();
}

DVGO_API int dvgo_hit_coarse_geo(const dvgo_scene_t* scene, const float* rays_o, const float* rays_d, int64_t n_rays,
                                 uint8_t* hit, dvgo_stream_t stream) {
  if (!scene || n_rays < 0 || scene->ndc) return DVGO_EINVAL;
  if (n_rays == 0) return 0;
  rays_hit_kernel<<<grid_for(n_rays * 32, 256, 16), 256, 0, as_stream(stream)>>>(rays_o, rays_d, to_args(scene), n_rays, hit);
  return launch_status();
}

DVGO_API int dvgo_view_hit_coarse_geo(const dvgo_view_t* v, const dvgo_scene_t* scene, uint8_t* hit,
                                      dvgo_stream_t stream) {
  if (!v || !scene || scene->ndc) return DVGO_EINVAL;
  const int64_t n = static_cast<int64_t>(v->H) * v->W;
  if (n == 0) return 0;
  view_hit_kernel<<<grid_for(n * 32, 256, 16), 256, 0, as_stream(stream)>>>(*v, to_args(scene), hit);
  return launch_status();
}

DVGO_API int dvgo_view_gather_rays(const dvgo_view_t* v, const uint8_t* hit, const int64_t* pos_incl,
                                   const int64_t* top, const float* img, float* rgb_tr, float* rays_o_tr,
                                   float* rays_d_tr, float* viewdirs_tr, dvgo_stream_t stream) {
  if (!v || !hit || !pos_incl || !top) return DVGO_EINVAL;
  const int64_t n = static_cast<int64_t>(v->H) * v->W;
  if (n == 0) return 0;
  view_gather_kernel<<<blocks_for(n, 256), 256, 0, as_stream(stream)>>>(*v, hit, pos_incl, top, img, rgb_tr, rays_o_tr,
                                                                       rays_d_tr, viewdirs_tr);
  return launch_status();
}

DVGO_API int dvgo_voxel_count_scatter(const float* rays_o, const float* rays_d, int64_t n_rays, const float* xyz_min,
                                      const float* xyz_max, int X, int Y, int Z, float near, float far, float stepdist,
                                      int n_samples, float* acc, dvgo_stream_t stream) {
  if (n_rays < 0 || n_samples < 0 || X <= 0 || Y <= 0 || Z <= 0) return DVGO_EINVAL;
  if (n_rays == 0 || n_samples == 0) return 0;
  voxel_count_scatter_kernel<<<grid_for(n_rays * 32, 256, 16), 256, 0, as_stream(stream)>>>(
      rays_o, rays_d, n_rays, xyz_min, xyz_max, X, Y, Z, near, far, stepdist, n_samples, acc);
  return launch_status();
}

DVGO_API int dvgo_voxel_count_commit(float* acc, float* count, int64_t n, dvgo_stream_t stream) {
  if (n < 0) return DVGO_EINVAL;
  if (n == 0) return 0;
  voxel_count_commit_kernel<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(acc, count, n);
  return launch_status();
}

DVGO_API int dvgo_alpha_maxpool_mask(const float* density, int X, int Y, int Z, float act_shift, float interval,
                                     float thres, const uint8_t* mask_in, uint8_t* mask_out, float* alpha_tmp,
                                     dvgo_stream_t stream) {
  if (X <= 0 || Y <= 0 || Z <= 0 || !alpha_tmp || !mask_out) return DVGO_EINVAL;
  const int64_t n = static_cast<int64_t>(X) * Y * Z;
  alpha_grid_kernel<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(density, n, act_shift, interval, alpha_tmp);
  maxpool_mask_kernel<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(alpha_tmp, X, Y, Z, thres, mask_in, mask_out);
  return launch_status(2);
}

DVGO_API int dvgo_resize_trilinear(const float* src, int C, int X, int Y, int Z, float* dst, int X2, int Y2, int Z2,
                                   dvgo_stream_t stream) {
  if (C <= 0 || X <= 0 || Y <= 0 || Z <= 0 || X2 <= 0 || Y2 <= 0 || Z2 <= 0) return DVGO_EINVAL;
  const int64_t n2 = static_cast<int64_t>(X2) * Y2 * Z2;
  resize_trilinear_kernel<<<grid_for(n2, 256, 8), 256, 0, as_stream(stream)>>>(src, C, X, Y, Z, dst, X2, Y2, Z2);
  return launch_status();
}
