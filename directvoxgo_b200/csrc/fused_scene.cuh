// fused_scene.cuh -- scene description + per-ray / per-sample geometry shared by the fused kernels
// (fused_march.cu, fused_prep.cu).  Arithmetic follows the reference expression trees exactly where
// integer / boolean results depend on it (common.cuh).
#pragma once
#include "common.cuh"
#include "../../include/dvgo_b200_fused.h"

namespace dvgo {

struct SceneDev {
  int X, Y, Z, C;
  float lo[3], hi[3];
  const uint8_t* mask;
  int mx, my, mz;
  float mscale[3], mshift[3];
  float near, far, stepdist, act_shift, interval, thres;
  int ndc, ndc_samples;
};

// The three [3] vectors live in device memory (they are torch buffers); fetch them once per thread.
struct SceneArgs {
  int X, Y, Z, C;
  const float* xyz_min;
  const float* xyz_max;
  const uint8_t* mask;
  int mx, my, mz;
  const float* mask_scale;
  const float* mask_shift;
  float near, far, stepdist, act_shift, interval, thres;
  int ndc, ndc_samples;
};

static inline SceneArgs to_args(const dvgo_scene_t* s) {
  SceneArgs a;
  a.X = s->X; a.Y = s->Y; a.Z = s->Z; a.C = s->C;
  a.xyz_min = s->xyz_min; a.xyz_max = s->xyz_max;
  a.mask = s->mask; a.mx = s->mx; a.my = s->my; a.mz = s->mz;
  a.mask_scale = s->mask_scale; a.mask_shift = s->mask_shift;
  a.near = s->near; a.far = s->far; a.stepdist = s->stepdist; a.act_shift = s->act_shift;
  a.interval = s->interval; a.thres = s->fast_color_thres;
  a.ndc = s->ndc; a.ndc_samples = s->ndc_samples;
  return a;
}

__device__ __forceinline__ SceneDev load_scene(const SceneArgs& a) {
  SceneDev s;
  s.X = a.X; s.Y = a.Y; s.Z = a.Z; s.C = a.C;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    s.lo[i] = __ldg(a.xyz_min + i);
    s.hi[i] = __ldg(a.xyz_max + i);
    s.mscale[i] = a.mask ? __ldg(a.mask_scale + i) : 0.f;
    s.mshift[i] = a.mask ? __ldg(a.mask_shift + i) : 0.f;
  }
  s.mask = a.mask; s.mx = a.mx; s.my = a.my; s.mz = a.mz;
  s.near = a.near; s.far = a.far; s.stepdist = a.stepdist; s.act_shift = a.act_shift;
  s.interval = a.interval; s.thres = a.thres; s.ndc = a.ndc; s.ndc_samples = a.ndc_samples;
  return s;
}

// Per-ray constants: where sample i of the ray lies.
struct RayGeom {
  float sx, sy, sz, ux, uy, uz;  // p_i = s + u * dist_i
  float inv_ndc;                 // ndc: dist_i = i / (N-1)
};

__device__ __forceinline__ RayGeom ray_geom_v(const SceneDev& sc, float ox, float oy, float oz, float dx,
                                              float dy, float dz, float t_min) {
  RayGeom g;
  if (sc.ndc) {  // lib/cuda/render_utils_kernel.cu:254-257: p = o + d * (i/(N-1))
    g.sx = ox; g.sy = oy; g.sz = oz; g.ux = dx; g.uy = dy; g.uz = dz;
    g.inv_ndc = static_cast<float>(sc.ndc_samples - 1);
  } else {       // :62-71, :178-181: p = (o + d*t_min) + (d/|d|) * (stepdist*i)
    const StartDir s = ray_start_dir(ox, oy, oz, dx, dy, dz, t_min);
    g.sx = s.sx; g.sy = s.sy; g.sz = s.sz; g.ux = s.ux; g.uy = s.uy; g.uz = s.uz;
    g.inv_ndc = 0.f;
  }
  return g;
}
__device__ __forceinline__ RayGeom ray_geom(const SceneDev& sc, const float* __restrict__ rays_o,
                                            const float* __restrict__ rays_d, int r, float t_min) {
  return ray_geom_v(sc, rays_o[3 * r], rays_o[3 * r + 1], rays_o[3 * r + 2], rays_d[3 * r], rays_d[3 * r + 1],
                    rays_d[3 * r + 2], t_min);
}

__device__ __forceinline__ void sample_point(const SceneDev& sc, const RayGeom& g, int i, float& px,
                                             float& py, float& pz) {
  const float dist = sc.ndc ? fdiv(static_cast<float>(i), g.inv_ndc)
                            : fmul(sc.stepdist, static_cast<float>(i));
  px = fma_(g.ux, dist, g.sx);
  py = fma_(g.uy, dist, g.sy);
  pz = fma_(g.uz, dist, g.sz);
}

__device__ __forceinline__ bool occupancy(const SceneDev& sc, float px, float py, float pz) {
  if (!sc.mask) return true;
  const int i = static_cast<int>(roundf(fma_(px, sc.mscale[0], sc.mshift[0])));
  const int j = static_cast<int>(roundf(fma_(py, sc.mscale[1], sc.mshift[1])));
  const int k = static_cast<int>(roundf(fma_(pz, sc.mscale[2], sc.mshift[2])));
  if ((0 <= i) & (i < sc.mx) & (0 <= j) & (j < sc.my) & (0 <= k) & (k < sc.mz))
    return sc.mask[(static_cast<int64_t>(i) * sc.my + j) * sc.mz + k] != 0;
  return false;
}

// 8-corner geometry (ATen order x0y0z0, x0y0z1, x0y1z0, ... ; weight = (wz*wy)*wx; zero padding),
// kept compact (one base voxel index + validity bits + six axis weights) to save registers.
struct Corner8 {
  int base;        // voxel index of (x0,y0,z0) = (x0*Y + y0)*Z + z0 (may be "virtual" when padded)
  int sy, sx;      // voxel-index strides of y+1 and x+1
  unsigned valid;  // bit k set <=> corner k lies inside the grid
  float wx0, wx1, wy0, wy1, wz0, wz1;
  __device__ __forceinline__ bool ok(int k) const { return (valid >> k) & 1u; }
  __device__ __forceinline__ int off(int k) const { return base + (k >> 2) * sx + ((k >> 1) & 1) * sy + (k & 1); }
  __device__ __forceinline__ float w(int k) const {
    return fmul(fmul((k & 1) ? wz1 : wz0, ((k >> 1) & 1) ? wy1 : wy0), (k >> 2) ? wx1 : wx0);
  }
};
// from the continuous voxel coordinates (what tri_setup derives from the point): the per-survivor record march_fwd
// leaves for the k0 gather / scatter kernels, so that they do not redo the ray geometry (two IEEE divisions, a square
// root and ~200 more instructions per thread -- the bulk of those kernels' instruction count in round 2's ncu profile)
__device__ __forceinline__ Corner8 corner8_idx(const SceneDev& sc, float fx, float fy, float fz) {
  const Tri t = tri_from_index(fx, fy, fz);
  Corner8 c;
  c.sy = sc.Z;
  c.sx = sc.Y * sc.Z;
  c.base = (t.x0 * sc.Y + t.y0) * sc.Z + t.z0;
  c.wx0 = t.wx0; c.wx1 = t.wx1; c.wy0 = t.wy0; c.wy1 = t.wy1; c.wz0 = t.wz0; c.wz1 = t.wz1;
  const unsigned vx = (t.x0 >= 0 && t.x0 < sc.X ? 1u : 0u) | (t.x0 + 1 >= 0 && t.x0 + 1 < sc.X ? 2u : 0u);
  const unsigned vy = (t.y0 >= 0 && t.y0 < sc.Y ? 1u : 0u) | (t.y0 + 1 >= 0 && t.y0 + 1 < sc.Y ? 2u : 0u);
  const unsigned vz = (t.z0 >= 0 && t.z0 < sc.Z ? 1u : 0u) | (t.z0 + 1 >= 0 && t.z0 + 1 < sc.Z ? 2u : 0u);
  unsigned v = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (((vx >> (k >> 2)) & 1u) & ((vy >> ((k >> 1) & 1)) & 1u) & ((vz >> (k & 1)) & 1u)) v |= 1u << k;
  c.valid = v;
  return c;
}
__device__ __forceinline__ void voxel_coords(const SceneDev& sc, float px, float py, float pz, float& fx, float& fy,
                                             float& fz) {
  fx = unnorm_coord(px, sc.lo[0], sc.hi[0], sc.X);
  fy = unnorm_coord(py, sc.lo[1], sc.hi[1], sc.Y);
  fz = unnorm_coord(pz, sc.lo[2], sc.hi[2], sc.Z);
}
__device__ __forceinline__ Corner8 corner8(const SceneDev& sc, float px, float py, float pz) {
  float fx, fy, fz;
  voxel_coords(sc, px, py, pz, fx, fy, fz);
  return corner8_idx(sc, fx, fy, fz);
}

}  // namespace dvgo

// k0 channel counts (rgbnet_dim) the fused kernels are instantiated for; any other returns DVGO_EINVAL
#define DVGO_DISPATCH_C(Cval, ...)                       \
  switch (Cval) {                                        \
    case 3: { constexpr int kC = 3; __VA_ARGS__; } break;   \
    case 4: { constexpr int kC = 4; __VA_ARGS__; } break;   \
    case 6: { constexpr int kC = 6; __VA_ARGS__; } break;   \
    case 8: { constexpr int kC = 8; __VA_ARGS__; } break;   \
    case 9: { constexpr int kC = 9; __VA_ARGS__; } break;   \
    case 12: { constexpr int kC = 12; __VA_ARGS__; } break; \
    case 16: { constexpr int kC = 16; __VA_ARGS__; } break; \
    default: return DVGO_EINVAL;                         \
  }
