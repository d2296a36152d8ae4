// common.cuh -- shared helpers for the sm_100a kernels behind include/dvgo_b200.h.
// Only CUDA headers here (no torch): the kernel translation units compile in seconds.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dvgo_b200.h"

#define DVGO_API extern "C" __attribute__((visibility("default")))

namespace dvgo {

constexpr int kWarp = 32;
constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

static inline cudaStream_t as_stream(dvgo_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// Number of kernels this library has launched (bench.py reports it as `gpu_launches`).
extern unsigned long long g_launch_count;
static inline void count_launches(int n) { __atomic_fetch_add(&g_launch_count, n, __ATOMIC_RELAXED); }

// Launch-error check after the last of `n_kernels` launches of an entry point.
static inline int launch_status(int n_kernels = 1) {
  count_launches(n_kernels);
  return static_cast<int>(cudaGetLastError());
}

static inline int blocks_for(int64_t n, int threads) {
  return static_cast<int>((n + threads - 1) / threads);
}

// ---- exactly-rounded fp32 building blocks --------------------------------------------------------
// The reference kernels are compiled with nvcc's default -fmad=true, so which a*b+c become FFMA is
// a compiler decision (SURVEY.md appendix B lists what the SASS shows).  Integer-valued outputs
// (N_steps, mask_outbbox, maskcache) depend on those roundings, so we spell every operation with a
// round-to-nearest intrinsic: intrinsics are never re-contracted, whatever surrounds them.
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fma_(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

// Ray/AABB slab test, reference lib/cuda/render_utils_kernel.cu:23-33.
struct TMinMax { float t_min, t_max; };
__device__ __forceinline__ TMinMax slab_test(float ox, float oy, float oz, float dx, float dy,
                                             float dz, const float* __restrict__ xyz_min,
                                             const float* __restrict__ xyz_max, float near,
                                             float far) {
  const float vx = (dx == 0.f) ? 1e-6f : dx;  // (float)1e-6, :23-25
  const float vy = (dy == 0.f) ? 1e-6f : dy;
  const float vz = (dz == 0.f) ? 1e-6f : dz;
  const float ax = fdiv(fsub(xyz_max[0], ox), vx), bx = fdiv(fsub(xyz_min[0], ox), vx);
  const float ay = fdiv(fsub(xyz_max[1], oy), vy), by = fdiv(fsub(xyz_min[1], oy), vy);
  const float az = fdiv(fsub(xyz_max[2], oz), vz), bz = fdiv(fsub(xyz_min[2], oz), vz);
  TMinMax r;
  r.t_min = fmaxf(fminf(fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz)), far), near);
  r.t_max = fmaxf(fminf(fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz)), far), near);
  return r;
}

// Number of samples on a ray, reference :47: float subtract + IEEE divide + ceilf, max in double.
__device__ __forceinline__ int64_t n_samples_of(float t_min, float t_max, float stepdist) {
  const double c = static_cast<double>(ceilf(fdiv(fsub(t_max, t_min), stepdist)));
  return static_cast<int64_t>(fmax(c, 1.));
}

// Ray start point and unit direction, reference :62-71 (FMA contractions as in the SASS).
struct StartDir { float sx, sy, sz, ux, uy, uz; };
__device__ __forceinline__ StartDir ray_start_dir(float ox, float oy, float oz, float dx, float dy,
                                                  float dz, float t_min) {
  // SASS of the reference: FMUL(dy,dy); FFMA(dx,dx,.); FFMA(dz,dz,.)
  const float rnorm = sqrtf(fma_(dz, dz, fma_(dx, dx, fmul(dy, dy))));
  StartDir r;
  r.sx = fma_(dx, t_min, ox);
  r.sy = fma_(dy, t_min, oy);
  r.sz = fma_(dz, t_min, oz);
  r.ux = fdiv(dx, rnorm);
  r.uy = fdiv(dy, rnorm);
  r.uz = fdiv(dz, rnorm);
  return r;
}

__device__ __forceinline__ bool out_of_bbox(float px, float py, float pz,
                                            const float* __restrict__ lo,
                                            const float* __restrict__ hi) {
  return (lo[0] > px) | (lo[1] > py) | (lo[2] > pz) | (hi[0] < px) | (hi[1] < py) | (hi[2] < pz);
}

// Trilinear corner geometry following ATen's grid_sampler_3d (align_corners=True, zero padding)
// applied to DirectVoxGO's ind_norm (lib/dvgo.py:316): see oracle/dvgo_oracle.c tri_setup.
struct Tri {
  int x0, y0, z0;
  float wx0, wx1, wy0, wy1, wz0, wz1;
};
__device__ __forceinline__ float unnorm_coord(float x, float lo, float hi, int size) {
  const float u = fdiv(fsub(x, lo), fsub(hi, lo));
  const float n = fsub(fmul(u, 2.f), 1.f);
  return fmul(fmul(fadd(n, 1.f), 0.5f), static_cast<float>(size - 1));
}
__device__ __forceinline__ Tri tri_from_index(float fx, float fy, float fz);
__device__ __forceinline__ Tri tri_setup(float px, float py, float pz,
                                         const float* __restrict__ lo,
                                         const float* __restrict__ hi, int X, int Y, int Z) {
  return tri_from_index(unnorm_coord(px, lo[0], hi[0], X), unnorm_coord(py, lo[1], hi[1], Y),
                        unnorm_coord(pz, lo[2], hi[2], Z));
}
// Same, for coordinates that are ALREADY normalised to [-1, 1] (the `grid` argument of F.grid_sample itself): only
// ATen's align_corners=True un-normalisation ((n + 1) / 2) * (size - 1) is applied.
__device__ __forceinline__ float unnorm_only(float n, int size) {
  return fmul(fmul(fadd(n, 1.f), 0.5f), static_cast<float>(size - 1));
}
__device__ __forceinline__ Tri tri_from_index(float fx, float fy, float fz) {
  Tri t;
  const float x0f = floorf(fx), y0f = floorf(fy), z0f = floorf(fz);
  t.x0 = static_cast<int>(x0f);
  t.y0 = static_cast<int>(y0f);
  t.z0 = static_cast<int>(z0f);
  t.wx0 = fsub(x0f + 1.f, fx); t.wx1 = fsub(fx, x0f);
  t.wy0 = fsub(y0f + 1.f, fy); t.wy1 = fsub(fy, y0f);
  t.wz0 = fsub(z0f + 1.f, fz); t.wz1 = fsub(fz, z0f);
  return t;
}

}  // namespace dvgo
