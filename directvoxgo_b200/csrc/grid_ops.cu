// grid_ops.cu -- DenseGrid trilinear sampling in the reference's NCDHW layout (row a7) and the
// segment_coo compositing reduction (row a10).
//
// Reference: lib/dvgo.py:312-328 -> ATen grid_sampler_3d (bilinear, align_corners=True, zero pad)
// and torch_scatter.segment_coo at lib/dvgo.py:554-558,571-575.
//
// These are the drop-in, layout-compatible ops.  The fused trainer (fused_*.cu) keeps its own
// channel-last copy of k0 so one corner fetch is a single 48-byte vector access; here every
// channel lives in its own X*Y*Z plane (NCDHW), so we at least (i) do the ind_norm arithmetic of
// dvgo.py:316 in-register instead of five elementwise launches, (ii) compute corner offsets and
// weights once per point for all channels, (iii) write the [P,C] result directly instead of the
// reference's [C,P] -> transpose copy (dvgo.py:321).
#include "common.cuh"

namespace dvgo {

struct Corners {
  int64_t off[8];
  float w[8];
};

// Corner order and weight association follow ATen (see oracle/dvgo_oracle.c tri_setup).
__device__ __forceinline__ Corners corners_of(const Tri& t, int X, int Y, int Z) {
  Corners c;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int dx = k >> 2, dy = (k >> 1) & 1, dz = k & 1;
    const int xi = t.x0 + dx, yi = t.y0 + dy, zi = t.z0 + dz;
    const bool in = (xi >= 0) & (xi < X) & (yi >= 0) & (yi < Y) & (zi >= 0) & (zi < Z);
    const float w = fmul(fmul(dz ? t.wz1 : t.wz0, dy ? t.wy1 : t.wy0), dx ? t.wx1 : t.wx0);
    c.off[k] = in ? (static_cast<int64_t>(xi) * Y + yi) * Z + zi : -1;
    c.w[k] = w;
  }
  return c;
}

// NORM: `xyz` holds F.grid_sample's own normalised (w, h, d) = (z, y, x) coordinates instead of world points.
template <bool NORM>
__device__ __forceinline__ Tri tri_of(const float* __restrict__ xyz, int64_t p, const float* __restrict__ lo,
                                      const float* __restrict__ hi, int X, int Y, int Z) {
  if (NORM)
    return tri_from_index(unnorm_only(xyz[3 * p + 2], X), unnorm_only(xyz[3 * p + 1], Y), unnorm_only(xyz[3 * p], Z));
  return tri_setup(xyz[3 * p], xyz[3 * p + 1], xyz[3 * p + 2], lo, hi, X, Y, Z);
}

template <bool NORM>
__global__ void __launch_bounds__(256) grid_sample_3d_kernel(
    const float* __restrict__ grid, int C, int X, int Y, int Z, const float* __restrict__ xyz,
    const float* __restrict__ xyz_min, const float* __restrict__ xyz_max, int64_t n_pts,
    float* __restrict__ out) {
  const int64_t plane = static_cast<int64_t>(X) * Y * Z;
  for (int64_t p = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; p < n_pts;
       p += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const Tri t = tri_of<NORM>(xyz, p, xyz_min, xyz_max, X, Y, Z);
    const Corners cn = corners_of(t, X, Y, Z);
    for (int c = 0; c < C; ++c) {
      const float* __restrict__ g = grid + c * plane;
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = cn.off[k] >= 0 ? __ldg(g + cn.off[k]) : 0.f;
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (cn.off[k] >= 0) acc = fma_(v[k], cn.w[k], acc);  // ATen: out_acc += v * w
      out[p * C + c] = acc;
    }
  }
}

template <bool NORM>
__global__ void __launch_bounds__(256) grid_sample_3d_backward_kernel(
    const float* __restrict__ grad_out, int C, int X, int Y, int Z, const float* __restrict__ xyz,
    const float* __restrict__ xyz_min, const float* __restrict__ xyz_max, int64_t n_pts,
    float* __restrict__ grad_grid) {
  const int64_t plane = static_cast<int64_t>(X) * Y * Z;
  for (int64_t p = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; p < n_pts;
       p += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const Tri t = tri_of<NORM>(xyz, p, xyz_min, xyz_max, X, Y, Z);
    const Corners cn = corners_of(t, X, Y, Z);
    for (int c = 0; c < C; ++c) {
      const float g = grad_out[p * C + c];
      float* __restrict__ gg = grad_grid + c * plane;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (cn.off[k] >= 0) atomicAdd(gg + cn.off[k], fmul(cn.w[k], g));
    }
  }
}

// ---- a10: segment_coo (sum) ----------------------------------------------------------------------
// index is sorted: each warp reduces runs of equal indices among its 32 consecutive points with a
// segmented shuffle reduction and only the head of each run touches memory (one atomicAdd per
// run per warp instead of one per point).
__global__ void __launch_bounds__(256) segment_coo_sum_kernel(const float* __restrict__ src,
                                                              const int64_t* __restrict__ index,
                                                              int64_t n_pts, int D,
                                                              float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t n_round = (n_pts + 31) / 32 * 32;
  for (int64_t p = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; p < n_round;
       p += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const bool valid = p < n_pts;
    const int64_t key = valid ? index[p] : -1;
    const int64_t key_prev = __shfl_up_sync(0xffffffffu, key, 1);
    const bool head_any = lane == 0 || key_prev != key;
    const bool head = valid && head_any;
    // runs are identified by position (ballot of run heads), not by key equality at a shuffle distance, so an
    // unsorted index (A.. B.. A..) degrades to more atomics instead of double-counting
    const unsigned heads = __ballot_sync(0xffffffffu, head_any);
    const int seg = __popc(heads & (0xffffffffu >> (31 - lane)));
    for (int d = 0; d < D; ++d) {
      float v = valid ? src[p * D + d] : 0.f;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const float vd = __shfl_down_sync(0xffffffffu, v, off);
        const int sd = __shfl_down_sync(0xffffffffu, seg, off);
        if (lane + off < 32 && sd == seg) v += vd;
      }
      if (head) atomicAdd(out + key * D + d, v);
    }
  }
}

__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ table,
                                                          const int64_t* __restrict__ index,
                                                          int64_t n_pts, int D,
                                                          float* __restrict__ out) {
  const int64_t total = n_pts * D;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t p = i / D;
    const int d = static_cast<int>(i - p * D);
    out[i] = table[index[p] * D + d];
  }
}

static inline int grid_for(int64_t n, int threads) {
  const int64_t want = (n + threads - 1) / threads;
  const int64_t cap = static_cast<int64_t>(kNumSMs) * 32;
  return static_cast<int>(want < cap ? (want > 0 ? want : 1) : cap);
}

// ---- 2-D tri-plane sampling (lib/tri_dvgo.py:456-464): ATen grid_sampler_2d, bilinear, align_corners, zero pad --
struct Bil { int x0, y0; float w[4]; };   // corner order nw, ne, sw, se as ATen accumulates them
template <bool NORM>   // NORM: p3 -> this point's (w, h) pair of F.grid_sample's own normalised coordinates
__device__ __forceinline__ Bil bil_setup(const float* __restrict__ p3, const float* __restrict__ lo,
                                         const float* __restrict__ hi, int axis_w, int axis_h, int W, int H) {
  const float ix = NORM ? unnorm_only(p3[0], W) : unnorm_coord(p3[axis_w], lo[axis_w], hi[axis_w], W);
  const float iy = NORM ? unnorm_only(p3[1], H) : unnorm_coord(p3[axis_h], lo[axis_h], hi[axis_h], H);
  const float x0f = floorf(ix), y0f = floorf(iy);
  Bil b;
  b.x0 = static_cast<int>(x0f);
  b.y0 = static_cast<int>(y0f);
  const float ex = fsub(x0f + 1.f, ix), wx = fsub(ix, x0f);   // (ix_se - ix), (ix - ix_nw)
  const float ey = fsub(y0f + 1.f, iy), wy = fsub(iy, y0f);
  b.w[0] = fmul(ex, ey);   // nw
  b.w[1] = fmul(wx, ey);   // ne
  b.w[2] = fmul(ex, wy);   // sw
  b.w[3] = fmul(wx, wy);   // se
  return b;
}

template <bool NORM>
__global__ void __launch_bounds__(256) grid_sample_2d_kernel(
    const float* __restrict__ plane, int C, int H, int W, const float* __restrict__ xyz,
    const float* __restrict__ xyz_min, const float* __restrict__ xyz_max, int axis_w, int axis_h, int64_t n_pts,
    float* __restrict__ out) {
  const int64_t hw = static_cast<int64_t>(H) * W;
  for (int64_t p = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; p < n_pts;
       p += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const Bil b = bil_setup<NORM>(xyz + (NORM ? 2 : 3) * p, xyz_min, xyz_max, axis_w, axis_h, W, H);
    int64_t off[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int x = b.x0 + (k & 1), y = b.y0 + (k >> 1);
      off[k] = (x >= 0 && x < W && y >= 0 && y < H) ? static_cast<int64_t>(y) * W + x : -1;
    }
    for (int c = 0; c < C; ++c) {
      const float* __restrict__ g = plane + c * hw;
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (off[k] >= 0) acc = fma_(__ldg(g + off[k]), b.w[k], acc);
      out[p * C + c] = acc;
    }
  }
}

template <bool NORM>
__global__ void __launch_bounds__(256) grid_sample_2d_backward_kernel(
    const float* __restrict__ grad_out, int C, int H, int W, const float* __restrict__ xyz,
    const float* __restrict__ xyz_min, const float* __restrict__ xyz_max, int axis_w, int axis_h, int64_t n_pts,
    float* __restrict__ grad_plane) {
  const int64_t hw = static_cast<int64_t>(H) * W;
  for (int64_t p = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; p < n_pts;
       p += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const Bil b = bil_setup<NORM>(xyz + (NORM ? 2 : 3) * p, xyz_min, xyz_max, axis_w, axis_h, W, H);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int x = b.x0 + (k & 1), y = b.y0 + (k >> 1);
      if (x < 0 || x >= W || y < 0 || y >= H) continue;
      const int64_t off = static_cast<int64_t>(y) * W + x;
      for (int c = 0; c < C; ++c) atomicAdd(grad_plane + c * hw + off, fmul(b.w[k], grad_out[p * C + c]));
    }
  }
}

}  // namespace dvgo

using namespace dvgo;

DVGO_API int dvgo_grid_sample_3d(const float* grid, int C, int X, int Y, int Z, const float* xyz,
                                 const float* xyz_min, const float* xyz_max, int64_t n_pts,
                                 float* out, dvgo_stream_t stream) {
  if (n_pts < 0 || C < 0 || X <= 0 || Y <= 0 || Z <= 0) return DVGO_EINVAL;
  if (n_pts == 0 || C == 0) return 0;
  if (!grid || !xyz || !xyz_min || !xyz_max || !out) return DVGO_EINVAL;
  grid_sample_3d_kernel<false><<<grid_for(n_pts, 256), 256, 0, as_stream(stream)>>>(
      grid, C, X, Y, Z, xyz, xyz_min, xyz_max, n_pts, out);
  return launch_status();
}

DVGO_API int dvgo_grid_sample_3d_norm(const float* grid, int C, int X, int Y, int Z, const float* ind_norm,
                                      int64_t n_pts, float* out, dvgo_stream_t stream) {
  if (n_pts < 0 || C < 0 || X <= 0 || Y <= 0 || Z <= 0) return DVGO_EINVAL;
  if (n_pts == 0 || C == 0) return 0;
  if (!grid || !ind_norm || !out) return DVGO_EINVAL;
  grid_sample_3d_kernel<true><<<grid_for(n_pts, 256), 256, 0, as_stream(stream)>>>(
      grid, C, X, Y, Z, ind_norm, nullptr, nullptr, n_pts, out);
  return launch_status();
}

DVGO_API int dvgo_grid_sample_3d_norm_backward(const float* grad_out, int C, int X, int Y, int Z,
                                               const float* ind_norm, int64_t n_pts, float* grad_grid,
                                               dvgo_stream_t stream) {
  if (n_pts < 0 || C < 0 || X <= 0 || Y <= 0 || Z <= 0) return DVGO_EINVAL;
  if (n_pts == 0 || C == 0) return 0;
  if (!grad_out || !ind_norm || !grad_grid) return DVGO_EINVAL;
  grid_sample_3d_backward_kernel<true><<<grid_for(n_pts, 256), 256, 0, as_stream(stream)>>>(
      grad_out, C, X, Y, Z, ind_norm, nullptr, nullptr, n_pts, grad_grid);
  return launch_status();
}

DVGO_API int dvgo_grid_sample_3d_backward(const float* grad_out, int C, int X, int Y, int Z,
                                          const float* xyz, const float* xyz_min,
                                          const float* xyz_max, int64_t n_pts, float* grad_grid,
                                          dvgo_stream_t stream) {
  if (n_pts < 0 || C < 0 || X <= 0 || Y <= 0 || Z <= 0) return DVGO_EINVAL;
  if (n_pts == 0 || C == 0) return 0;
  if (!grad_out || !xyz || !xyz_min || !xyz_max || !grad_grid) return DVGO_EINVAL;
  grid_sample_3d_backward_kernel<false><<<grid_for(n_pts, 256), 256, 0, as_stream(stream)>>>(
      grad_out, C, X, Y, Z, xyz, xyz_min, xyz_max, n_pts, grad_grid);
  return launch_status();
}

DVGO_API int dvgo_segment_coo_sum(const float* src, const int64_t* index, int64_t n_pts, int D,
                                  int64_t n_seg, float* out, dvgo_stream_t stream) {
  if (n_pts < 0 || D < 0 || n_seg < 0) return DVGO_EINVAL;
  if (n_pts == 0 || D == 0) return 0;
  if (!src || !index || !out) return DVGO_EINVAL;
  segment_coo_sum_kernel<<<grid_for(n_pts, 256), 256, 0, as_stream(stream)>>>(src, index, n_pts, D,
                                                                              out);
  return launch_status();
}

DVGO_API int dvgo_gather_rows(const float* table, const int64_t* index, int64_t n_pts, int D,
                              float* out, dvgo_stream_t stream) {
  if (n_pts < 0 || D < 0) return DVGO_EINVAL;
  if (n_pts == 0 || D == 0) return 0;
  if (!table || !index || !out) return DVGO_EINVAL;
  gather_rows_kernel<<<grid_for(n_pts * D, 256), 256, 0, as_stream(stream)>>>(table, index, n_pts,
                                                                              D, out);
  return launch_status();
}

DVGO_API int dvgo_grid_sample_2d(const float* plane, int C, int H, int W, const float* xyz, const float* xyz_min,
                                 const float* xyz_max, int axis_w, int axis_h, int64_t n_pts, float* out,
                                 dvgo_stream_t stream) {
  if (C <= 0 || H <= 0 || W <= 0 || n_pts < 0 || axis_w < 0 || axis_w > 2 || axis_h < 0 || axis_h > 2) return DVGO_EINVAL;
  if (n_pts == 0) return 0;
  if (!plane || !xyz || !xyz_min || !xyz_max || !out) return DVGO_EINVAL;
  grid_sample_2d_kernel<false><<<grid_for(n_pts, 256), 256, 0, as_stream(stream)>>>(plane, C, H, W, xyz, xyz_min,
                                                                                    xyz_max, axis_w, axis_h, n_pts, out);
  return launch_status();
}

DVGO_API int dvgo_grid_sample_2d_norm(const float* plane, int C, int H, int W, const float* ind_norm, int64_t n_pts,
                                      float* out, dvgo_stream_t stream) {
  if (C <= 0 || H <= 0 || W <= 0 || n_pts < 0) return DVGO_EINVAL;
  if (n_pts == 0) return 0;
  if (!plane || !ind_norm || !out) return DVGO_EINVAL;
  grid_sample_2d_kernel<true><<<grid_for(n_pts, 256), 256, 0, as_stream(stream)>>>(plane, C, H, W, ind_norm, nullptr,
                                                                                   nullptr, 0, 1, n_pts, out);
  return launch_status();
}

DVGO_API int dvgo_grid_sample_2d_norm_backward(const float* grad_out, int C, int H, int W, const float* ind_norm,
                                               int64_t n_pts, float* grad_plane, dvgo_stream_t stream) {
  if (C <= 0 || H <= 0 || W <= 0 || n_pts < 0) return DVGO_EINVAL;
  if (n_pts == 0) return 0;
  if (!grad_out || !ind_norm || !grad_plane) return DVGO_EINVAL;
  grid_sample_2d_backward_kernel<true><<<grid_for(n_pts, 256), 256, 0, as_stream(stream)>>>(
      grad_out, C, H, W, ind_norm, nullptr, nullptr, 0, 1, n_pts, grad_plane);
  return launch_status();
}

DVGO_API int dvgo_grid_sample_2d_backward(const float* grad_out, int C, int H, int W, const float* xyz,
                                          const float* xyz_min, const float* xyz_max, int axis_w, int axis_h,
                                          int64_t n_pts, float* grad_plane, dvgo_stream_t stream) {
  if (C <= 0 || H <= 0 || W <= 0 || n_pts < 0 || axis_w < 0 || axis_w > 2 || axis_h < 0 || axis_h > 2) return DVGO_EINVAL;
  if (n_pts == 0) return 0;
  if (!grad_out || !xyz || !xyz_min || !xyz_max || !grad_plane) return DVGO_EINVAL;
  grid_sample_2d_backward_kernel<false><<<grid_for(n_pts, 256), 256, 0, as_stream(stream)>>>(
      grad_out, C, H, W, xyz, xyz_min, xyz_max, axis_w, axis_h, n_pts, grad_plane);
  return launch_status();
}
