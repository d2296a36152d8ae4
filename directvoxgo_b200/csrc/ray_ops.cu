// ray_ops.cu -- ray/AABB clipping, step generation, point sampling and occupancy lookup
// (rows a1-a6 of SURVEY.md section 8).  Reference: lib/cuda/render_utils_kernel.cu:12-351.
//
// Design notes (B200): these are latency/launch-bound at 8192 rays, so the win over the reference
// is structural: its ~12 launches + 2 ATen scans + memset for sample_pts_on_rays become
// 2 launches (setup+scan, warp-per-ray fill); every store of the [M,...] outputs is coalesced
// because a warp owns a contiguous run of samples of one ray.
#include "common.cuh"
#include "scan.cuh"

namespace dvgo {

// ---- a1: infer_t_minmax --------------------------------------------------------------------------
__global__ void __launch_bounds__(256) infer_t_minmax_kernel(
    const float* __restrict__ rays_o, const float* __restrict__ rays_d,
    const float* __restrict__ xyz_min, const float* __restrict__ xyz_max, float near, float far,
    int n_rays, float* __restrict__ t_min, float* __restrict__ t_max) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rays) return;
  const TMinMax t = slab_test(rays_o[3 * r], rays_o[3 * r + 1], rays_o[3 * r + 2], rays_d[3 * r],
                              rays_d[3 * r + 1], rays_d[3 * r + 2], xyz_min, xyz_max, near, far);
  t_min[r] = t.t_min;
  t_max[r] = t.t_max;
}

// ---- a2: infer_n_samples -------------------------------------------------------------------------
__global__ void __launch_bounds__(256) infer_n_samples_kernel(const float* __restrict__ t_min,
                                                              const float* __restrict__ t_max,
                                                              float stepdist, int n_rays,
                                                              int64_t* __restrict__ n_samples) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n_rays) n_samples[r] = n_samples_of(t_min[r], t_max[r], stepdist);
}

// ---- a3: infer_ray_start_dir ---------------------------------------------------------------------
__global__ void __launch_bounds__(256) infer_ray_start_dir_kernel(
    const float* __restrict__ rays_o, const float* __restrict__ rays_d,
    const float* __restrict__ t_min, int n_rays, float* __restrict__ rays_start,
    float* __restrict__ rays_dir) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rays) return;
  const StartDir s = ray_start_dir(rays_o[3 * r], rays_o[3 * r + 1], rays_o[3 * r + 2],
                                   rays_d[3 * r], rays_d[3 * r + 1], rays_d[3 * r + 2], t_min[r]);
  rays_start[3 * r] = s.sx; rays_start[3 * r + 1] = s.sy; rays_start[3 * r + 2] = s.sz;
  rays_dir[3 * r] = s.ux;   rays_dir[3 * r + 1] = s.uy;   rays_dir[3 * r + 2] = s.uz;
}

// ---- a4 phase 1: per-ray clip + step count, then an inclusive scan ---------------------------------
__global__ void __launch_bounds__(256) ray_count_kernel(
    const float* __restrict__ rays_o, const float* __restrict__ rays_d,
    const float* __restrict__ xyz_min, const float* __restrict__ xyz_max, float near, float far,
    float stepdist, int n_rays, float* __restrict__ t_min, float* __restrict__ t_max,
    int64_t* __restrict__ N_steps) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rays) return;
  const TMinMax t = slab_test(rays_o[3 * r], rays_o[3 * r + 1], rays_o[3 * r + 2], rays_d[3 * r],
                              rays_d[3 * r + 1], rays_d[3 * r + 2], xyz_min, xyz_max, near, far);
  t_min[r] = t.t_min;
  t_max[r] = t.t_max;
  N_steps[r] = n_samples_of(t.t_min, t.t_max, stepdist);
}

// (the single-CTA inclusive scan of the int64 counts lives in scan.cuh, shared with f64_ops.cu)

// ---- a4 phase 2: one warp per ray writes that ray's contiguous run of samples -----------------------
__global__ void __launch_bounds__(256) sample_pts_fill_kernel(
    const float* __restrict__ rays_o, const float* __restrict__ rays_d,
    const float* __restrict__ xyz_min, const float* __restrict__ xyz_max,
    const float* __restrict__ t_min, const int64_t* __restrict__ cumsum, float stepdist, int n_rays,
    float* __restrict__ rays_pts, uint8_t* __restrict__ mask_outbbox, int64_t* __restrict__ ray_id,
    int64_t* __restrict__ step_id) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < n_rays;
       r += gridDim.x * warps_per_block) {
    const int64_t end = cumsum[r];
    const int64_t begin = r ? cumsum[r - 1] : 0;
    const int n = static_cast<int>(end - begin);
    const StartDir s = ray_start_dir(rays_o[3 * r], rays_o[3 * r + 1], rays_o[3 * r + 2],
                                     rays_d[3 * r], rays_d[3 * r + 1], rays_d[3 * r + 2], t_min[r]);
    for (int i = lane; i < n; i += 32) {
      const float dist = fmul(stepdist, static_cast<float>(i));  // :178
      const float px = fma_(s.ux, dist, s.sx);                   // :179-181 (FFMA in the SASS)
      const float py = fma_(s.uy, dist, s.sy);
      const float pz = fma_(s.uz, dist, s.sz);
      const int64_t idx = begin + i;
      rays_pts[3 * idx] = px;
      rays_pts[3 * idx + 1] = py;
      rays_pts[3 * idx + 2] = pz;
      mask_outbbox[idx] = out_of_bbox(px, py, pz, xyz_min, xyz_max);
      ray_id[idx] = r;
      step_id[idx] = i;
    }
  }
}

// ---- a5: sample_ndc_pts_on_rays ------------------------------------------------------------------
__global__ void __launch_bounds__(256) sample_ndc_kernel(
    const float* __restrict__ rays_o, const float* __restrict__ rays_d,
    const float* __restrict__ xyz_min, const float* __restrict__ xyz_max, int N_samples,
    int64_t total, float* __restrict__ rays_pts, uint8_t* __restrict__ mask_outbbox) {
  const float denom = static_cast<float>(N_samples - 1);
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(idx / N_samples);
    const int s = static_cast<int>(idx - static_cast<int64_t>(r) * N_samples);
    const float dist = fdiv(static_cast<float>(s), denom);  // :254
    const float px = fma_(rays_d[3 * r], dist, rays_o[3 * r]);
    const float py = fma_(rays_d[3 * r + 1], dist, rays_o[3 * r + 1]);
    const float pz = fma_(rays_d[3 * r + 2], dist, rays_o[3 * r + 2]);
    rays_pts[3 * idx] = px;
    rays_pts[3 * idx + 1] = py;
    rays_pts[3 * idx + 2] = pz;
    mask_outbbox[idx] = out_of_bbox(px, py, pz, xyz_min, xyz_max);
  }
}

// ---- a6: maskcache_lookup ------------------------------------------------------------------------
__device__ __forceinline__ bool mask_lookup(const uint8_t* __restrict__ world, float x, float y,
                                            float z, const float* __restrict__ scale,
                                            const float* __restrict__ shift, int sz_i, int sz_j,
                                            int sz_k) {
  // :312-314: FFMA then round() (half away from zero) then int conversion.
  const int i = static_cast<int>(roundf(fma_(x, scale[0], shift[0])));
  const int j = static_cast<int>(roundf(fma_(y, scale[1], shift[1])));
  const int k = static_cast<int>(roundf(fma_(z, scale[2], shift[2])));
  if ((0 <= i) & (i < sz_i) & (0 <= j) & (j < sz_j) & (0 <= k) & (k < sz_k))
    return world[(static_cast<int64_t>(i) * sz_j + j) * sz_k + k] != 0;
  return false;
}

__global__ void __launch_bounds__(256) maskcache_lookup_kernel(
    const uint8_t* __restrict__ world, const float* __restrict__ xyz,
    const float* __restrict__ scale, const float* __restrict__ shift, int sz_i, int sz_j, int sz_k,
    int64_t n_pts, uint8_t* __restrict__ out) {
  for (int64_t p = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; p < n_pts;
       p += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    out[p] = mask_lookup(world, xyz[3 * p], xyz[3 * p + 1], xyz[3 * p + 2], scale, shift, sz_i,
                         sz_j, sz_k);
  }
}

static inline int grid_for(int64_t n, int threads) {
  // Grid-stride kernels: cap at 32 resident CTAs' worth per SM so huge n does not over-launch.
  const int64_t want = (n + threads - 1) / threads;
  const int64_t cap = static_cast<int64_t>(kNumSMs) * 32;
  return static_cast<int>(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace dvgo

using namespace dvgo;

namespace dvgo { unsigned long long g_launch_count = 0; }

DVGO_API int dvgo_abi_version(void) { return DVGO_ABI_VERSION; }
DVGO_API unsigned long long dvgo_launch_count(void) {
  return __atomic_load_n(&dvgo::g_launch_count, __ATOMIC_RELAXED);
}
DVGO_API const char* dvgo_build_arch(void) { return "sm_100a"; }

DVGO_API int dvgo_infer_t_minmax(const float* rays_o, const float* rays_d, const float* xyz_min,
                                 const float* xyz_max, float near, float far, int n_rays,
                                 float* t_min, float* t_max, dvgo_stream_t stream) {
  if (n_rays < 0) return DVGO_EINVAL;
  if (n_rays == 0) return 0;
  if (!rays_o || !rays_d || !xyz_min || !xyz_max || !t_min || !t_max) return DVGO_EINVAL;
  infer_t_minmax_kernel<<<blocks_for(n_rays, 256), 256, 0, as_stream(stream)>>>(
      rays_o, rays_d, xyz_min, xyz_max, near, far, n_rays, t_min, t_max);
  return launch_status();
}

DVGO_API int dvgo_infer_n_samples(const float* t_min, const float* t_max, float stepdist,
                                  int n_rays, int64_t* n_samples, dvgo_stream_t stream) {
  if (n_rays < 0) return DVGO_EINVAL;
  if (n_rays == 0) return 0;
  if (!t_min || !t_max || !n_samples) return DVGO_EINVAL;
  infer_n_samples_kernel<<<blocks_for(n_rays, 256), 256, 0, as_stream(stream)>>>(
      t_min, t_max, stepdist, n_rays, n_samples);
  return launch_status();
}

DVGO_API int dvgo_infer_ray_start_dir(const float* rays_o, const float* rays_d, const float* t_min,
                                      int n_rays, float* rays_start, float* rays_dir,
                                      dvgo_stream_t stream) {
  if (n_rays < 0) return DVGO_EINVAL;
  if (n_rays == 0) return 0;
  if (!rays_o || !rays_d || !t_min || !rays_start || !rays_dir) return DVGO_EINVAL;
  infer_ray_start_dir_kernel<<<blocks_for(n_rays, 256), 256, 0, as_stream(stream)>>>(
      rays_o, rays_d, t_min, n_rays, rays_start, rays_dir);
  return launch_status();
}

DVGO_API int dvgo_sample_pts_count(const float* rays_o, const float* rays_d, const float* xyz_min,
                                   const float* xyz_max, float near, float far, float stepdist,
                                   int n_rays, float* t_min, float* t_max, int64_t* N_steps,
                                   int64_t* N_steps_cumsum, int64_t* total_host,
                                   dvgo_stream_t stream) {
  if (n_rays < 0 || !total_host) return DVGO_EINVAL;
  *total_host = 0;
  if (n_rays == 0) return 0;
  if (!rays_o || !rays_d || !xyz_min || !xyz_max || !t_min || !t_max || !N_steps || !N_steps_cumsum)
    return DVGO_EINVAL;
  cudaStream_t s = as_stream(stream);
  ray_count_kernel<<<blocks_for(n_rays, 256), 256, 0, s>>>(rays_o, rays_d, xyz_min, xyz_max, near,
                                                           far, stepdist, n_rays, t_min, t_max,
                                                           N_steps);
  inclusive_scan_i64_kernel<<<1, kScanThreads, 0, s>>>(N_steps, n_rays, N_steps_cumsum);
  int err = launch_status(2);
  if (err) return err;
  // The reference's single host sync (render_utils_kernel.cu:206): the caller must size outputs.
  err = static_cast<int>(cudaMemcpyAsync(total_host, N_steps_cumsum + (n_rays - 1), sizeof(int64_t),
                                         cudaMemcpyDeviceToHost, s));
  if (err) return err;
  return static_cast<int>(cudaStreamSynchronize(s));
}

DVGO_API int dvgo_sample_pts_fill(const float* rays_o, const float* rays_d, const float* xyz_min,
                                  const float* xyz_max, const float* t_min,
                                  const int64_t* N_steps_cumsum, float stepdist, int n_rays,
                                  int64_t total, float* rays_pts, uint8_t* mask_outbbox,
                                  int64_t* ray_id, int64_t* step_id, dvgo_stream_t stream) {
  if (n_rays < 0 || total < 0 || total >= (int64_t(1) << 31)) return DVGO_EINVAL;
  if (n_rays == 0 || total == 0) return 0;
  if (!rays_o || !rays_d || !xyz_min || !xyz_max || !t_min || !N_steps_cumsum || !rays_pts ||
      !mask_outbbox || !ray_id || !step_id)
    return DVGO_EINVAL;
  const int warps_per_block = 8;
  const int64_t want = (static_cast<int64_t>(n_rays) + warps_per_block - 1) / warps_per_block;
  const int blocks = static_cast<int>(want < kNumSMs * 16 ? want : kNumSMs * 16);
  sample_pts_fill_kernel<<<blocks, warps_per_block * 32, 0, as_stream(stream)>>>(
      rays_o, rays_d, xyz_min, xyz_max, t_min, N_steps_cumsum, stepdist, n_rays, rays_pts,
      mask_outbbox, ray_id, step_id);
  return launch_status();
}

DVGO_API int dvgo_sample_ndc_pts_on_rays(const float* rays_o, const float* rays_d,
                                         const float* xyz_min, const float* xyz_max, int N_samples,
                                         int n_rays, float* rays_pts, uint8_t* mask_outbbox,
                                         dvgo_stream_t stream) {
  if (n_rays < 0 || N_samples < 0) return DVGO_EINVAL;
  const int64_t total = static_cast<int64_t>(n_rays) * N_samples;
  if (total == 0) return 0;
  if (!rays_o || !rays_d || !xyz_min || !xyz_max || !rays_pts || !mask_outbbox) return DVGO_EINVAL;
  sample_ndc_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(
      rays_o, rays_d, xyz_min, xyz_max, N_samples, total, rays_pts, mask_outbbox);
  return launch_status();
}

DVGO_API int dvgo_maskcache_lookup(const uint8_t* world, const float* xyz,
                                   const float* xyz2ijk_scale, const float* xyz2ijk_shift, int sz_i,
                                   int sz_j, int sz_k, int64_t n_pts, uint8_t* out,
                                   dvgo_stream_t stream) {
  if (n_pts < 0 || sz_i < 0 || sz_j < 0 || sz_k < 0) return DVGO_EINVAL;
  if (n_pts == 0) return 0;  // reference short-circuit, render_utils_kernel.cu:333-335
  if (!world || !xyz || !xyz2ijk_scale || !xyz2ijk_shift || !out) return DVGO_EINVAL;
  maskcache_lookup_kernel<<<grid_for(n_pts, 256), 256, 0, as_stream(stream)>>>(
      world, xyz, xyz2ijk_scale, xyz2ijk_shift, sz_i, sz_j, sz_k, n_pts, out);
  return launch_status();
}
