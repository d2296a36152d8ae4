/*
 * dvgo_b200.h -- C ABI of libdvgo_b200.so: the B200 (sm_100a) implementation of DirectVoxGO's
 * per-ray volume-rendering + grid-optimisation hot path.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - tensors are dense, row-major, fp32 / int64 / uint8(bool) exactly as the reference's torch
 *     tensors are (lib/cuda/render_utils.cpp:40-42 CHECK_CUDA + CHECK_CONTIGUOUS); the float64
 *     instantiation the reference's AT_DISPATCH_FLOATING_TYPES also provides is declared in
 *     dvgo_b200_f64.h (same entry points with an _f64 suffix and double* tensors);
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream, what the reference
 *     uses, lib/cuda/render_utils_kernel.cu:87); all work is enqueued asynchronously on it, no
 *     entry point synchronises the device or allocates memory unless it says so;
 *   - the CUDA device must be current (the torch wrapper guards the device of the inputs);
 *   - return value: 0 on success, a positive cudaError_t if a launch failed, DVGO_EINVAL (-1) for
 *     an invalid argument.  No exceptions cross the boundary.
 *
 * Each entry point names the reference interface it replaces (path:line under hbell99/DirectVoxGO).
 * The reference binds that interface through pybind11 (lib/cuda/render_utils.cpp:144-155,
 * lib/cuda/total_variation.cpp:22-24, lib/cuda/adam_upd.cpp:79-86); INTEGRATION.md shows the
 * binding a maintainer adds to call this ABI instead.
 */
#ifndef DVGO_B200_H_
#define DVGO_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DVGO_EINVAL (-1)
/* 2: dvgo_scene_t (dvgo_b200_fused.h) gained `exact_transmittance`; the float64 entry points (dvgo_b200_f64.h) were added */
#define DVGO_ABI_VERSION 2

typedef void* dvgo_stream_t; /* cudaStream_t */

/* ABI version, for the binding to check at load time. */
int dvgo_abi_version(void);
/* Name of the compiled architecture ("sm_100a"). */
const char* dvgo_build_arch(void);
/* Number of CUDA kernels this library has launched in this process (monotonic counter). */
unsigned long long dvgo_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * a1  render_utils_cuda.infer_t_minmax        lib/cuda/render_utils.cpp:44-52,
 *                                              lib/cuda/render_utils_kernel.cu:12-35,76-98
 * rays_o, rays_d [n_rays,3]; xyz_min, xyz_max [3]; t_min, t_max [n_rays] (outputs).
 */
int dvgo_infer_t_minmax(const float* rays_o, const float* rays_d, const float* xyz_min,
                        const float* xyz_max, float near, float far, int n_rays, float* t_min,
                        float* t_max, dvgo_stream_t stream);

/* a2  render_utils_cuda.infer_n_samples       render_utils.cpp:54-58, render_utils_kernel.cu:38-49,100-114 */
int dvgo_infer_n_samples(const float* t_min, const float* t_max, float stepdist, int n_rays,
                         int64_t* n_samples, dvgo_stream_t stream);

/* a3  render_utils_cuda.infer_ray_start_dir   render_utils.cpp:60-65, render_utils_kernel.cu:52-73,116-132 */
int dvgo_infer_ray_start_dir(const float* rays_o, const float* rays_d, const float* t_min,
                             int n_rays, float* rays_start, float* rays_dir, dvgo_stream_t stream);

/* a4  render_utils_cuda.sample_pts_on_rays    render_utils.cpp:67-78, render_utils_kernel.cu:138-236
 * The output length is data dependent (the reference syncs at render_utils_kernel.cu:206), so the
 * op is split in two: `count` writes t_min,t_max [n_rays], N_steps [n_rays] (int64), their
 * inclusive prefix sum N_steps_cumsum [n_rays] (int64) and the grand total into *total_host, a
 * HOST int64 (this call synchronises `stream` once -- the reference's one sync); the caller
 * allocates the five [total] outputs; `fill` writes them.
 * n_rays must be >= 0 and the total must be < 2^31 (the reference truncates to int, :206).
 */
int dvgo_sample_pts_count(const float* rays_o, const float* rays_d, const float* xyz_min,
                          const float* xyz_max, float near, float far, float stepdist, int n_rays,
                          float* t_min, float* t_max, int64_t* N_steps, int64_t* N_steps_cumsum,
                          int64_t* total_host, dvgo_stream_t stream);
int dvgo_sample_pts_fill(const float* rays_o, const float* rays_d, const float* xyz_min,
                         const float* xyz_max, const float* t_min, const int64_t* N_steps_cumsum,
                         float stepdist, int n_rays, int64_t total, float* rays_pts /*[total,3]*/,
                         uint8_t* mask_outbbox /*[total] bool*/, int64_t* ray_id, int64_t* step_id,
                         dvgo_stream_t stream);

/* a5  render_utils_cuda.sample_ndc_pts_on_rays  render_utils.cpp:80-91, render_utils_kernel.cu:239-287
 * rays_pts [n_rays,N_samples,3], mask_outbbox [n_rays,N_samples]. */
int dvgo_sample_ndc_pts_on_rays(const float* rays_o, const float* rays_d, const float* xyz_min,
                                const float* xyz_max, int N_samples, int n_rays, float* rays_pts,
                                uint8_t* mask_outbbox, dvgo_stream_t stream);

/* a6  render_utils_cuda.maskcache_lookup      render_utils.cpp:93-102, render_utils_kernel.cu:294-351
 * world [sz_i,sz_j,sz_k] bool; xyz [n_pts,3]; out [n_pts] bool: every element is written
 * (False when the rounded index is out of range), so `out` need not be pre-zeroed. */
int dvgo_maskcache_lookup(const uint8_t* world, const float* xyz, const float* xyz2ijk_scale,
                          const float* xyz2ijk_shift, int sz_i, int sz_j, int sz_k, int64_t n_pts,
                          uint8_t* out, dvgo_stream_t stream);

/* a8  render_utils_cuda.raw2alpha / raw2alpha_backward   render_utils.cpp:104-114,
 *                                                        render_utils_kernel.cu:358-428 */
int dvgo_raw2alpha(const float* density, float shift, float interval, int64_t n_pts, float* exp_d,
                   float* alpha, dvgo_stream_t stream);
int dvgo_raw2alpha_backward(const float* exp_d, const float* grad_back, float interval,
                            int64_t n_pts, float* grad, dvgo_stream_t stream);

/* a9  render_utils_cuda.alpha2weight / alpha2weight_backward   render_utils.cpp:116-141,
 *                                                              render_utils_kernel.cu:431-561
 * ray_id [n_pts] int64 sorted ascending.  All five outputs are fully written (fills included:
 * weight=0, T=1, alphainv_last=1, i_start=i_end=0 as at :478-482), no pre-initialisation needed.
 * One warp per ray: a shuffle-based segmented exclusive product replaces the serial loop. */
int dvgo_alpha2weight(const float* alpha, const int64_t* ray_id, int n_rays, int64_t n_pts,
                      float* weight, float* T, float* alphainv_last, int64_t* i_start,
                      int64_t* i_end, dvgo_stream_t stream);
int dvgo_alpha2weight_backward(const float* alpha, const float* weight, const float* T,
                               const float* alphainv_last, const int64_t* i_start,
                               const int64_t* i_end, int n_rays, int64_t n_pts,
                               const float* grad_weights, const float* grad_last, float* grad,
                               dvgo_stream_t stream);

/* a7  DenseGrid trilinear sampling = DirectVoxGO.grid_sampler   lib/dvgo.py:312-328
 *     (F.grid_sample(grid[1,C,X,Y,Z], ind_norm, 'bilinear', align_corners=True), zero padding;
 *     also lib/dmpigo.py:164-171, lib/tri_dvgo.py:609-627).
 * grid [C,X,Y,Z] (the reference's NCDHW with N=1); xyz [n_pts,3] world coordinates;
 * out [n_pts,C].  The ind_norm arithmetic of dvgo.py:316 is done inside the kernel.
 * backward: grad_grid [C,X,Y,Z] += scatter(grad_out [n_pts,C]) with fp32 atomics (unordered). */
int dvgo_grid_sample_3d(const float* grid, int C, int X, int Y, int Z, const float* xyz,
                        const float* xyz_min, const float* xyz_max, int64_t n_pts, float* out,
                        dvgo_stream_t stream);
int dvgo_grid_sample_3d_backward(const float* grad_out, int C, int X, int Y, int Z,
                                 const float* xyz, const float* xyz_min, const float* xyz_max,
                                 int64_t n_pts, float* grad_grid, dvgo_stream_t stream);

/* a7 (2-D)  tri-plane bilinear sampling   lib/tri_dvgo.py:456-464 (grid_sampler2D):
 *     F.grid_sample(plane[1,C,H,W], ind_norm[..., [a, b]], 'bilinear', align_corners=True), zero padding.
 * The reference feeds the FLIPPED normalised coordinate ind_norm = (z_n, y_n, x_n); axis_w / axis_h name the
 * world axis (0 = x, 1 = y, 2 = z) whose normalised coordinate indexes the plane's W / H dimension:
 *     'xy': ind_norm[..., [0,1]] -> axis_w = 2, axis_h = 1;   'yz': [1,2] -> axis_w = 1, axis_h = 0;
 *     'zx': [2,0] -> axis_w = 0, axis_h = 2.
 * plane [C,H,W]; xyz [n_pts,3] world coordinates; out [n_pts,C].  backward: grad_plane += scatter(grad_out). */
int dvgo_grid_sample_2d(const float* plane, int C, int H, int W, const float* xyz, const float* xyz_min,
                        const float* xyz_max, int axis_w, int axis_h, int64_t n_pts, float* out,
                        dvgo_stream_t stream);
int dvgo_grid_sample_2d_backward(const float* grad_out, int C, int H, int W, const float* xyz,
                                 const float* xyz_min, const float* xyz_max, int axis_w, int axis_h,
                                 int64_t n_pts, float* grad_plane, dvgo_stream_t stream);

/* a7 (literal F.grid_sample signature)  The unmodified lib/*.py call torch.nn.functional.grid_sample themselves
 * (lib/dvgo.py:321, lib/dmpigo.py:169, lib/tri_dvgo.py:462-464,618) with ALREADY-normalised coordinates.
 * These variants take that `grid` argument as it is: ind_norm [n_pts,3] = (w, h, d) = (z_n, y_n, x_n) for the 5-D
 * case, [n_pts,2] = (w, h) for the 4-D case; only ATen's align_corners=True un-normalisation is applied.
 * Outputs / gradients as above ([n_pts,C]; the binding returns the [1,C,...] view ATen would).
 * `directvoxgo_b200.dropin.install(grid_sample=True)` routes F.grid_sample here. */
int dvgo_grid_sample_3d_norm(const float* grid, int C, int X, int Y, int Z, const float* ind_norm,
                             int64_t n_pts, float* out, dvgo_stream_t stream);
int dvgo_grid_sample_3d_norm_backward(const float* grad_out, int C, int X, int Y, int Z,
                                      const float* ind_norm, int64_t n_pts, float* grad_grid,
                                      dvgo_stream_t stream);
int dvgo_grid_sample_2d_norm(const float* plane, int C, int H, int W, const float* ind_norm, int64_t n_pts,
                             float* out, dvgo_stream_t stream);
int dvgo_grid_sample_2d_norm_backward(const float* grad_out, int C, int H, int W, const float* ind_norm,
                                      int64_t n_pts, float* grad_plane, dvgo_stream_t stream);

/* a10 torch_scatter.segment_coo(src, index, out, reduce='sum')   lib/dvgo.py:554-558,571-575
 * src [n_pts,D] fp32, index [n_pts] int64 sorted ascending, out [n_seg,D] accumulated INTO (the
 * caller passes zeros, as the reference does).  Deterministic: one warp per run of equal indices.
 * The backward is a row gather: grad_src[p,:] = grad_out[index[p],:]. */
int dvgo_segment_coo_sum(const float* src, const int64_t* index, int64_t n_pts, int D,
                         int64_t n_seg, float* out, dvgo_stream_t stream);
int dvgo_gather_rows(const float* table, const int64_t* index, int64_t n_pts, int D, float* out,
                     dvgo_stream_t stream);

/* a12 total_variation_cuda.total_variation_add_grad   lib/cuda/total_variation.cpp:16-20,
 *                                                     lib/cuda/total_variation_kernel.cu:13-67
 * param, grad [1,C,sz_i,sz_j,sz_k] -> N = C*sz_i*sz_j*sz_k elements.  Reference quirks kept:
 * weights are divided by 6, wx is unused and the i-axis uses wz (:31-32). */
int dvgo_total_variation_add_grad(const float* param, float* grad, float wx, float wy, float wz,
                                  int dense_mode, int64_t N, int64_t sz_i, int64_t sz_j,
                                  int64_t sz_k, dvgo_stream_t stream);

/* a13 adam_upd_cuda.{adam_upd, masked_adam_upd, adam_upd_with_perlr}   lib/cuda/adam_upd.cpp:36-77,
 *                                                                      lib/cuda/adam_upd_kernel.cu:8-132 */
int dvgo_adam_upd(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t N,
                  int step, float beta1, float beta2, float lr, float eps, dvgo_stream_t stream);
int dvgo_masked_adam_upd(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                         int64_t N, int step, float beta1, float beta2, float lr, float eps,
                         dvgo_stream_t stream);
int dvgo_adam_upd_with_perlr(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                             const float* perlr, int64_t N, int step, float beta1, float beta2,
                             float lr, float eps, dvgo_stream_t stream);

#ifdef __cplusplus
} /* extern "C" */
#endif
#endif /* DVGO_B200_H_ */
