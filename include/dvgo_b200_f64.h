/*
 * dvgo_b200_f64.h -- the float64 instantiations of the reference's three extensions, C ABI.
 *
 * The reference dispatches every kernel of render_utils_cuda, total_variation_cuda and adam_upd_cuda
 * over AT_DISPATCH_FLOATING_TYPES (lib/cuda/render_utils_kernel.cu:86,105,122,223,276,340,384,419,493,546;
 * lib/cuda/total_variation_kernel.cu:50,59; lib/cuda/adam_upd_kernel.cu:74,98,123), i.e. the same Python
 * call also accepts float64 tensors.  Its models are fp32, so this instantiation is not on the training
 * hot path; it exists for `torch.autograd.gradcheck` of the autograd glue and for callers that keep
 * rays in double.  These entry points are that instantiation; the torch binding selects them when the
 * tensors are float64, exactly where the reference's dispatch macro would.
 *
 * "double" here means what the reference's template means, NOT uniformly double arithmetic: the
 * reference keeps several temporaries in `float` whatever scalar_t is (vx/ax/.. and the clip result at
 * render_utils_kernel.cu:23-33, rnorm :62, dist and px/py/pz :178-181 and :254-257, T_cum :447,
 * back_cum :522, grad_to_add total_variation_kernel.cu:25, the host-computed step_size and all
 * scalar arguments).  Each kernel here reproduces those roundings, so the results equal the
 * reference's double instantiation (bit for bit except libm-free FMA contraction choices, which follow
 * the reference's SASS; see directvoxgo_b200/csrc/f64_ops.cu).
 *
 * Conventions, return codes and the meaning of every argument are those of dvgo_b200.h with
 * `float*` tensors replaced by `double*`; scalar arguments stay `float` as in the reference's
 * signatures (render_utils.cpp:9-38).
 */
#ifndef DVGO_B200_F64_H_
#define DVGO_B200_F64_H_

#include "dvgo_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* a1  infer_t_minmax<double>        render_utils_kernel.cu:12-35 */
int dvgo_infer_t_minmax_f64(const double* rays_o, const double* rays_d, const double* xyz_min,
                            const double* xyz_max, float near, float far, int n_rays, double* t_min,
                            double* t_max, dvgo_stream_t stream);
/* a2  infer_n_samples<double>       render_utils_kernel.cu:38-49 (ceil and max in double) */
int dvgo_infer_n_samples_f64(const double* t_min, const double* t_max, float stepdist, int n_rays,
                             int64_t* n_samples, dvgo_stream_t stream);
/* a3  infer_ray_start_dir<double>   render_utils_kernel.cu:52-73 (rnorm rounded to float, :62) */
int dvgo_infer_ray_start_dir_f64(const double* rays_o, const double* rays_d, const double* t_min,
                                 int n_rays, double* rays_start, double* rays_dir,
                                 dvgo_stream_t stream);
/* a4  sample_pts_on_rays<double>    render_utils_kernel.cu:138-236; two-phase like dvgo_sample_pts_count/_fill */
int dvgo_sample_pts_count_f64(const double* rays_o, const double* rays_d, const double* xyz_min,
                              const double* xyz_max, float near, float far, float stepdist,
                              int n_rays, double* t_min, double* t_max, int64_t* N_steps,
                              int64_t* N_steps_cumsum, int64_t* total_host, dvgo_stream_t stream);
int dvgo_sample_pts_fill_f64(const double* rays_o, const double* rays_d, const double* xyz_min,
                             const double* xyz_max, const double* t_min,
                             const int64_t* N_steps_cumsum, float stepdist, int n_rays, int64_t total,
                             double* rays_pts, uint8_t* mask_outbbox, int64_t* ray_id,
                             int64_t* step_id, dvgo_stream_t stream);
/* a5  sample_ndc_pts_on_rays<double>  render_utils_kernel.cu:239-287 */
int dvgo_sample_ndc_pts_on_rays_f64(const double* rays_o, const double* rays_d,
                                    const double* xyz_min, const double* xyz_max, int N_samples,
                                    int n_rays, double* rays_pts, uint8_t* mask_outbbox,
                                    dvgo_stream_t stream);
/* a6  maskcache_lookup<double>      render_utils_kernel.cu:294-351 */
int dvgo_maskcache_lookup_f64(const uint8_t* world, const double* xyz, const double* xyz2ijk_scale,
                              const double* xyz2ijk_shift, int sz_i, int sz_j, int sz_k,
                              int64_t n_pts, uint8_t* out, dvgo_stream_t stream);
/* a8  raw2alpha<double> / raw2alpha_backward<double>   render_utils_kernel.cu:358-428 */
int dvgo_raw2alpha_f64(const double* density, float shift, float interval, int64_t n_pts,
                       double* exp_d, double* alpha, dvgo_stream_t stream);
int dvgo_raw2alpha_backward_f64(const double* exp_d, const double* grad_back, float interval,
                                int64_t n_pts, double* grad, dvgo_stream_t stream);
/* a9  alpha2weight<double> / alpha2weight_backward<double>   render_utils_kernel.cu:431-561
 * (the running transmittance and the backward accumulator are `float`, :447 and :522) */
int dvgo_alpha2weight_f64(const double* alpha, const int64_t* ray_id, int n_rays, int64_t n_pts,
                          double* weight, double* T, double* alphainv_last, int64_t* i_start,
                          int64_t* i_end, dvgo_stream_t stream);
int dvgo_alpha2weight_backward_f64(const double* alpha, const double* weight, const double* T,
                                   const double* alphainv_last, const int64_t* i_start,
                                   const int64_t* i_end, int n_rays, int64_t n_pts,
                                   const double* grad_weights, const double* grad_last, double* grad,
                                   dvgo_stream_t stream);
/* a12 total_variation_add_grad<double>   total_variation_kernel.cu:13-67 (float accumulator, :25) */
int dvgo_total_variation_add_grad_f64(const double* param, double* grad, float wx, float wy, float wz,
                                      int dense_mode, int64_t N, int64_t sz_i, int64_t sz_j,
                                      int64_t sz_k, dvgo_stream_t stream);
/* a13 adam_upd<double>, masked_adam_upd<double>, adam_upd_with_perlr<double>   adam_upd_kernel.cu:8-132 */
int dvgo_adam_upd_f64(double* param, const double* grad, double* exp_avg, double* exp_avg_sq,
                      int64_t N, int step, float beta1, float beta2, float lr, float eps,
                      dvgo_stream_t stream);
int dvgo_masked_adam_upd_f64(double* param, const double* grad, double* exp_avg, double* exp_avg_sq,
                             int64_t N, int step, float beta1, float beta2, float lr, float eps,
                             dvgo_stream_t stream);
int dvgo_adam_upd_with_perlr_f64(double* param, const double* grad, double* exp_avg,
                                 double* exp_avg_sq, const double* perlr, int64_t N, int step,
                                 float beta1, float beta2, float lr, float eps, dvgo_stream_t stream);

#ifdef __cplusplus
} /* extern "C" */
#endif
#endif /* DVGO_B200_F64_H_ */
