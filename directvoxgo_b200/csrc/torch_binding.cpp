// torch_binding.cpp -- the thin torch <-> C-ABI adaptor.
//
// Exposes, with the reference's exact Python-visible names, argument order and return structure,
//   render_utils_cuda     (lib/cuda/render_utils.cpp:144-155)
//   total_variation_cuda  (lib/cuda/total_variation.cpp:22-24)
//   adam_upd_cuda         (lib/cuda/adam_upd.cpp:79-86)
// as sub-modules of directvoxgo_b200._C, plus `ext` with the ops the reference takes from ATen /
// torch_scatter (grid_sample, segment_coo) and the fused trainer entry points.
//
// All it does is: check the reference's preconditions (CHECK_CUDA / CHECK_CONTIGUOUS ->
// RuntimeError, shape asserts), guard the device, allocate outputs with torch, and pass raw
// pointers + the current CUDA stream to include/dvgo_b200.h.  There is NO CPU path: a non-CUDA
// tensor is an error, exactly as in the reference (render_utils.cpp:40).
//
// The three reference modules dispatch on the tensor dtype (AT_DISPATCH_FLOATING_TYPES, e.g.
// lib/cuda/render_utils_kernel.cu:86): float32 goes to include/dvgo_b200.h, float64 to the
// double instantiation in include/dvgo_b200_f64.h.  Any other dtype, or tensors of mixed
// floating dtypes, is an error (the reference's data<scalar_t>() throws there too).
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include <vector>

#include "../../include/dvgo_b200.h"
#include "../../include/dvgo_b200_f64.h"

namespace {

using torch::Tensor;

#define CHECK_CUDA(x) TORCH_CHECK((x).is_cuda(), #x " must be a CUDA tensor")
#define CHECK_CONTIGUOUS(x) TORCH_CHECK((x).is_contiguous(), #x " must be contiguous")
#define CHECK_INPUT(x) \
  CHECK_CUDA(x);       \
  CHECK_CONTIGUOUS(x)
#define CHECK_F32(x) TORCH_CHECK((x).scalar_type() == torch::kFloat32, #x " must be float32")
// float32 or float64, and every further floating tensor of the call has the dtype of the first
#define CHECK_FLT(x)                                                                              \
  TORCH_CHECK((x).scalar_type() == torch::kFloat32 || (x).scalar_type() == torch::kFloat64, #x    \
              " must be float32 or float64")
#define CHECK_SAME(x, ref) \
  TORCH_CHECK((x).scalar_type() == (ref).scalar_type(), #x " must have the dtype of " #ref)
#define CHECK_I64(x) TORCH_CHECK((x).scalar_type() == torch::kInt64, #x " must be int64")
#define CHECK_BOOL(x) TORCH_CHECK((x).scalar_type() == torch::kBool, #x " must be bool")

inline dvgo_stream_t cur_stream() {
  return reinterpret_cast<dvgo_stream_t>(at::cuda::getCurrentCUDAStream().stream());
}

inline void check_rc(int rc, const char* what) {
  TORCH_CHECK(rc == 0, "dvgo_b200: ", what, " failed with code ", rc,
              rc > 0 ? (std::string(" (") + cudaGetErrorString(static_cast<cudaError_t>(rc)) + ")")
                     : std::string(" (invalid argument)"));
}

inline const float* fp(const Tensor& t) { return t.data_ptr<float>(); }
inline float* fpm(Tensor& t) { return t.data_ptr<float>(); }
inline bool is_f64(const Tensor& t) { return t.scalar_type() == torch::kFloat64; }
inline const double* dp(const Tensor& t) { return t.data_ptr<double>(); }
inline double* dpm(Tensor& t) { return t.data_ptr<double>(); }
inline const int64_t* ip(const Tensor& t) { return t.data_ptr<int64_t>(); }
inline const uint8_t* bp(const Tensor& t) { return reinterpret_cast<const uint8_t*>(t.data_ptr<bool>()); }
inline uint8_t* bpm(Tensor& t) { return reinterpret_cast<uint8_t*>(t.data_ptr<bool>()); }

// ------------------------------------------------------------------------------------------------
// render_utils_cuda
// ------------------------------------------------------------------------------------------------
std::vector<Tensor> infer_t_minmax(Tensor rays_o, Tensor rays_d, Tensor xyz_min, Tensor xyz_max,
                                   const float near, const float far) {
  CHECK_INPUT(rays_o); CHECK_INPUT(rays_d); CHECK_INPUT(xyz_min); CHECK_INPUT(xyz_max);
  CHECK_FLT(rays_o); CHECK_SAME(rays_d, rays_o); CHECK_SAME(xyz_min, rays_o); CHECK_SAME(xyz_max, rays_o);
  const c10::cuda::CUDAGuard guard(rays_o.device());
  const int n_rays = rays_o.size(0);
  auto t_min = torch::empty({n_rays}, rays_o.options());
  auto t_max = torch::empty({n_rays}, rays_o.options());
  if (is_f64(rays_o)) {
    check_rc(dvgo_infer_t_minmax_f64(dp(rays_o), dp(rays_d), dp(xyz_min), dp(xyz_max), near, far, n_rays,
                                     dpm(t_min), dpm(t_max), cur_stream()),
             "infer_t_minmax<double>");
    return {t_min, t_max};
  }
  check_rc(dvgo_infer_t_minmax(fp(rays_o), fp(rays_d), fp(xyz_min), fp(xyz_max), near, far, n_rays,
                               fpm(t_min), fpm(t_max), cur_stream()),
           "infer_t_minmax");
  return {t_min, t_max};
}

Tensor infer_n_samples(Tensor t_min, Tensor t_max, const float stepdist) {
  CHECK_INPUT(t_min); CHECK_INPUT(t_max); CHECK_FLT(t_min); CHECK_SAME(t_max, t_min);
  const c10::cuda::CUDAGuard guard(t_min.device());
  const int n_rays = t_min.size(0);
  auto n_samples = torch::empty({n_rays}, t_min.options().dtype(torch::kInt64));
  if (is_f64(t_min)) {
    check_rc(dvgo_infer_n_samples_f64(dp(t_min), dp(t_max), stepdist, n_rays, n_samples.data_ptr<int64_t>(),
                                      cur_stream()),
             "infer_n_samples<double>");
    return n_samples;
  }
  check_rc(dvgo_infer_n_samples(fp(t_min), fp(t_max), stepdist, n_rays,
                                n_samples.data_ptr<int64_t>(), cur_stream()),
           "infer_n_samples");
  return n_samples;
}

std::vector<Tensor> infer_ray_start_dir(Tensor rays_o, Tensor rays_d, Tensor t_min) {
  CHECK_INPUT(rays_o); CHECK_INPUT(rays_d); CHECK_INPUT(t_min);
  CHECK_FLT(rays_o); CHECK_SAME(rays_d, rays_o); CHECK_SAME(t_min, rays_o);
  const c10::cuda::CUDAGuard guard(rays_o.device());
  const int n_rays = rays_o.size(0);
  auto rays_start = torch::empty_like(rays_o);
  auto rays_dir = torch::empty_like(rays_o);
  if (is_f64(rays_o)) {
    check_rc(dvgo_infer_ray_start_dir_f64(dp(rays_o), dp(rays_d), dp(t_min), n_rays, dpm(rays_start),
                                          dpm(rays_dir), cur_stream()),
             "infer_ray_start_dir<double>");
    return {rays_start, rays_dir};
  }
  check_rc(dvgo_infer_ray_start_dir(fp(rays_o), fp(rays_d), fp(t_min), n_rays, fpm(rays_start),
                                    fpm(rays_dir), cur_stream()),
           "infer_ray_start_dir");
  return {rays_start, rays_dir};
}

std::vector<Tensor> sample_pts_on_rays(Tensor rays_o, Tensor rays_d, Tensor xyz_min, Tensor xyz_max,
                                       const float near, const float far, const float stepdist) {
  CHECK_INPUT(rays_o); CHECK_INPUT(rays_d); CHECK_INPUT(xyz_min); CHECK_INPUT(xyz_max);
  CHECK_FLT(rays_o); CHECK_SAME(rays_d, rays_o); CHECK_SAME(xyz_min, rays_o); CHECK_SAME(xyz_max, rays_o);
  TORCH_CHECK(rays_o.dim() == 2 && rays_o.size(1) == 3, "rays_o must be [N,3]");
  TORCH_CHECK(rays_d.sizes() == rays_o.sizes(), "rays_d must match rays_o");
  const c10::cuda::CUDAGuard guard(rays_o.device());
  const int n_rays = rays_o.size(0);
  auto fopt = rays_o.options();
  auto iopt = rays_o.options().dtype(torch::kInt64);
  auto t_min = torch::empty({n_rays}, fopt);
  auto t_max = torch::empty({n_rays}, fopt);
  auto N_steps = torch::empty({n_rays}, iopt);
  auto cumsum = torch::empty({n_rays}, iopt);
  int64_t total = 0;
  if (is_f64(rays_o)) {
    check_rc(dvgo_sample_pts_count_f64(dp(rays_o), dp(rays_d), dp(xyz_min), dp(xyz_max), near, far, stepdist,
                                       n_rays, dpm(t_min), dpm(t_max), N_steps.data_ptr<int64_t>(),
                                       cumsum.data_ptr<int64_t>(), &total, cur_stream()),
             "sample_pts_count<double>");
    auto rays_pts = torch::empty({total, 3}, fopt);
    auto mask_outbbox = torch::empty({total}, fopt.dtype(torch::kBool));
    auto ray_id = torch::empty({total}, iopt);
    auto step_id = torch::empty({total}, iopt);
    check_rc(dvgo_sample_pts_fill_f64(dp(rays_o), dp(rays_d), dp(xyz_min), dp(xyz_max), dp(t_min), ip(cumsum),
                                      stepdist, n_rays, total, dpm(rays_pts), bpm(mask_outbbox),
                                      ray_id.data_ptr<int64_t>(), step_id.data_ptr<int64_t>(), cur_stream()),
             "sample_pts_fill<double>");
    return {rays_pts, mask_outbbox, ray_id, step_id, N_steps, t_min, t_max};
  }
  check_rc(dvgo_sample_pts_count(fp(rays_o), fp(rays_d), fp(xyz_min), fp(xyz_max), near, far,
                                 stepdist, n_rays, fpm(t_min), fpm(t_max),
                                 N_steps.data_ptr<int64_t>(), cumsum.data_ptr<int64_t>(), &total,
                                 cur_stream()),
           "sample_pts_count");
  auto rays_pts = torch::empty({total, 3}, fopt);
  auto mask_outbbox = torch::empty({total}, fopt.dtype(torch::kBool));
  auto ray_id = torch::empty({total}, iopt);
  auto step_id = torch::empty({total}, iopt);
  check_rc(dvgo_sample_pts_fill(fp(rays_o), fp(rays_d), fp(xyz_min), fp(xyz_max), fp(t_min),
                                ip(cumsum), stepdist, n_rays, total, fpm(rays_pts),
                                bpm(mask_outbbox), ray_id.data_ptr<int64_t>(),
                                step_id.data_ptr<int64_t>(), cur_stream()),
           "sample_pts_fill");
  return {rays_pts, mask_outbbox, ray_id, step_id, N_steps, t_min, t_max};
}

std::vector<Tensor> sample_ndc_pts_on_rays(Tensor rays_o, Tensor rays_d, Tensor xyz_min,
                                           Tensor xyz_max, const int N_samples) {
  CHECK_INPUT(rays_o); CHECK_INPUT(rays_d); CHECK_INPUT(xyz_min); CHECK_INPUT(xyz_max);
  CHECK_FLT(rays_o); CHECK_SAME(rays_d, rays_o); CHECK_SAME(xyz_min, rays_o); CHECK_SAME(xyz_max, rays_o);
  TORCH_CHECK(rays_o.dim() == 2 && rays_o.size(1) == 3, "rays_o must be [N,3]");
  const c10::cuda::CUDAGuard guard(rays_o.device());
  const int n_rays = rays_o.size(0);
  auto rays_pts = torch::empty({n_rays, N_samples, 3}, rays_o.options());
  auto mask_outbbox = torch::empty({n_rays, N_samples}, rays_o.options().dtype(torch::kBool));
  if (is_f64(rays_o)) {
    check_rc(dvgo_sample_ndc_pts_on_rays_f64(dp(rays_o), dp(rays_d), dp(xyz_min), dp(xyz_max), N_samples, n_rays,
                                             dpm(rays_pts), bpm(mask_outbbox), cur_stream()),
             "sample_ndc_pts_on_rays<double>");
    return {rays_pts, mask_outbbox};
  }
  check_rc(dvgo_sample_ndc_pts_on_rays(fp(rays_o), fp(rays_d), fp(xyz_min), fp(xyz_max), N_samples,
                                       n_rays, fpm(rays_pts), bpm(mask_outbbox), cur_stream()),
           "sample_ndc_pts_on_rays");
  return {rays_pts, mask_outbbox};
}

Tensor maskcache_lookup(Tensor world, Tensor xyz, Tensor xyz2ijk_scale, Tensor xyz2ijk_shift) {
  CHECK_INPUT(world); CHECK_INPUT(xyz); CHECK_INPUT(xyz2ijk_scale); CHECK_INPUT(xyz2ijk_shift);
  CHECK_BOOL(world); CHECK_FLT(xyz); CHECK_SAME(xyz2ijk_scale, xyz); CHECK_SAME(xyz2ijk_shift, xyz);
  TORCH_CHECK(world.dim() == 3, "world must be [X,Y,Z]");
  TORCH_CHECK(xyz.dim() == 2 && xyz.size(1) == 3, "xyz must be [P,3]");
  const c10::cuda::CUDAGuard guard(xyz.device());
  const int64_t n_pts = xyz.size(0);
  auto out = torch::empty({n_pts}, xyz.options().dtype(torch::kBool));
  if (is_f64(xyz)) {
    check_rc(dvgo_maskcache_lookup_f64(bp(world), dp(xyz), dp(xyz2ijk_scale), dp(xyz2ijk_shift), world.size(0),
                                       world.size(1), world.size(2), n_pts, bpm(out), cur_stream()),
             "maskcache_lookup<double>");
    return out;
  }
  check_rc(dvgo_maskcache_lookup(bp(world), fp(xyz), fp(xyz2ijk_scale), fp(xyz2ijk_shift),
                                 world.size(0), world.size(1), world.size(2), n_pts, bpm(out),
                                 cur_stream()),
           "maskcache_lookup");
  return out;
}

std::vector<Tensor> raw2alpha(Tensor density, const float shift, const float interval) {
  CHECK_INPUT(density); CHECK_FLT(density);
  TORCH_CHECK(density.dim() == 1, "density must be 1-D");
  const c10::cuda::CUDAGuard guard(density.device());
  auto exp_d = torch::empty_like(density);
  auto alpha = torch::empty_like(density);
  if (is_f64(density)) {
    check_rc(dvgo_raw2alpha_f64(dp(density), shift, interval, density.size(0), dpm(exp_d), dpm(alpha),
                                cur_stream()),
             "raw2alpha<double>");
    return {exp_d, alpha};
  }
  check_rc(dvgo_raw2alpha(fp(density), shift, interval, density.size(0), fpm(exp_d), fpm(alpha),
                          cur_stream()),
           "raw2alpha");
  return {exp_d, alpha};
}

Tensor raw2alpha_backward(Tensor exp, Tensor grad_back, const float interval) {
  CHECK_INPUT(exp); CHECK_INPUT(grad_back); CHECK_FLT(exp); CHECK_SAME(grad_back, exp);
  TORCH_CHECK(exp.numel() == grad_back.numel(), "exp and grad_back must have the same numel");
  const c10::cuda::CUDAGuard guard(exp.device());
  auto grad = torch::empty_like(exp);
  if (is_f64(exp)) {
    check_rc(dvgo_raw2alpha_backward_f64(dp(exp), dp(grad_back), interval, exp.numel(), dpm(grad), cur_stream()),
             "raw2alpha_backward<double>");
    return grad;
  }
  check_rc(dvgo_raw2alpha_backward(fp(exp), fp(grad_back), interval, exp.numel(), fpm(grad),
                                   cur_stream()),
           "raw2alpha_backward");
  return grad;
}

std::vector<Tensor> alpha2weight(Tensor alpha, Tensor ray_id, const int n_rays) {
  CHECK_INPUT(alpha); CHECK_INPUT(ray_id); CHECK_FLT(alpha); CHECK_I64(ray_id);
  TORCH_CHECK(alpha.dim() == 1 && ray_id.dim() == 1 && alpha.sizes() == ray_id.sizes(),
              "alpha and ray_id must be 1-D of equal length");
  const c10::cuda::CUDAGuard guard(alpha.device());
  const int64_t n_pts = alpha.size(0);
  auto weight = torch::empty_like(alpha);
  auto T = torch::empty_like(alpha);
  auto alphainv_last = torch::empty({n_rays}, alpha.options());
  auto i_start = torch::empty({n_rays}, alpha.options().dtype(torch::kInt64));
  auto i_end = torch::empty({n_rays}, alpha.options().dtype(torch::kInt64));
  if (is_f64(alpha)) {
    check_rc(dvgo_alpha2weight_f64(dp(alpha), ip(ray_id), n_rays, n_pts, dpm(weight), dpm(T), dpm(alphainv_last),
                                   i_start.data_ptr<int64_t>(), i_end.data_ptr<int64_t>(), cur_stream()),
             "alpha2weight<double>");
    return {weight, T, alphainv_last, i_start, i_end};
  }
  check_rc(dvgo_alpha2weight(fp(alpha), ip(ray_id), n_rays, n_pts, fpm(weight), fpm(T),
                             fpm(alphainv_last), i_start.data_ptr<int64_t>(),
                             i_end.data_ptr<int64_t>(), cur_stream()),
           "alpha2weight");
  return {weight, T, alphainv_last, i_start, i_end};
}

Tensor alpha2weight_backward(Tensor alpha, Tensor weight, Tensor T, Tensor alphainv_last,
                             Tensor i_start, Tensor i_end, const int n_rays, Tensor grad_weights,
                             Tensor grad_last) {
  CHECK_INPUT(alpha); CHECK_INPUT(weight); CHECK_INPUT(T); CHECK_INPUT(alphainv_last);
  CHECK_INPUT(i_start); CHECK_INPUT(i_end); CHECK_INPUT(grad_weights); CHECK_INPUT(grad_last);
  CHECK_FLT(alpha); CHECK_SAME(weight, alpha); CHECK_SAME(T, alpha); CHECK_SAME(alphainv_last, alpha);
  CHECK_I64(i_start); CHECK_I64(i_end); CHECK_SAME(grad_weights, alpha); CHECK_SAME(grad_last, alpha);
  const c10::cuda::CUDAGuard guard(alpha.device());
  auto grad = torch::empty_like(alpha);
  if (is_f64(alpha)) {
    check_rc(dvgo_alpha2weight_backward_f64(dp(alpha), dp(weight), dp(T), dp(alphainv_last), ip(i_start),
                                            ip(i_end), n_rays, alpha.numel(), dp(grad_weights), dp(grad_last),
                                            dpm(grad), cur_stream()),
             "alpha2weight_backward<double>");
    return grad;
  }
  check_rc(dvgo_alpha2weight_backward(fp(alpha), fp(weight), fp(T), fp(alphainv_last), ip(i_start),
                                      ip(i_end), n_rays, alpha.numel(), fp(grad_weights),
                                      fp(grad_last), fpm(grad), cur_stream()),
           "alpha2weight_backward");
  return grad;
}

// ------------------------------------------------------------------------------------------------
// total_variation_cuda / adam_upd_cuda
// ------------------------------------------------------------------------------------------------
void total_variation_add_grad(Tensor param, Tensor grad, float wx, float wy, float wz,
                              bool dense_mode) {
  CHECK_INPUT(param); CHECK_INPUT(grad); CHECK_FLT(param); CHECK_SAME(grad, param);
  TORCH_CHECK(param.dim() == 5, "param must be [1,C,X,Y,Z]");
  TORCH_CHECK(param.sizes() == grad.sizes(), "param and grad must have the same shape");
  const c10::cuda::CUDAGuard guard(param.device());
  if (is_f64(param)) {
    check_rc(dvgo_total_variation_add_grad_f64(dp(param), dpm(grad), wx, wy, wz, dense_mode ? 1 : 0, param.numel(),
                                               param.size(2), param.size(3), param.size(4), cur_stream()),
             "total_variation_add_grad<double>");
    return;
  }
  check_rc(dvgo_total_variation_add_grad(fp(param), fpm(grad), wx, wy, wz, dense_mode ? 1 : 0,
                                         param.numel(), param.size(2), param.size(3), param.size(4),
                                         cur_stream()),
           "total_variation_add_grad");
}

#define ADAM_CHECKS()                                                                       \
  CHECK_INPUT(param); CHECK_INPUT(grad); CHECK_INPUT(exp_avg); CHECK_INPUT(exp_avg_sq);     \
  CHECK_FLT(param); CHECK_SAME(grad, param); CHECK_SAME(exp_avg, param); CHECK_SAME(exp_avg_sq, param); \
  TORCH_CHECK(param.numel() == grad.numel() && param.numel() == exp_avg.numel() &&          \
                  param.numel() == exp_avg_sq.numel(),                                      \
              "param, grad, exp_avg, exp_avg_sq must have the same numel");                 \
  const c10::cuda::CUDAGuard guard(param.device())

void adam_upd(Tensor param, Tensor grad, Tensor exp_avg, Tensor exp_avg_sq, int step, float beta1,
              float beta2, float lr, float eps) {
  ADAM_CHECKS();
  if (is_f64(param)) {
    check_rc(dvgo_adam_upd_f64(dpm(param), dp(grad), dpm(exp_avg), dpm(exp_avg_sq), param.numel(), step, beta1,
                               beta2, lr, eps, cur_stream()),
             "adam_upd<double>");
    return;
  }
  check_rc(dvgo_adam_upd(fpm(param), fp(grad), fpm(exp_avg), fpm(exp_avg_sq), param.numel(), step,
                         beta1, beta2, lr, eps, cur_stream()),
           "adam_upd");
}

void masked_adam_upd(Tensor param, Tensor grad, Tensor exp_avg, Tensor exp_avg_sq, int step,
                     float beta1, float beta2, float lr, float eps) {
  ADAM_CHECKS();
  if (is_f64(param)) {
    check_rc(dvgo_masked_adam_upd_f64(dpm(param), dp(grad), dpm(exp_avg), dpm(exp_avg_sq), param.numel(), step,
                                      beta1, beta2, lr, eps, cur_stream()),
             "masked_adam_upd<double>");
    return;
  }
  check_rc(dvgo_masked_adam_upd(fpm(param), fp(grad), fpm(exp_avg), fpm(exp_avg_sq), param.numel(),
                                step, beta1, beta2, lr, eps, cur_stream()),
           "masked_adam_upd");
}

void adam_upd_with_perlr(Tensor param, Tensor grad, Tensor exp_avg, Tensor exp_avg_sq, Tensor perlr,
                         int step, float beta1, float beta2, float lr, float eps) {
  ADAM_CHECKS();
  CHECK_INPUT(perlr); CHECK_SAME(perlr, param);
  TORCH_CHECK(perlr.numel() == param.numel(), "perlr must match param");
  if (is_f64(param)) {
    check_rc(dvgo_adam_upd_with_perlr_f64(dpm(param), dp(grad), dpm(exp_avg), dpm(exp_avg_sq), dp(perlr),
                                          param.numel(), step, beta1, beta2, lr, eps, cur_stream()),
             "adam_upd_with_perlr<double>");
    return;
  }
  check_rc(dvgo_adam_upd_with_perlr(fpm(param), fp(grad), fpm(exp_avg), fpm(exp_avg_sq), fp(perlr),
                                    param.numel(), step, beta1, beta2, lr, eps, cur_stream()),
           "adam_upd_with_perlr");
}

// ------------------------------------------------------------------------------------------------
// ext: ops the reference takes from ATen / torch_scatter
// ------------------------------------------------------------------------------------------------
// grid [1,C,X,Y,Z] (or [C,X,Y,Z]); xyz [P,3] world coords -> [P,C]
Tensor grid_sample_3d(Tensor grid, Tensor xyz, Tensor xyz_min, Tensor xyz_max) {
  CHECK_INPUT(grid); CHECK_INPUT(xyz); CHECK_INPUT(xyz_min); CHECK_INPUT(xyz_max);
  CHECK_F32(grid); CHECK_F32(xyz); CHECK_F32(xyz_min); CHECK_F32(xyz_max);
  TORCH_CHECK(grid.dim() == 5 && grid.size(0) == 1, "grid must be [1,C,X,Y,Z]");
  TORCH_CHECK(xyz.dim() == 2 && xyz.size(1) == 3, "xyz must be [P,3]");
  const c10::cuda::CUDAGuard guard(grid.device());
  const int C = grid.size(1);
  auto out = torch::empty({xyz.size(0), C}, xyz.options());
  check_rc(dvgo_grid_sample_3d(fp(grid), C, grid.size(2), grid.size(3), grid.size(4), fp(xyz),
                               fp(xyz_min), fp(xyz_max), xyz.size(0), fpm(out), cur_stream()),
           "grid_sample_3d");
  return out;
}

// grad_out [P,C] -> accumulates into grad_grid [1,C,X,Y,Z] (must be pre-initialised by the caller)
void grid_sample_3d_backward(Tensor grad_out, Tensor xyz, Tensor xyz_min, Tensor xyz_max,
                             Tensor grad_grid) {
  CHECK_INPUT(grad_out); CHECK_INPUT(xyz); CHECK_INPUT(xyz_min); CHECK_INPUT(xyz_max);
  CHECK_INPUT(grad_grid);
  CHECK_F32(grad_out); CHECK_F32(xyz); CHECK_F32(xyz_min); CHECK_F32(xyz_max); CHECK_F32(grad_grid);
  TORCH_CHECK(grad_grid.dim() == 5 && grad_grid.size(0) == 1, "grad_grid must be [1,C,X,Y,Z]");
  const int C = grad_grid.size(1);
  TORCH_CHECK(grad_out.dim() == 2 && grad_out.size(0) == xyz.size(0) && grad_out.size(1) == C,
              "grad_out must be [P,C]");
  const c10::cuda::CUDAGuard guard(grad_grid.device());
  check_rc(dvgo_grid_sample_3d_backward(fp(grad_out), C, grad_grid.size(2), grad_grid.size(3),
                                        grad_grid.size(4), fp(xyz), fp(xyz_min), fp(xyz_max),
                                        xyz.size(0), fpm(grad_grid), cur_stream()),
           "grid_sample_3d_backward");
}

// tri-plane: plane [1,C,H,W], xyz [P,3] -> [P,C]; axis_w / axis_h as in include/dvgo_b200.h
Tensor grid_sample_2d(Tensor plane, Tensor xyz, Tensor xyz_min, Tensor xyz_max, int axis_w, int axis_h) {
  CHECK_INPUT(plane); CHECK_INPUT(xyz); CHECK_INPUT(xyz_min); CHECK_INPUT(xyz_max);
  CHECK_F32(plane); CHECK_F32(xyz); CHECK_F32(xyz_min); CHECK_F32(xyz_max);
  TORCH_CHECK(plane.dim() == 4 && plane.size(0) == 1, "plane must be [1,C,H,W]");
  TORCH_CHECK(xyz.dim() == 2 && xyz.size(1) == 3, "xyz must be [P,3]");
  const c10::cuda::CUDAGuard guard(plane.device());
  const int C = plane.size(1);
  auto out = torch::empty({xyz.size(0), C}, xyz.options());
  check_rc(dvgo_grid_sample_2d(fp(plane), C, plane.size(2), plane.size(3), fp(xyz), fp(xyz_min), fp(xyz_max), axis_w,
                               axis_h, xyz.size(0), fpm(out), cur_stream()), "grid_sample_2d");
  return out;
}

void grid_sample_2d_backward(Tensor grad_out, Tensor xyz, Tensor xyz_min, Tensor xyz_max, int axis_w, int axis_h,
                             Tensor grad_plane) {
  CHECK_INPUT(grad_out); CHECK_INPUT(xyz); CHECK_INPUT(xyz_min); CHECK_INPUT(xyz_max); CHECK_INPUT(grad_plane);
  CHECK_F32(grad_out); CHECK_F32(xyz); CHECK_F32(xyz_min); CHECK_F32(xyz_max); CHECK_F32(grad_plane);
  TORCH_CHECK(grad_plane.dim() == 4 && grad_plane.size(0) == 1, "grad_plane must be [1,C,H,W]");
  const int C = grad_plane.size(1);
  TORCH_CHECK(grad_out.dim() == 2 && grad_out.size(0) == xyz.size(0) && grad_out.size(1) == C, "grad_out must be [P,C]");
  const c10::cuda::CUDAGuard guard(grad_plane.device());
  check_rc(dvgo_grid_sample_2d_backward(fp(grad_out), C, grad_plane.size(2), grad_plane.size(3), fp(xyz), fp(xyz_min),
                                        fp(xyz_max), axis_w, axis_h, xyz.size(0), fpm(grad_plane), cur_stream()),
           "grid_sample_2d_backward");
}

// F.grid_sample's own signature pieces: input [1,C,D,H,W] or [1,C,H,W]; pts [P,3] / [P,2] normalised coordinates -> [P,C]
Tensor grid_sample_norm(Tensor input, Tensor pts) {
  CHECK_INPUT(input); CHECK_INPUT(pts); CHECK_F32(input); CHECK_F32(pts);
  const int nd = input.dim() - 2;
  TORCH_CHECK((nd == 2 || nd == 3) && input.size(0) == 1, "input must be [1,C,H,W] or [1,C,D,H,W]");
  TORCH_CHECK(pts.dim() == 2 && pts.size(1) == nd, "pts must be [P,2] or [P,3]");
  const c10::cuda::CUDAGuard guard(input.device());
  const int C = input.size(1);
  auto out = torch::empty({pts.size(0), C}, pts.options());
  if (nd == 3)
    check_rc(dvgo_grid_sample_3d_norm(fp(input), C, input.size(2), input.size(3), input.size(4), fp(pts), pts.size(0),
                                      fpm(out), cur_stream()), "grid_sample_3d_norm");
  else
    check_rc(dvgo_grid_sample_2d_norm(fp(input), C, input.size(2), input.size(3), fp(pts), pts.size(0), fpm(out),
                                      cur_stream()), "grid_sample_2d_norm");
  return out;
}

void grid_sample_norm_backward(Tensor grad_out, Tensor pts, Tensor grad_input) {
  CHECK_INPUT(grad_out); CHECK_INPUT(pts); CHECK_INPUT(grad_input);
  CHECK_F32(grad_out); CHECK_F32(pts); CHECK_F32(grad_input);
  const int nd = grad_input.dim() - 2;
  TORCH_CHECK((nd == 2 || nd == 3) && grad_input.size(0) == 1, "grad_input must be [1,C,H,W] or [1,C,D,H,W]");
  const int C = grad_input.size(1);
  TORCH_CHECK(pts.dim() == 2 && pts.size(1) == nd && grad_out.dim() == 2 && grad_out.size(0) == pts.size(0) &&
              grad_out.size(1) == C, "grad_out must be [P,C], pts [P,nd]");
  const c10::cuda::CUDAGuard guard(grad_input.device());
  if (nd == 3)
    check_rc(dvgo_grid_sample_3d_norm_backward(fp(grad_out), C, grad_input.size(2), grad_input.size(3), grad_input.size(4),
                                               fp(pts), pts.size(0), fpm(grad_input), cur_stream()),
             "grid_sample_3d_norm_backward");
  else
    check_rc(dvgo_grid_sample_2d_norm_backward(fp(grad_out), C, grad_input.size(2), grad_input.size(3), fp(pts),
                                               pts.size(0), fpm(grad_input), cur_stream()),
             "grid_sample_2d_norm_backward");
}

// out[index[p], :] += src[p, :]   (src [P] or [P,D]; index sorted; out [N] or [N,D], in place)
void segment_coo_sum(Tensor src, Tensor index, Tensor out) {
  CHECK_INPUT(src); CHECK_INPUT(index); CHECK_INPUT(out);
  CHECK_F32(src); CHECK_I64(index); CHECK_F32(out);
  TORCH_CHECK(index.dim() == 1 && src.dim() >= 1 && src.size(0) == index.size(0),
              "index must be 1-D and match src.size(0)");
  const int64_t P = index.size(0);
  const int64_t D = P ? src.numel() / P : (out.dim() > 1 ? out.size(1) : 1);
  TORCH_CHECK(out.dim() >= 1 && (out.size(0) == 0 || out.numel() / out.size(0) == D),
              "out row width must match src");
  const c10::cuda::CUDAGuard guard(src.device());
  check_rc(dvgo_segment_coo_sum(fp(src), ip(index), P, static_cast<int>(D), out.size(0), fpm(out),
                                cur_stream()),
           "segment_coo_sum");
}

Tensor gather_rows(Tensor table, Tensor index) {
  CHECK_INPUT(table); CHECK_INPUT(index); CHECK_F32(table); CHECK_I64(index);
  TORCH_CHECK(index.dim() == 1 && table.dim() >= 1, "index must be 1-D");
  const int64_t D = table.size(0) ? table.numel() / table.size(0) : 1;
  auto sizes = table.sizes().vec();
  sizes[0] = index.size(0);
  const c10::cuda::CUDAGuard guard(table.device());
  auto out = torch::empty(sizes, table.options());
  check_rc(dvgo_gather_rows(fp(table), ip(index), index.size(0), static_cast<int>(D), fpm(out),
                            cur_stream()),
           "gather_rows");
  return out;
}

}  // namespace

void dvgo_bind_fused(pybind11::module_& m);  // fused_binding.cpp

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "directvoxgo_b200: sm_100a kernels behind the DirectVoxGO operator surface";
  m.def("abi_version", []() { return dvgo_abi_version(); });
  m.def("build_arch", []() { return std::string(dvgo_build_arch()); });
  m.def("launch_count", []() { return dvgo_launch_count(); });

  auto ru = m.def_submodule("render_utils_cuda");
  ru.def("infer_t_minmax", &infer_t_minmax, "Inference t_min and t_max of ray-bbox intersection");
  ru.def("infer_n_samples", &infer_n_samples, "Inference the number of points to sample on each ray");
  ru.def("infer_ray_start_dir", &infer_ray_start_dir, "Inference the starting point and shooting direction of each ray");
  ru.def("sample_pts_on_rays", &sample_pts_on_rays, "Sample points on rays");
  ru.def("sample_ndc_pts_on_rays", &sample_ndc_pts_on_rays, "Sample points on rays");
  ru.def("maskcache_lookup", &maskcache_lookup, "Lookup to skip know freespace.");
  ru.def("raw2alpha", &raw2alpha, "Raw values [-inf, inf] to alpha [0, 1].");
  ru.def("raw2alpha_backward", &raw2alpha_backward, "Backward pass of the raw to alpha");
  ru.def("alpha2weight", &alpha2weight, "Per-point alpha to accumulated blending weight");
  ru.def("alpha2weight_backward", &alpha2weight_backward, "Backward pass of alpha2weight");

  auto tv = m.def_submodule("total_variation_cuda");
  tv.def("total_variation_add_grad", &total_variation_add_grad, "Add total variation grad");

  auto ad = m.def_submodule("adam_upd_cuda");
  ad.def("adam_upd", &adam_upd, "Adam update");
  ad.def("masked_adam_upd", &masked_adam_upd, "Adam update ignoring zero grad");
  ad.def("adam_upd_with_perlr", &adam_upd_with_perlr, "Adam update ignoring zero grad with per-voxel lr");

  auto ext = m.def_submodule("ext");
  ext.def("grid_sample_3d", &grid_sample_3d);
  ext.def("grid_sample_3d_backward", &grid_sample_3d_backward);
  ext.def("grid_sample_2d", &grid_sample_2d);
  ext.def("grid_sample_2d_backward", &grid_sample_2d_backward);
  ext.def("grid_sample_norm", &grid_sample_norm);
  ext.def("grid_sample_norm_backward", &grid_sample_norm_backward);
  ext.def("segment_coo_sum", &segment_coo_sum);
  ext.def("gather_rows", &gather_rows);
  dvgo_bind_fused(ext);
}
