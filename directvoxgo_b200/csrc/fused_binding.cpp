// fused_binding.cpp -- torch adaptor for the fused entry points of include/dvgo_b200_fused.h.
// Every buffer is a preallocated torch tensor owned by the Python FusedTrainer / FusedRenderer; the
// functions below only validate (CUDA, contiguous, dtype), pass raw pointers + the current stream
// and turn non-zero return codes into RuntimeError.  No allocation, no synchronisation.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include "../../include/dvgo_b200_fused.h"
#include "../../include/dvgo_b200_prep.h"

namespace {

using torch::Tensor;

inline dvgo_stream_t cur_stream() {
  return reinterpret_cast<dvgo_stream_t>(at::cuda::getCurrentCUDAStream().stream());
}

inline void chk(const Tensor& t, const char* name, c10::ScalarType st) {
  TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor");
  TORCH_CHECK(t.is_contiguous(), name, " must be contiguous");
  TORCH_CHECK(t.scalar_type() == st, name, " has the wrong dtype");
}
#define F32(t) chk(t, #t, torch::kFloat32)
#define I32(t) chk(t, #t, torch::kInt32)

inline void rc_check(int rc, const char* what) {
  TORCH_CHECK(rc == 0, "dvgo_b200 fused: ", what, " failed with code ", rc,
              rc > 0 ? (std::string(" (") + cudaGetErrorString(static_cast<cudaError_t>(rc)) + ")")
                     : std::string(" (invalid argument)"));
}

inline const float* fp(const Tensor& t) { return t.data_ptr<float>(); }
inline float* fpm(const Tensor& t) { return t.data_ptr<float>(); }
inline int32_t* ipm(const Tensor& t) { return t.data_ptr<int32_t>(); }
inline const float* fp_opt(const c10::optional<Tensor>& t) { return t.has_value() ? t->data_ptr<float>() : nullptr; }

// Scene constants + the tensors they point into (kept alive by the object).
struct Scene {
  dvgo_scene_t s;
  Tensor xyz_min, xyz_max, mask, mask_scale, mask_shift;
  Scene(int X, int Y, int Z, int C, Tensor xyz_min_, Tensor xyz_max_, c10::optional<Tensor> mask_,
        c10::optional<Tensor> mask_scale_, c10::optional<Tensor> mask_shift_, double near, double far,
        double stepdist, double act_shift, double interval, double thres, bool ndc, int ndc_samples,
        bool exact_transmittance)
      : xyz_min(xyz_min_), xyz_max(xyz_max_) {
    F32(xyz_min); F32(xyz_max);
    s.X = X; s.Y = Y; s.Z = Z; s.C = C;
    s.xyz_min = fp(xyz_min); s.xyz_max = fp(xyz_max);
    s.mask = nullptr; s.mx = s.my = s.mz = 0; s.mask_scale = s.mask_shift = nullptr;
    if (mask_.has_value()) {
      mask = *mask_; mask_scale = *mask_scale_; mask_shift = *mask_shift_;
      chk(mask, "mask", torch::kBool); F32(mask_scale); F32(mask_shift);
      TORCH_CHECK(mask.dim() == 3, "mask must be [X,Y,Z]");
      s.mask = reinterpret_cast<const uint8_t*>(mask.data_ptr<bool>());
      s.mx = mask.size(0); s.my = mask.size(1); s.mz = mask.size(2);
      s.mask_scale = fp(mask_scale); s.mask_shift = fp(mask_shift);
    }
    s.near = static_cast<float>(near); s.far = static_cast<float>(far);
    s.stepdist = static_cast<float>(stepdist); s.act_shift = static_cast<float>(act_shift);
    s.interval = static_cast<float>(interval); s.fast_color_thres = static_cast<float>(thres);
    s.ndc = ndc ? 1 : 0; s.ndc_samples = ndc_samples;
    s.exact_transmittance = exact_transmittance ? 1 : 0;
  }
  int max_steps() const { return dvgo_fused_max_steps(&s); }
};

void ray_setup(const Scene& sc, Tensor rays_o, Tensor rays_d, Tensor t_min, Tensor n_steps, Tensor ray_off) {
  F32(rays_o); F32(rays_d); F32(t_min); I32(n_steps); I32(ray_off);
  const int n = rays_o.size(0);
  TORCH_CHECK(ray_off.numel() >= n + 1 && t_min.numel() >= n && n_steps.numel() >= n, "workspace too small");
  const c10::cuda::CUDAGuard guard(rays_o.device());
  rc_check(dvgo_fused_ray_setup(fp(rays_o), fp(rays_d), &sc.s, n, fpm(t_min), ipm(n_steps), ipm(ray_off),
                                cur_stream()), "ray_setup");
}

void march_fwd(const Scene& sc, Tensor rays_o, Tensor rays_d, Tensor density, c10::optional<Tensor> k0_cl,
               Tensor t_min, Tensor n_steps, Tensor ray_off, Tensor slot_alpha, Tensor slot_T,
               Tensor slot_expd, Tensor slot_code, Tensor feat, Tensor s_ray, Tensor s_slot, Tensor s_weight,
               Tensor alphainv_last, Tensor counters, c10::optional<Tensor> s_pos) {
  F32(rays_o); F32(rays_d); F32(density); F32(t_min); I32(n_steps); I32(ray_off); F32(slot_alpha);
  F32(slot_T); F32(slot_expd); I32(slot_code); F32(feat); I32(s_ray); I32(s_slot); F32(s_weight);
  F32(alphainv_last); I32(counters);
  const bool keep_slots = slot_code.numel() > 0;   // empty slot arrays = forward-only (rendering): no per-slot record
  if (k0_cl.has_value()) { F32((*k0_cl)); TORCH_CHECK(k0_cl->numel() == density.numel() * sc.s.C, "k0_cl must be [X,Y,Z,C]"); }
  TORCH_CHECK(density.numel() == (int64_t)sc.s.X * sc.s.Y * sc.s.Z, "density must be [X,Y,Z]");
  const int n = rays_o.size(0);
  const int64_t slot_cap = slot_code.numel(), surv_cap = s_ray.numel();
  TORCH_CHECK(slot_alpha.numel() >= slot_cap && slot_T.numel() >= slot_cap && slot_expd.numel() >= slot_cap, "slot arrays");
  TORCH_CHECK(s_slot.numel() >= surv_cap && s_weight.numel() >= surv_cap && feat.numel() >= surv_cap * sc.s.C, "survivor arrays");
  TORCH_CHECK(alphainv_last.numel() >= n && counters.numel() >= 2, "per-ray arrays");
  if (s_pos.has_value()) { F32((*s_pos)); TORCH_CHECK(s_pos->numel() >= surv_cap * 4, "s_pos must be [surv_cap,4]"); }
  const c10::cuda::CUDAGuard guard(rays_o.device());
  rc_check(dvgo_fused_march_fwd(fp(rays_o), fp(rays_d), &sc.s, fp(density), fp_opt(k0_cl), n, fp(t_min),
                                ipm(n_steps), ipm(ray_off), slot_cap, surv_cap, keep_slots ? fpm(slot_alpha) : nullptr,
                                keep_slots ? fpm(slot_T) : nullptr, keep_slots ? fpm(slot_expd) : nullptr,
                                keep_slots ? ipm(slot_code) : nullptr, fpm(feat), ipm(s_ray), ipm(s_slot),
                                fpm(s_weight), fpm(alphainv_last), ipm(counters),
                                s_pos.has_value() ? s_pos->data_ptr<float>() : nullptr, cur_stream()), "march_fwd");
}

void k0_gather(const Scene& sc, Tensor rays_o, Tensor rays_d, Tensor k0_cl, Tensor t_min, Tensor ray_off, Tensor s_ray,
               Tensor s_slot, Tensor counters, Tensor feat) {
  F32(rays_o); F32(rays_d); F32(k0_cl); F32(t_min); I32(ray_off); I32(s_ray); I32(s_slot); I32(counters); F32(feat);
  const c10::cuda::CUDAGuard guard(rays_o.device());
  rc_check(dvgo_fused_k0_gather(fp(rays_o), fp(rays_d), &sc.s, fp(k0_cl), fp(t_min), ipm(ray_off), ipm(s_ray),
                                ipm(s_slot), ipm(counters), s_ray.numel(), fpm(feat), cur_stream()), "k0_gather");
}

void k0_gather_tiles(const Scene& sc, Tensor rays_o, Tensor rays_d, Tensor k0_cl, Tensor t_min, Tensor ray_off,
                     Tensor s_ray, Tensor s_slot, Tensor counters, c10::optional<Tensor> s_pos, Tensor pe16, int pe_stride,
                     Tensor xt) {
  if (s_pos.has_value()) { F32((*s_pos)); TORCH_CHECK(s_pos->numel() >= s_ray.numel() * 4, "s_pos must be [surv_cap,4]"); }
  F32(rays_o); F32(rays_d); F32(k0_cl); F32(t_min); I32(ray_off); I32(s_ray); I32(s_slot); I32(counters);
  TORCH_CHECK(pe16.is_cuda() && pe16.is_contiguous() && pe16.scalar_type() == torch::kHalf && pe16.dim() == 2 &&
              pe16.size(0) >= rays_o.size(0) && pe16.size(1) == ((sc.s.C + pe_stride + 15) / 16) * 16,
              "pe16 must be the [n_rays, K1] half table view_embedding(..., C) returns");
  TORCH_CHECK(xt.is_cuda() && xt.is_contiguous() && xt.scalar_type() == torch::kUInt8 &&
              xt.numel() >= dvgo_mlp_xtile_bytes(s_ray.numel(), sc.s.C, pe_stride) &&
              reinterpret_cast<uintptr_t>(xt.data_ptr()) % 16 == 0, "xt: uint8 CUDA tensor of mlp_xtile_bytes bytes");
  const c10::cuda::CUDAGuard guard(rays_o.device());
  rc_check(dvgo_fused_k0_gather_tiles(fp(rays_o), fp(rays_d), &sc.s, fp(k0_cl), fp(t_min), ipm(ray_off), ipm(s_ray),
                                      ipm(s_slot), ipm(counters), s_ray.numel(),
                                      s_pos.has_value() ? s_pos->data_ptr<float>() : nullptr, pe16.data_ptr(), pe_stride,
                                      xt.data_ptr(), cur_stream()), "k0_gather_tiles");
}

// k0 gather + rgbnet forward in one kernel (mlp_fwd_gather_kernel): rgb [surv_cap,3]; xt (training): the X~ tiles the
// backward kernel reads, written by one bulk copy per tile; None (rendering): the tiles never leave shared memory.
void mlp_fwd_gather(const Scene& sc, Tensor k0_cl, Tensor s_pos, Tensor pe16, int P, int pe_stride, Tensor counters,
                    int64_t cap, Tensor wpack, Tensor rgb, c10::optional<Tensor> xt) {
  F32(k0_cl); F32(s_pos); I32(counters); F32(rgb);
  TORCH_CHECK(s_pos.numel() >= cap * 4, "s_pos must be [surv_cap,4]");
  TORCH_CHECK(rgb.numel() >= cap * 3, "rgb too small");
  TORCH_CHECK(pe16.is_cuda() && pe16.is_contiguous() && pe16.scalar_type() == torch::kHalf && pe16.dim() == 2 &&
              pe16.size(1) == ((sc.s.C + pe_stride + 15) / 16) * 16,
              "pe16 must be the [n_rays, K1] half table view_embedding(..., C) returns");
  TORCH_CHECK(wpack.is_cuda() && wpack.is_contiguous() && wpack.scalar_type() == torch::kUInt8 &&
              wpack.numel() >= dvgo_mlp_wpack_bytes(sc.s.C, pe_stride) &&
              reinterpret_cast<uintptr_t>(wpack.data_ptr()) % 16 == 0, "wpack: uint8 CUDA tensor of mlp_wpack_bytes bytes");
  if (xt.has_value())
    TORCH_CHECK(xt->is_cuda() && xt->is_contiguous() && xt->scalar_type() == torch::kUInt8 &&
                xt->numel() >= dvgo_mlp_xtile_bytes(cap, sc.s.C, pe_stride) &&
                reinterpret_cast<uintptr_t>(xt->data_ptr()) % 16 == 0, "xt: uint8 CUDA tensor of mlp_xtile_bytes bytes");
  const c10::cuda::CUDAGuard guard(k0_cl.device());
  rc_check(dvgo_mlp_fwd_gather(&sc.s, fp(k0_cl), fp(s_pos), pe16.data_ptr(), P, pe_stride, ipm(counters), cap,
                               wpack.data_ptr(), fpm(rgb), xt.has_value() ? xt->data_ptr() : nullptr, cur_stream()),
           "mlp_fwd_gather");
}

void k0_scatter(const Scene& sc, Tensor rays_o, Tensor rays_d, Tensor t_min, Tensor ray_off, Tensor s_ray, Tensor s_slot,
                Tensor counters, Tensor d_feat, Tensor grad_k0_cl, c10::optional<Tensor> s_pos) {
  F32(rays_o); F32(rays_d); F32(t_min); I32(ray_off); I32(s_ray); I32(s_slot); I32(counters); F32(d_feat);
  F32(grad_k0_cl);
  if (s_pos.has_value()) { F32((*s_pos)); TORCH_CHECK(s_pos->numel() >= s_ray.numel() * 4, "s_pos must be [surv_cap,4]"); }
  const c10::cuda::CUDAGuard guard(rays_o.device());
  rc_check(dvgo_fused_k0_scatter(fp(rays_o), fp(rays_d), &sc.s, fp(t_min), ipm(ray_off), ipm(s_ray), ipm(s_slot),
                                 ipm(counters), s_ray.numel(), s_pos.has_value() ? s_pos->data_ptr<float>() : nullptr,
                                 fp(d_feat), fpm(grad_k0_cl), cur_stream()), "k0_scatter");
}

void rgb_direct(Tensor feat, Tensor counters, Tensor rgb) {
  F32(feat); I32(counters); F32(rgb);
  const c10::cuda::CUDAGuard guard(feat.device());
  rc_check(dvgo_fused_rgb_direct(fp(feat), ipm(counters), rgb.numel() / 3, fpm(rgb), cur_stream()), "rgb_direct");
}

void rgb_direct_bwd(Tensor rgb, Tensor d_rgb, Tensor counters, Tensor d_feat) {
  F32(rgb); F32(d_rgb); I32(counters); F32(d_feat);
  const c10::cuda::CUDAGuard guard(rgb.device());
  rc_check(dvgo_fused_rgb_direct_bwd(fp(rgb), fp(d_rgb), ipm(counters), rgb.numel() / 3, fpm(d_feat), cur_stream()),
           "rgb_direct_bwd");
}

void composite(Tensor rgb, Tensor s_weight, Tensor s_ray, Tensor s_slot, Tensor ray_off, Tensor counters,
               Tensor rgb_acc, c10::optional<Tensor> depth_acc) {
  F32(rgb); F32(s_weight); I32(s_ray); I32(s_slot); I32(ray_off); I32(counters); F32(rgb_acc);
  if (depth_acc.has_value()) F32((*depth_acc));
  const c10::cuda::CUDAGuard guard(rgb.device());
  rc_check(dvgo_fused_composite(fp(rgb), fp(s_weight), ipm(s_ray), ipm(s_slot), ipm(ray_off), ipm(counters),
                                rgb.numel() / 3, fpm(rgb_acc),
                                depth_acc.has_value() ? depth_acc->data_ptr<float>() : nullptr, cur_stream()),
           "composite");
}

void ray_finish(Tensor rgb_acc, Tensor alphainv_last, c10::optional<Tensor> target, double bg, int n_rays,
                int n_global, double weight_main, double weight_entropy_last, c10::optional<Tensor> G,
                c10::optional<Tensor> g_last, c10::optional<Tensor> loss_acc) {
  F32(rgb_acc); F32(alphainv_last);
  if (target.has_value()) { F32((*target)); F32((*G)); F32((*g_last)); }
  const c10::cuda::CUDAGuard guard(rgb_acc.device());
  rc_check(dvgo_fused_ray_finish(fpm(rgb_acc), fp(alphainv_last), fp_opt(target), static_cast<float>(bg), n_rays,
                                 n_global, static_cast<float>(weight_main), static_cast<float>(weight_entropy_last),
                                 G.has_value() ? G->data_ptr<float>() : nullptr,
                                 g_last.has_value() ? g_last->data_ptr<float>() : nullptr,
                                 loss_acc.has_value() ? loss_acc->data_ptr<float>() : nullptr, cur_stream()),
           "ray_finish");
}

void sample_grad(Tensor rgb, Tensor s_weight, Tensor s_ray, Tensor G, Tensor target, Tensor counters, int n_global,
                 double weight_rgbper, c10::optional<Tensor> d_rgb, Tensor d_w, Tensor loss_acc, c10::optional<Tensor> dzt,
                 double grad_scale) {
  F32(rgb); F32(s_weight); I32(s_ray); F32(G); F32(target); I32(counters); F32(d_w); F32(loss_acc);
  if (d_rgb.has_value()) F32((*d_rgb));
  TORCH_CHECK(d_rgb.has_value() || dzt.has_value(), "sample_grad: d_rgb may only be omitted when dzt is given");
  if (dzt.has_value())
    TORCH_CHECK(dzt->is_cuda() && dzt->is_contiguous() && dzt->scalar_type() == torch::kUInt8 &&
                dzt->numel() >= dvgo_mlp_dztile_bytes(rgb.numel() / 3) &&
                reinterpret_cast<uintptr_t>(dzt->data_ptr()) % 16 == 0, "dzt: uint8 CUDA tensor of mlp_dztile_bytes bytes");
  const c10::cuda::CUDAGuard guard(rgb.device());
  rc_check(dvgo_fused_sample_grad(fp(rgb), fp(s_weight), ipm(s_ray), fp(G), fp(target), ipm(counters),
                                  rgb.numel() / 3, n_global, static_cast<float>(weight_rgbper),
                                  d_rgb.has_value() ? d_rgb->data_ptr<float>() : nullptr, fpm(d_w),
                                  fpm(loss_acc), dzt.has_value() ? dzt->data_ptr() : nullptr,
                                  static_cast<float>(grad_scale), cur_stream()),
           "sample_grad");
}

void march_bwd(const Scene& sc, Tensor rays_o, Tensor rays_d, Tensor t_min, Tensor n_steps, Tensor ray_off,
               Tensor slot_alpha, Tensor slot_T, Tensor slot_expd, Tensor slot_code, Tensor d_feat, Tensor d_w,
               Tensor alphainv_last, Tensor g_last, Tensor grad_density, c10::optional<Tensor> grad_k0_cl) {
  F32(rays_o); F32(rays_d); F32(t_min); I32(n_steps); I32(ray_off); F32(slot_alpha); F32(slot_T); F32(slot_expd);
  I32(slot_code); F32(d_feat); F32(d_w); F32(alphainv_last); F32(g_last); F32(grad_density);
  if (grad_k0_cl.has_value()) F32((*grad_k0_cl));
  const c10::cuda::CUDAGuard guard(rays_o.device());
  rc_check(dvgo_fused_march_bwd(fp(rays_o), fp(rays_d), &sc.s, rays_o.size(0), fp(t_min), ipm(n_steps), ipm(ray_off),
                                fp(slot_alpha), fp(slot_T), fp(slot_expd), ipm(slot_code), fp(d_feat), fp(d_w),
                                fp(alphainv_last), fp(g_last), fpm(grad_density),
                                grad_k0_cl.has_value() ? grad_k0_cl->data_ptr<float>() : nullptr, cur_stream()),
           "march_bwd");
}

void sweep(Tensor param_in, Tensor param_out, Tensor grad, Tensor exp_avg, Tensor exp_avg_sq,
           c10::optional<Tensor> perlr, int X, int Y, int Z, int C, bool tv, bool tv_dense, double wx, double wy,
           double wz, bool masked, int step, double beta1, double beta2, double lr, double eps, int x_begin,
           int x_end) {
  F32(param_in); F32(param_out); F32(grad); F32(exp_avg); F32(exp_avg_sq);
  const int64_t n = (int64_t)X * Y * Z * C;
  TORCH_CHECK(param_in.numel() == n && param_out.numel() == n && grad.numel() == n && exp_avg.numel() == n &&
                  exp_avg_sq.numel() == n, "sweep: all buffers must have X*Y*Z*C elements");
  const c10::cuda::CUDAGuard guard(param_in.device());
  rc_check(dvgo_fused_sweep(fp(param_in), fpm(param_out), fpm(grad), fpm(exp_avg), fpm(exp_avg_sq), fp_opt(perlr), X,
                            Y, Z, C, x_begin, x_end, tv, tv_dense, static_cast<float>(wx), static_cast<float>(wy),
                            static_cast<float>(wz), masked, step, static_cast<float>(beta1),
                            static_cast<float>(beta2), static_cast<float>(lr), static_cast<float>(eps), cur_stream()),
           "sweep");
}

// The sweep with the gradient exchange folded in (include/dvgo_b200_fused.h: dvgo_fused_sweep_peer).  The peer lists
// hold raw device addresses of every rank's buffer as mapped in THIS process (torch symmetric memory); param_in,
// grad_local, exp_avg, exp_avg_sq are this rank's tensors (grad_local / the local output are only size-checked).
void sweep_peer(Tensor param_in, std::vector<int64_t> param_out_ptrs, std::vector<int64_t> grad_ptrs,
                int64_t param_out_mc, int64_t grad_mc, int self_rank,
                Tensor grad_local, Tensor exp_avg, Tensor exp_avg_sq, c10::optional<Tensor> perlr, int X, int Y, int Z,
                int C, bool tv, bool tv_dense, double wx, double wy, double wz, bool masked, int step, double beta1,
                double beta2, double lr, double eps, int x_begin, int x_end) {
  F32(param_in); F32(grad_local); F32(exp_avg); F32(exp_avg_sq);
  const int64_t n = (int64_t)X * Y * Z * C;
  const int np = static_cast<int>(grad_ptrs.size());
  TORCH_CHECK(np >= 1 && np <= 8 && param_out_ptrs.size() == grad_ptrs.size(), "sweep_peer: 1..8 peers");
  TORCH_CHECK(self_rank >= 0 && self_rank < np, "sweep_peer: bad rank");
  TORCH_CHECK(param_in.numel() == n && grad_local.numel() == n && exp_avg.numel() == n && exp_avg_sq.numel() == n,
              "sweep_peer: all buffers must have X*Y*Z*C elements");
  TORCH_CHECK(reinterpret_cast<int64_t>(grad_local.data_ptr<float>()) == grad_ptrs[self_rank],
              "sweep_peer: grad_ptrs[self_rank] must be the local gradient buffer");
  float* pout[8];
  float* grads[8];
  for (int r = 0; r < np; ++r) {
    pout[r] = reinterpret_cast<float*>(param_out_ptrs[r]);
    grads[r] = reinterpret_cast<float*>(grad_ptrs[r]);
  }
  const c10::cuda::CUDAGuard guard(param_in.device());
  rc_check(dvgo_fused_sweep_peer(fp(param_in), pout, grads, reinterpret_cast<float*>(param_out_mc),
                                 reinterpret_cast<const float*>(grad_mc), np, self_rank, fpm(exp_avg), fpm(exp_avg_sq), fp_opt(perlr),
                                 X, Y, Z, C, x_begin, x_end, tv, tv_dense, static_cast<float>(wx),
                                 static_cast<float>(wy), static_cast<float>(wz), masked, step,
                                 static_cast<float>(beta1), static_cast<float>(beta2), static_cast<float>(lr),
                                 static_cast<float>(eps), cur_stream()), "sweep_peer");
}

void peer_barrier(std::vector<int64_t> flag_ptrs, int self_rank, int epoch) {
  const int np = static_cast<int>(flag_ptrs.size());
  TORCH_CHECK(np >= 1 && np <= 8 && self_rank >= 0 && self_rank < np, "peer_barrier: 1..8 peers");
  int32_t* flags[8];
  for (int r = 0; r < np; ++r) flags[r] = reinterpret_cast<int32_t*>(flag_ptrs[r]);
  rc_check(dvgo_peer_barrier(flags, np, self_rank, epoch, cur_stream()), "peer_barrier");
}

void adam_upd_peer(Tensor param, std::vector<int64_t> grad_ptrs, int64_t grad_mc, Tensor exp_avg, Tensor exp_avg_sq,
                   int step, double beta1, double beta2, double lr, double eps) {
  F32(param); F32(exp_avg); F32(exp_avg_sq);
  const int np = static_cast<int>(grad_ptrs.size());
  TORCH_CHECK(np >= 1 && np <= 8, "adam_upd_peer: 1..8 peers");
  TORCH_CHECK(exp_avg.numel() == param.numel() && exp_avg_sq.numel() == param.numel(), "adam_upd_peer: shape mismatch");
  const float* grads[8];
  for (int r = 0; r < np; ++r) grads[r] = reinterpret_cast<const float*>(grad_ptrs[r]);
  const c10::cuda::CUDAGuard guard(param.device());
  rc_check(dvgo_adam_upd_peer(fpm(param), grads, reinterpret_cast<const float*>(grad_mc), np, fpm(exp_avg), fpm(exp_avg_sq),
                              param.numel(), step, static_cast<float>(beta1), static_cast<float>(beta2),
                              static_cast<float>(lr), static_cast<float>(eps), cur_stream()), "adam_upd_peer");
}

Tensor ncdhw_to_cl(Tensor src) {  // [1,C,X,Y,Z] -> [X,Y,Z,C]
  F32(src);
  TORCH_CHECK(src.dim() == 5 && src.size(0) == 1, "expected [1,C,X,Y,Z]");
  const c10::cuda::CUDAGuard guard(src.device());
  const int C = src.size(1);
  auto dst = torch::empty({src.size(2), src.size(3), src.size(4), C}, src.options());
  rc_check(dvgo_grid_ncdhw_to_cl(fp(src), fpm(dst), C, src.numel() / std::max(C, 1), cur_stream()), "ncdhw_to_cl");
  return dst;
}

Tensor cl_to_ncdhw(Tensor src) {  // [X,Y,Z,C] -> [1,C,X,Y,Z]
  F32(src);
  TORCH_CHECK(src.dim() == 4, "expected [X,Y,Z,C]");
  const c10::cuda::CUDAGuard guard(src.device());
  const int C = src.size(3);
  auto dst = torch::empty({1, C, src.size(0), src.size(1), src.size(2)}, src.options());
  rc_check(dvgo_grid_cl_to_ncdhw(fp(src), fpm(dst), C, src.numel() / std::max(C, 1), cur_stream()), "cl_to_ncdhw");
  return dst;
}

void zero_(Tensor t) {
  TORCH_CHECK(t.is_cuda() && t.is_contiguous() && t.element_size() == 4, "zero_: 4-byte contiguous CUDA tensor");
  const c10::cuda::CUDAGuard guard(t.device());
  rc_check(dvgo_fused_zero(t.data_ptr(), t.numel(), cur_stream()), "zero");
}

void step_begin(Tensor zblock, c10::optional<Tensor> stats) {
  TORCH_CHECK(zblock.is_cuda() && zblock.is_contiguous() && zblock.element_size() == 4 && zblock.numel() >= 2,
              "step_begin: 4-byte contiguous CUDA tensor with >= 2 words");
  long long* st = nullptr;
  if (stats.has_value()) {
    TORCH_CHECK(stats->is_cuda() && stats->is_contiguous() && stats->scalar_type() == torch::kInt64 && stats->numel() >= 4,
                "step_begin: stats must be an int64 CUDA tensor with 4 elements");
    st = reinterpret_cast<long long*>(stats->data_ptr<int64_t>());
  }
  const c10::cuda::CUDAGuard guard(zblock.device());
  rc_check(dvgo_fused_step_begin(zblock.data_ptr(), zblock.numel(), st, cur_stream()), "step_begin");
}

// ---- include/dvgo_b200_prep.h: ray generation, training-ray preparation, whole-grid sweeps -------------------
struct View {
  dvgo_view_t v;
  View(int H, int W, double fx, double fy, double cx, double cy, std::vector<double> c2w, bool inverse_y, bool flip_x,
       bool flip_y, int mode, bool ndc, double ndc_sx, double ndc_sy) {
    TORCH_CHECK(c2w.size() == 12, "c2w: 12 values (rows 0..2 of the 3x4 / 4x4 matrix)");
    TORCH_CHECK(mode == 0 || mode == 1, "mode: 0 = lefttop, 1 = center");
    v.H = H; v.W = W;
    v.fx = static_cast<float>(fx); v.fy = static_cast<float>(fy); v.cx = static_cast<float>(cx); v.cy = static_cast<float>(cy);
    for (int i = 0; i < 12; ++i) v.c2w[i] = static_cast<float>(c2w[i]);
    v.inverse_y = inverse_y; v.flip_x = flip_x; v.flip_y = flip_y; v.mode = mode; v.ndc = ndc;
    v.ndc_sx = static_cast<float>(ndc_sx); v.ndc_sy = static_cast<float>(ndc_sy);
  }
};

void rays_of_view(const View& v, int64_t pix_begin, int64_t n_pix, c10::optional<Tensor> rays_o,
                  c10::optional<Tensor> rays_d, c10::optional<Tensor> viewdirs) {
  const Tensor* any = nullptr;
  for (auto* t : {&rays_o, &rays_d, &viewdirs})
    if (t->has_value()) {
      chk(**t, "rays_of_view output", torch::kFloat32);
      TORCH_CHECK((*t)->numel() >= 3 * n_pix, "rays_of_view: output too small");
      any = &**t;
    }
  TORCH_CHECK(any, "rays_of_view: no output given");
  const c10::cuda::CUDAGuard guard(any->device());
  rc_check(dvgo_rays_of_view(&v.v, pix_begin, n_pix, rays_o.has_value() ? fpm(*rays_o) : nullptr,
                             rays_d.has_value() ? fpm(*rays_d) : nullptr,
                             viewdirs.has_value() ? fpm(*viewdirs) : nullptr, cur_stream()), "rays_of_view");
}

Tensor hit_coarse_geo(const Scene& sc, Tensor rays_o, Tensor rays_d) {
  F32(rays_o); F32(rays_d);
  TORCH_CHECK(rays_o.dim() == 2 && rays_o.size(1) == 3 && rays_d.sizes() == rays_o.sizes(), "rays must be [N,3]");
  const c10::cuda::CUDAGuard guard(rays_o.device());
  auto hit = torch::empty({rays_o.size(0)}, rays_o.options().dtype(torch::kBool));
  rc_check(dvgo_hit_coarse_geo(&sc.s, fp(rays_o), fp(rays_d), rays_o.size(0),
                               reinterpret_cast<uint8_t*>(hit.data_ptr<bool>()), cur_stream()), "hit_coarse_geo");
  return hit;
}

Tensor view_hit_coarse_geo(const View& v, const Scene& sc) {
  const c10::cuda::CUDAGuard guard(sc.xyz_min.device());
  auto hit = torch::empty({v.v.H, v.v.W}, sc.xyz_min.options().dtype(torch::kBool));
  rc_check(dvgo_view_hit_coarse_geo(&v.v, &sc.s, reinterpret_cast<uint8_t*>(hit.data_ptr<bool>()), cur_stream()),
           "view_hit_coarse_geo");
  return hit;
}

void view_gather_rays(const View& v, Tensor hit, Tensor pos_incl, Tensor top, c10::optional<Tensor> img, Tensor rgb_tr,
                      Tensor rays_o_tr, Tensor rays_d_tr, Tensor viewdirs_tr) {
  chk(hit, "hit", torch::kBool); chk(pos_incl, "pos_incl", torch::kInt64); chk(top, "top", torch::kInt64);
  F32(rgb_tr); F32(rays_o_tr); F32(rays_d_tr); F32(viewdirs_tr);
  const int64_t n = static_cast<int64_t>(v.v.H) * v.v.W;
  TORCH_CHECK(hit.numel() == n && pos_incl.numel() == n && top.numel() == 1, "view_gather_rays: size mismatch");
  if (img.has_value()) { chk(*img, "img", torch::kFloat32); TORCH_CHECK(img->numel() == 3 * n, "img must be [H,W,3]"); }
  const c10::cuda::CUDAGuard guard(hit.device());
  rc_check(dvgo_view_gather_rays(&v.v, reinterpret_cast<const uint8_t*>(hit.data_ptr<bool>()),
                                 pos_incl.data_ptr<int64_t>(), top.data_ptr<int64_t>(),
                                 img.has_value() ? fp(*img) : nullptr, fpm(rgb_tr), fpm(rays_o_tr), fpm(rays_d_tr),
                                 fpm(viewdirs_tr), cur_stream()), "view_gather_rays");
}

void voxel_count_scatter(Tensor rays_o, Tensor rays_d, Tensor xyz_min, Tensor xyz_max, double near, double far,
                         double stepdist, int n_samples, Tensor acc) {
  F32(rays_o); F32(rays_d); F32(xyz_min); F32(xyz_max); F32(acc);
  TORCH_CHECK(acc.dim() == 3, "acc must be [X,Y,Z]");
  const c10::cuda::CUDAGuard guard(acc.device());
  rc_check(dvgo_voxel_count_scatter(fp(rays_o), fp(rays_d), rays_o.numel() / 3, fp(xyz_min), fp(xyz_max), acc.size(0),
                                    acc.size(1), acc.size(2), static_cast<float>(near), static_cast<float>(far),
                                    static_cast<float>(stepdist), n_samples, fpm(acc), cur_stream()), "voxel_count_scatter");
}

void voxel_count_commit(Tensor acc, Tensor count) {
  F32(acc); F32(count);
  TORCH_CHECK(acc.numel() == count.numel(), "voxel_count_commit: size mismatch");
  const c10::cuda::CUDAGuard guard(acc.device());
  rc_check(dvgo_voxel_count_commit(fpm(acc), fpm(count), acc.numel(), cur_stream()), "voxel_count_commit");
}

Tensor alpha_maxpool_mask(Tensor density, double act_shift, double interval, double thres, c10::optional<Tensor> mask_in) {
  F32(density);
  TORCH_CHECK(density.dim() >= 3, "density must be [...,X,Y,Z]");
  const int d = density.dim();
  const int X = density.size(d - 3), Y = density.size(d - 2), Z = density.size(d - 1);
  TORCH_CHECK(density.numel() == static_cast<int64_t>(X) * Y * Z, "density must have one channel");
  const c10::cuda::CUDAGuard guard(density.device());
  if (mask_in.has_value()) { chk(*mask_in, "mask_in", torch::kBool); TORCH_CHECK(mask_in->numel() == density.numel(), "mask size"); }
  auto out = torch::empty({X, Y, Z}, density.options().dtype(torch::kBool));
  auto tmp = torch::empty({X, Y, Z}, density.options());
  rc_check(dvgo_alpha_maxpool_mask(fp(density), X, Y, Z, static_cast<float>(act_shift), static_cast<float>(interval),
                                   static_cast<float>(thres),
                                   mask_in.has_value() ? reinterpret_cast<const uint8_t*>(mask_in->data_ptr<bool>()) : nullptr,
                                   reinterpret_cast<uint8_t*>(out.data_ptr<bool>()), fpm(tmp), cur_stream()),
           "alpha_maxpool_mask");
  return out;
}

Tensor resize_trilinear(Tensor src, int X2, int Y2, int Z2) {
  F32(src);
  TORCH_CHECK(src.dim() == 5 && src.size(0) == 1, "src must be [1,C,X,Y,Z]");
  const c10::cuda::CUDAGuard guard(src.device());
  auto dst = torch::empty({1, src.size(1), X2, Y2, Z2}, src.options());
  rc_check(dvgo_resize_trilinear(fp(src), src.size(1), src.size(2), src.size(3), src.size(4), fpm(dst), X2, Y2, Z2,
                                 cur_stream()), "resize_trilinear");
  return dst;
}

}  // namespace

void dvgo_bind_mlp(pybind11::module_& m);  // mlp_binding.cpp

void dvgo_bind_fused(pybind11::module_& m) {
  pybind11::class_<Scene>(m, "Scene")
      .def(pybind11::init<int, int, int, int, Tensor, Tensor, c10::optional<Tensor>, c10::optional<Tensor>,
                          c10::optional<Tensor>, double, double, double, double, double, double, bool, int, bool>(),
           pybind11::arg("X"), pybind11::arg("Y"), pybind11::arg("Z"), pybind11::arg("C"), pybind11::arg("xyz_min"),
           pybind11::arg("xyz_max"), pybind11::arg("mask"), pybind11::arg("mask_scale"), pybind11::arg("mask_shift"),
           pybind11::arg("near"), pybind11::arg("far"), pybind11::arg("stepdist"), pybind11::arg("act_shift"),
           pybind11::arg("interval"), pybind11::arg("thres"), pybind11::arg("ndc"), pybind11::arg("ndc_samples"),
           pybind11::arg("exact_transmittance") = false)
      .def("max_steps", &Scene::max_steps);
  m.def("ray_setup", &ray_setup);
  m.def("march_fwd", &march_fwd, pybind11::arg("scene"), pybind11::arg("rays_o"), pybind11::arg("rays_d"), pybind11::arg("density"), pybind11::arg("k0_cl"), pybind11::arg("t_min"), pybind11::arg("n_steps"), pybind11::arg("ray_off"), pybind11::arg("slot_alpha"), pybind11::arg("slot_T"), pybind11::arg("slot_expd"), pybind11::arg("slot_code"), pybind11::arg("feat"), pybind11::arg("s_ray"), pybind11::arg("s_slot"), pybind11::arg("s_weight"), pybind11::arg("alphainv_last"), pybind11::arg("counters"), pybind11::arg("s_pos") = pybind11::none());
  m.def("k0_gather", &k0_gather);
  m.def("k0_gather_tiles", &k0_gather_tiles);
  m.def("mlp_fwd_gather", &mlp_fwd_gather);
  m.def("k0_scatter", &k0_scatter, pybind11::arg("scene"), pybind11::arg("rays_o"), pybind11::arg("rays_d"), pybind11::arg("t_min"), pybind11::arg("ray_off"), pybind11::arg("s_ray"), pybind11::arg("s_slot"), pybind11::arg("counters"), pybind11::arg("d_feat"), pybind11::arg("grad_k0_cl"), pybind11::arg("s_pos") = pybind11::none());
  m.def("rgb_direct", &rgb_direct);
  m.def("rgb_direct_bwd", &rgb_direct_bwd);
  m.def("composite", &composite);
  m.def("ray_finish", &ray_finish);
  m.def("sample_grad", &sample_grad, pybind11::arg("rgb"), pybind11::arg("s_weight"), pybind11::arg("s_ray"),
        pybind11::arg("G"), pybind11::arg("target"), pybind11::arg("counters"), pybind11::arg("n_global"),
        pybind11::arg("weight_rgbper"), pybind11::arg("d_rgb"), pybind11::arg("d_w"), pybind11::arg("loss_acc"),
        pybind11::arg("dzt") = pybind11::none(), pybind11::arg("grad_scale") = 1.0);
  m.def("march_bwd", &march_bwd);
  m.def("sweep", &sweep, pybind11::arg("param_in"), pybind11::arg("param_out"), pybind11::arg("grad"),
        pybind11::arg("exp_avg"), pybind11::arg("exp_avg_sq"), pybind11::arg("perlr"), pybind11::arg("X"),
        pybind11::arg("Y"), pybind11::arg("Z"), pybind11::arg("C"), pybind11::arg("tv"), pybind11::arg("tv_dense"),
        pybind11::arg("wx"), pybind11::arg("wy"), pybind11::arg("wz"), pybind11::arg("masked"), pybind11::arg("step"),
        pybind11::arg("beta1"), pybind11::arg("beta2"), pybind11::arg("lr"), pybind11::arg("eps"),
        pybind11::arg("x_begin") = 0, pybind11::arg("x_end") = -1);
  m.def("sweep_peer", &sweep_peer);
  m.def("peer_barrier", &peer_barrier);
  m.def("adam_upd_peer", &adam_upd_peer);
  m.def("ncdhw_to_cl", &ncdhw_to_cl);
  m.def("cl_to_ncdhw", &cl_to_ncdhw);
  m.def("zero_", &zero_);
  m.def("step_begin", &step_begin);
  pybind11::class_<View>(m, "View")
      .def(pybind11::init<int, int, double, double, double, double, std::vector<double>, bool, bool, bool, int, bool,
                          double, double>());
  m.def("rays_of_view", &rays_of_view);
  m.def("hit_coarse_geo", &hit_coarse_geo);
  m.def("view_hit_coarse_geo", &view_hit_coarse_geo);
  m.def("view_gather_rays", &view_gather_rays);
  m.def("voxel_count_scatter", &voxel_count_scatter);
  m.def("voxel_count_commit", &voxel_count_commit);
  m.def("alpha_maxpool_mask", &alpha_maxpool_mask);
  m.def("resize_trilinear", &resize_trilinear);
  dvgo_bind_mlp(m);
}
