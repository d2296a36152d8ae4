"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of the float64 half of oracle/libdvgo_oracle.so
(oracle/dvgo_oracle_f64.c: the CPU restatement of the reference kernels' DOUBLE instantiation).

Same call surface as oracle/oracle.py (= the reference's pybind tables, lib/cuda/render_utils.cpp:144-155,
total_variation.cpp:22-24, adam_upd.cpp:79-86), on contiguous CPU float64 tensors.  Checks
directvoxgo_b200/csrc/f64_ops.cu; nothing in the product package imports this module.
"""
import torch

from .oracle import _F, _I, _L, _P, _p, lib

_D = torch.float64


def _f64(t):
    assert t.dtype == _D, "the float64 oracle takes float64 tensors"
    return t.detach().contiguous()


def infer_t_minmax(rays_o, rays_d, xyz_min, xyz_max, near, far):
    n = rays_o.shape[0]
    t_min, t_max = torch.empty(n, dtype=_D), torch.empty(n, dtype=_D)
    lib().orc64_infer_t_minmax(_p(_f64(rays_o)), _p(_f64(rays_d)), _p(_f64(xyz_min)), _p(_f64(xyz_max)),
                               _F(near), _F(far), _I(n), _p(t_min), _p(t_max))
    return [t_min, t_max]


def infer_n_samples(t_min, t_max, stepdist):
    n = t_min.shape[0]
    out = torch.empty(n, dtype=torch.int64)
    lib().orc64_infer_n_samples(_p(_f64(t_min)), _p(_f64(t_max)), _F(stepdist), _I(n), _p(out))
    return out


def infer_ray_start_dir(rays_o, rays_d, t_min):
    n = rays_o.shape[0]
    start, dirs = torch.empty(n, 3, dtype=_D), torch.empty(n, 3, dtype=_D)
    lib().orc64_infer_ray_start_dir(_p(_f64(rays_o)), _p(_f64(rays_d)), _p(_f64(t_min)), _I(n), _p(start), _p(dirs))
    return [start, dirs]


def sample_pts_on_rays(rays_o, rays_d, xyz_min, xyz_max, near, far, stepdist):
    ro, rd, lo, hi = _f64(rays_o), _f64(rays_d), _f64(xyz_min), _f64(xyz_max)
    n = ro.shape[0]
    t_min, t_max = torch.empty(n, dtype=_D), torch.empty(n, dtype=_D)
    N_steps = torch.empty(n, dtype=torch.int64)
    total = lib().orc64_sample_pts_count(_p(ro), _p(rd), _p(lo), _p(hi), _F(near), _F(far), _F(stepdist), _I(n),
                                         _p(t_min), _p(t_max), _p(N_steps))
    pts = torch.empty(total, 3, dtype=_D)
    mask = torch.empty(total, dtype=torch.uint8)
    ray_id = torch.empty(total, dtype=torch.int64)
    step_id = torch.empty(total, dtype=torch.int64)
    lib().orc64_sample_pts_fill(_p(ro), _p(rd), _p(lo), _p(hi), _p(t_min), _p(N_steps), _F(stepdist), _I(n),
                                _p(pts), _p(mask), _p(ray_id), _p(step_id))
    return [pts, mask.bool(), ray_id, step_id, N_steps, t_min, t_max]


def sample_ndc_pts_on_rays(rays_o, rays_d, xyz_min, xyz_max, N_samples):
    ro, rd = _f64(rays_o), _f64(rays_d)
    n = ro.shape[0]
    pts = torch.empty(n, N_samples, 3, dtype=_D)
    mask = torch.empty(n, N_samples, dtype=torch.uint8)
    lib().orc64_sample_ndc_pts_on_rays(_p(ro), _p(rd), _p(_f64(xyz_min)), _p(_f64(xyz_max)), _I(N_samples), _I(n),
                                       _p(pts), _p(mask))
    return [pts, mask.bool()]


def maskcache_lookup(world, xyz, xyz2ijk_scale, xyz2ijk_shift):
    w = world.to(torch.uint8).contiguous()
    x = _f64(xyz)
    out = torch.empty(x.shape[0], dtype=torch.uint8)
    lib().orc64_maskcache_lookup(_p(w), _p(x), _p(_f64(xyz2ijk_scale)), _p(_f64(xyz2ijk_shift)), _I(w.shape[0]),
                                 _I(w.shape[1]), _I(w.shape[2]), _L(x.shape[0]), _p(out))
    return out.bool()


def raw2alpha(density, shift, interval):
    d = _f64(density)
    e, a = torch.empty_like(d), torch.empty_like(d)
    lib().orc64_raw2alpha(_p(d), _F(shift), _F(interval), _L(d.numel()), _p(e), _p(a))
    return [e, a]


def raw2alpha_backward(exp, grad_back, interval):
    e, g = _f64(exp), _f64(grad_back)
    out = torch.empty_like(e)
    lib().orc64_raw2alpha_backward(_p(e), _p(g), _F(interval), _L(e.numel()), _p(out))
    return out


def alpha2weight(alpha, ray_id, n_rays):
    a = _f64(alpha)
    rid = ray_id.to(torch.int64).contiguous()
    n = a.numel()
    w, T = torch.empty(n, dtype=_D), torch.empty(n, dtype=_D)
    last = torch.empty(n_rays, dtype=_D)
    i_s = torch.empty(n_rays, dtype=torch.int64)
    i_e = torch.empty(n_rays, dtype=torch.int64)
    lib().orc64_alpha2weight(_p(a), _p(rid), _I(n_rays), _L(n), _p(w), _p(T), _p(last), _p(i_s), _p(i_e))
    return [w, T, last, i_s, i_e]


def alpha2weight_backward(alpha, weight, T, alphainv_last, i_start, i_end, n_rays, grad_weights, grad_last):
    a = _f64(alpha)
    out = torch.empty_like(a)
    lib().orc64_alpha2weight_backward(_p(a), _p(_f64(weight)), _p(_f64(T)), _p(_f64(alphainv_last)),
                                      _p(i_start.contiguous()), _p(i_end.contiguous()), _I(n_rays), _L(a.numel()),
                                      _p(_f64(grad_weights)), _p(_f64(grad_last)), _p(out))
    return out


def total_variation_add_grad(param, grad, wx, wy, wz, dense_mode):
    assert param.dtype == _D and grad.dtype == _D and grad.is_contiguous()
    lib().orc64_total_variation_add_grad(_p(_f64(param)), _p(grad), _F(wx), _F(wy), _F(wz), _I(1 if dense_mode else 0),
                                         _L(param.numel()), _L(param.shape[2]), _L(param.shape[3]), _L(param.shape[4]))


def _adam(mode, param, grad, exp_avg, exp_avg_sq, perlr, step, beta1, beta2, lr, eps):
    for t in (param, grad, exp_avg, exp_avg_sq):
        assert t.dtype == _D and t.is_contiguous()
    lib().orc64_adam_upd(_p(param), _p(grad), _p(exp_avg), _p(exp_avg_sq),
                         _p(_f64(perlr)) if perlr is not None else _P(0), _L(param.numel()), _I(step), _F(beta1),
                         _F(beta2), _F(lr), _F(eps), _I(mode))


def adam_upd(param, grad, exp_avg, exp_avg_sq, step, beta1, beta2, lr, eps):
    _adam(0, param, grad, exp_avg, exp_avg_sq, None, step, beta1, beta2, lr, eps)


def masked_adam_upd(param, grad, exp_avg, exp_avg_sq, step, beta1, beta2, lr, eps):
    _adam(1, param, grad, exp_avg, exp_avg_sq, None, step, beta1, beta2, lr, eps)


def adam_upd_with_perlr(param, grad, exp_avg, exp_avg_sq, perlr, step, beta1, beta2, lr, eps):
    _adam(2, param, grad, exp_avg, exp_avg_sq, perlr, step, beta1, beta2, lr, eps)
