"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of oracle/libdvgo_oracle.so (the plain-C CPU
restatement of the reference kernels, oracle/dvgo_oracle.c).

Exposes three namespaces with the reference's pybind surface (lib/cuda/render_utils.cpp:144-155,
total_variation.cpp:22-24, adam_upd.cpp:79-86) operating on CPU torch tensors, so that
  * tests can compare the CUDA product against it op by op, and
  * the reference's own Python (lib/dvgo.py ...) can be imported in the GPU-less build container
    with `directvoxgo_b200.dropin.install(oracle.as_modules())` to generate golden fixtures.
Nothing in the product package imports this module.
"""
import ctypes
import os
import subprocess
import types

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libdvgo_oracle.so")


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("dvgo_oracle.c", "dvgo_oracle_f64.c")]
    if force or not os.path.exists(_LIB_PATH) or max(map(os.path.getmtime, srcs)) > os.path.getmtime(_LIB_PATH):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-ffp-contract=off",
                               "-fno-fast-math", "-o", _LIB_PATH] + srcs + ["-lm"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.orc_sample_pts_count.restype = ctypes.c_int64
        _lib.orc64_sample_pts_count.restype = ctypes.c_int64
        _lib.orc_adam_step_size.restype = ctypes.c_float
    return _lib


_F, _I, _L, _P = ctypes.c_float, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p


def _p(t):
    assert t.device.type == "cpu" and t.is_contiguous(), "oracle works on contiguous CPU tensors"
    return _P(t.data_ptr())


def _f32(t):
    return t.detach().to(torch.float32).contiguous()


# ---------------------------------------------------------------- render_utils_cuda surface
def infer_t_minmax(rays_o, rays_d, xyz_min, xyz_max, near, far):
    n = rays_o.shape[0]
    t_min, t_max = torch.empty(n), torch.empty(n)
    lib().orc_infer_t_minmax(_p(_f32(rays_o)), _p(_f32(rays_d)), _p(_f32(xyz_min)), _p(_f32(xyz_max)),
                             _F(near), _F(far), _I(n), _p(t_min), _p(t_max))
    return [t_min, t_max]


def infer_n_samples(t_min, t_max, stepdist):
    n = t_min.shape[0]
    out = torch.empty(n, dtype=torch.int64)
    lib().orc_infer_n_samples(_p(_f32(t_min)), _p(_f32(t_max)), _F(stepdist), _I(n), _p(out))
    return out


def infer_ray_start_dir(rays_o, rays_d, t_min):
    n = rays_o.shape[0]
    start, dirs = torch.empty(n, 3), torch.empty(n, 3)
    lib().orc_infer_ray_start_dir(_p(_f32(rays_o)), _p(_f32(rays_d)), _p(_f32(t_min)), _I(n),
                                  _p(start), _p(dirs))
    return [start, dirs]


def sample_pts_on_rays(rays_o, rays_d, xyz_min, xyz_max, near, far, stepdist):
    ro, rd, lo, hi = _f32(rays_o), _f32(rays_d), _f32(xyz_min), _f32(xyz_max)
    n = ro.shape[0]
    t_min, t_max = torch.empty(n), torch.empty(n)
    N_steps = torch.empty(n, dtype=torch.int64)
    total = lib().orc_sample_pts_count(_p(ro), _p(rd), _p(lo), _p(hi), _F(near), _F(far), _F(stepdist),
                                       _I(n), _p(t_min), _p(t_max), _p(N_steps))
    pts = torch.empty(total, 3)
    mask = torch.empty(total, dtype=torch.uint8)
    ray_id = torch.empty(total, dtype=torch.int64)
    step_id = torch.empty(total, dtype=torch.int64)
    lib().orc_sample_pts_fill(_p(ro), _p(rd), _p(lo), _p(hi), _p(t_min), _p(N_steps), _F(stepdist), _I(n),
                              _p(pts), _p(mask), _p(ray_id), _p(step_id))
    return [pts, mask.bool(), ray_id, step_id, N_steps, t_min, t_max]


def sample_ndc_pts_on_rays(rays_o, rays_d, xyz_min, xyz_max, N_samples):
    ro, rd = _f32(rays_o), _f32(rays_d)
    n = ro.shape[0]
    pts = torch.empty(n, N_samples, 3)
    mask = torch.empty(n, N_samples, dtype=torch.uint8)
    lib().orc_sample_ndc_pts_on_rays(_p(ro), _p(rd), _p(_f32(xyz_min)), _p(_f32(xyz_max)),
                                     _I(N_samples), _I(n), _p(pts), _p(mask))
    return [pts, mask.bool()]


def maskcache_lookup(world, xyz, xyz2ijk_scale, xyz2ijk_shift):
    w = world.to(torch.uint8).contiguous()
    x = _f32(xyz)
    out = torch.empty(x.shape[0], dtype=torch.uint8)
    lib().orc_maskcache_lookup(_p(w), _p(x), _p(_f32(xyz2ijk_scale)), _p(_f32(xyz2ijk_shift)),
                               _I(w.shape[0]), _I(w.shape[1]), _I(w.shape[2]), _L(x.shape[0]), _p(out))
    return out.bool()


def raw2alpha(density, shift, interval):
    d = _f32(density)
    e, a = torch.empty_like(d), torch.empty_like(d)
    lib().orc_raw2alpha(_p(d), _F(shift), _F(interval), _L(d.numel()), _p(e), _p(a))
    return [e, a]


def raw2alpha_backward(exp, grad_back, interval):
    e, g = _f32(exp), _f32(grad_back)
    out = torch.empty_like(e)
    lib().orc_raw2alpha_backward(_p(e), _p(g), _F(interval), _L(e.numel()), _p(out))
    return out


def alpha2weight(alpha, ray_id, n_rays):
    a = _f32(alpha)
    rid = ray_id.to(torch.int64).contiguous()
    n = a.numel()
    w, T = torch.empty(n), torch.empty(n)
    last = torch.empty(n_rays)
    i_s = torch.empty(n_rays, dtype=torch.int64)
    i_e = torch.empty(n_rays, dtype=torch.int64)
    lib().orc_alpha2weight(_p(a), _p(rid), _I(n_rays), _L(n), _p(w), _p(T), _p(last), _p(i_s), _p(i_e))
    return [w, T, last, i_s, i_e]


def alpha2weight_backward(alpha, weight, T, alphainv_last, i_start, i_end, n_rays, grad_weights, grad_last):
    a = _f32(alpha)
    out = torch.empty_like(a)
    lib().orc_alpha2weight_backward(_p(a), _p(_f32(weight)), _p(_f32(T)), _p(_f32(alphainv_last)),
                                    _p(i_start.contiguous()), _p(i_end.contiguous()), _I(n_rays),
                                    _L(a.numel()), _p(_f32(grad_weights)), _p(_f32(grad_last)), _p(out))
    return out


# ---------------------------------------------------------------- total_variation_cuda / adam_upd_cuda
def total_variation_add_grad(param, grad, wx, wy, wz, dense_mode):
    assert param.dtype == torch.float32 and grad.dtype == torch.float32
    lib().orc_total_variation_add_grad(_p(param.detach()), _p(grad), _F(wx), _F(wy), _F(wz),
                                       _I(1 if dense_mode else 0), _L(param.numel()),
                                       _L(param.shape[2]), _L(param.shape[3]), _L(param.shape[4]))


def _adam(mode, param, grad, exp_avg, exp_avg_sq, perlr, step, beta1, beta2, lr, eps):
    lib().orc_adam_upd(_p(param.detach()), _p(grad), _p(exp_avg), _p(exp_avg_sq),
                       _p(perlr) if perlr is not None else _P(0), _L(param.numel()), _I(step),
                       _F(beta1), _F(beta2), _F(lr), _F(eps), _I(mode))


def adam_upd(param, grad, exp_avg, exp_avg_sq, step, beta1, beta2, lr, eps):
    _adam(0, param, grad, exp_avg, exp_avg_sq, None, step, beta1, beta2, lr, eps)


def masked_adam_upd(param, grad, exp_avg, exp_avg_sq, step, beta1, beta2, lr, eps):
    _adam(1, param, grad, exp_avg, exp_avg_sq, None, step, beta1, beta2, lr, eps)


def adam_upd_with_perlr(param, grad, exp_avg, exp_avg_sq, perlr, step, beta1, beta2, lr, eps):
    _adam(2, param, grad, exp_avg, exp_avg_sq, perlr.contiguous(), step, beta1, beta2, lr, eps)


# ---------------------------------------------------------------- ATen / torch_scatter restatements
def grid_sample_3d(grid, xyz, xyz_min, xyz_max):
    """grid [1,C,X,Y,Z], xyz [P,3] -> [P,C] (restates F.grid_sample as used at lib/dvgo.py:312-328)."""
    g, x = _f32(grid), _f32(xyz)
    C, X, Y, Z = g.shape[1:]
    out = torch.empty(x.shape[0], C)
    lib().orc_grid_sample_3d(_p(g), _I(C), _I(X), _I(Y), _I(Z), _p(x), _p(_f32(xyz_min)), _p(_f32(xyz_max)),
                             _L(x.shape[0]), _p(out))
    return out


def grid_sample_3d_backward(grad_out, xyz, xyz_min, xyz_max, grad_grid):
    go, x = _f32(grad_out), _f32(xyz)
    C, X, Y, Z = grad_grid.shape[1:]
    lib().orc_grid_sample_3d_backward(_p(go), _I(C), _I(X), _I(Y), _I(Z), _p(x), _p(_f32(xyz_min)),
                                      _p(_f32(xyz_max)), _L(x.shape[0]), _p(grad_grid))


def grid_sample_2d(plane, xyz, xyz_min, xyz_max, axis_w, axis_h):
    """plane [1,C,H,W], xyz [P,3] -> [P,C] (F.grid_sample 2-D as used at lib/tri_dvgo.py:456-464)."""
    g, x = _f32(plane), _f32(xyz)
    C, H, W = g.shape[1:]
    out = torch.empty(x.shape[0], C)
    lib().orc_grid_sample_2d(_p(g), _I(C), _I(H), _I(W), _p(x), _p(_f32(xyz_min)), _p(_f32(xyz_max)), _I(axis_w),
                             _I(axis_h), _L(x.shape[0]), _p(out))
    return out


def grid_sample_2d_backward(grad_out, xyz, xyz_min, xyz_max, axis_w, axis_h, grad_plane):
    go, x = _f32(grad_out), _f32(xyz)
    C, H, W = grad_plane.shape[1:]
    lib().orc_grid_sample_2d_backward(_p(go), _I(C), _I(H), _I(W), _p(x), _p(_f32(xyz_min)), _p(_f32(xyz_max)),
                                      _I(axis_w), _I(axis_h), _L(x.shape[0]), _p(grad_plane))


def segment_coo_sum(src, index, out):
    s = _f32(src)
    P = index.shape[0]
    D = s.numel() // P if P else 1
    lib().orc_segment_coo_sum(_p(s), _p(index.contiguous()), _L(P), _I(D), _p(out))
    return out


def segment_coo(src, index, out=None, dim_size=None, reduce="sum"):
    """Differentiable torch_scatter.segment_coo stand-in for the CPU run of the reference's Python
    (== out.index_add_(0, index, src))."""
    assert reduce in ("sum", "add")
    if out is None:
        out = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype)
    return out.index_add(0, index, src)


def scatter_add(src, index, dim=0, out=None, dim_size=None):
    if out is None:
        shape = list(src.shape); shape[dim] = dim_size
        out = torch.zeros(shape, dtype=src.dtype)
    return out.index_add(dim, index, src)


def as_modules():
    """{name: module-like} for directvoxgo_b200.dropin.install(): the reference's three extension
    names + torch_scatter, all served by the CPU oracle."""
    ru = types.SimpleNamespace(
        infer_t_minmax=infer_t_minmax, infer_n_samples=infer_n_samples,
        infer_ray_start_dir=infer_ray_start_dir, sample_pts_on_rays=sample_pts_on_rays,
        sample_ndc_pts_on_rays=sample_ndc_pts_on_rays, maskcache_lookup=maskcache_lookup,
        raw2alpha=raw2alpha, raw2alpha_backward=raw2alpha_backward, alpha2weight=alpha2weight,
        alpha2weight_backward=alpha2weight_backward)
    tv = types.SimpleNamespace(total_variation_add_grad=total_variation_add_grad)
    ad = types.SimpleNamespace(adam_upd=adam_upd, masked_adam_upd=masked_adam_upd,
                               adam_upd_with_perlr=adam_upd_with_perlr)
    ts = types.ModuleType("torch_scatter")
    ts.segment_coo, ts.scatter_add = segment_coo, scatter_add
    return {"render_utils_cuda": ru, "total_variation_cuda": tv, "adam_upd_cuda": ad, "torch_scatter": ts}
