"""Shared driver for the float64 parity tests: runs every op of the reference's three extension modules on one
set of inputs through a given implementation (the CUDA product, the reference's own double kernels in oracle/_ref,
or the CPU oracle oracle/oracle_f64.py) and returns the outputs under the key names of
tests/golden/ref_gpu_ops_f64.npz (oracle/make_golden_gpu.py ... f64)."""
import numpy as np
import torch

from tests.util import make_rays, sorted_ray_ids

INPUT_KEYS = ("rays_o", "rays_d", "xyz_min", "xyz_max", "near", "far", "stepdist", "world", "scale", "shift",
              "ndc_o", "ndc_d", "ndc_n", "density", "shift_a", "interval", "grad_back", "a2w_alpha", "a2w_ray_id",
              "a2w_n_rays", "a2w_gw", "a2w_gl", "tv_param", "tv_grad_in", "tv_wy", "tv_wz", "adam_p", "adam_m",
              "adam_v", "adam_g", "adam_perlr")
# outputs that are integers / booleans (bit-exact class) -- everything else is float64
EXACT_KEYS = ("mask_outbbox", "ray_id", "step_id", "N_steps", "maskcache", "ndc_mask", "i_start", "i_end")


def ulp_diff64(a, b):
    """Distance in float64 units-in-the-last-place (same shape)."""
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64)).view(np.int64)
    b = np.ascontiguousarray(np.asarray(b, dtype=np.float64)).view(np.int64)
    a = np.where(a < 0, np.int64(-2 ** 63) - a, a)   # monotone map of the bit patterns onto the integers
    b = np.where(b < 0, np.int64(-2 ** 63) - b, b)
    # int64 subtraction (wraps only for operands of opposite sign and huge magnitude: reported as a huge distance)
    with np.errstate(over="ignore"):
        d = np.abs(a - b)
    return np.where(d < 0, np.iinfo(np.int64).max, d).astype(np.float64)


def make_inputs(seed, n_rays=512, n_pts=60000, a2w_rays=700, n_adam=100003, grid=(3, 21, 18, 15)):
    """Seeded float64 inputs that are NOT float32-representable (so that the float temporaries of the reference's
    double instantiation matter), incl. rays that miss the box and exact-zero direction components."""
    g = torch.Generator().manual_seed(seed)
    w = lambda t: t.double() * (1.0 + 2.0 ** -30) + 0.0
    lo, hi = torch.tensor([-1.0, -0.9, -0.8]), torch.tensor([1.0, 0.9, 0.8])
    ro, rd, _, _ = make_rays(n_rays, seed + 1)
    d = {"rays_o": w(ro), "rays_d": w(rd), "xyz_min": w(lo), "xyz_max": w(hi), "near": 0.2, "far": 6.0,
         "stepdist": 0.0213}
    d["world"] = torch.rand(23, 19, 17, generator=g) > 0.4
    shape = torch.tensor([23.0, 19.0, 17.0]).double()
    d["scale"] = (shape - 1) / (d["xyz_max"] - d["xyz_min"])
    d["shift"] = -d["xyz_min"] * d["scale"]
    d["ndc_o"] = w(torch.cat([(torch.rand(64, 2, generator=g) - 0.5) * 2.4, -torch.ones(64, 1)], -1)).contiguous()
    d["ndc_d"] = w(torch.cat([(torch.rand(64, 2, generator=g) - 0.5) * 0.9, 2 * torch.ones(64, 1)], -1)).contiguous()
    d["ndc_n"] = 65
    d["density"] = w(torch.cat([torch.randn(n_pts, generator=g) * 4, torch.tensor([100.0, -100.0, 0.0, 88.0, 30.0])]))
    d["shift_a"], d["interval"] = -4.595, 0.5
    d["grad_back"] = w(torch.randn(d["density"].shape, generator=g))
    d["a2w_n_rays"] = a2w_rays
    d["a2w_ray_id"] = sorted_ray_ids(a2w_rays, n_pts, seed + 2)
    d["a2w_alpha"] = w(torch.rand(n_pts, generator=g) ** 3 * 0.6)
    d["a2w_gw"] = w(torch.randn(n_pts, generator=g))
    d["a2w_gl"] = w(torch.randn(a2w_rays, generator=g))
    d["tv_param"] = w(torch.randn(1, *grid, generator=g) * 1.5)
    gi = torch.randn(1, *grid, generator=g)
    gi[torch.rand(gi.shape, generator=g) < 0.5] = 0
    d["tv_grad_in"] = w(gi)
    d["tv_wy"], d["tv_wz"] = 0.7, 1.3
    d["adam_p"] = w(torch.randn(n_adam, generator=g))
    d["adam_m"] = w(torch.randn(n_adam, generator=g) * 0.01)
    d["adam_v"] = w(torch.rand(n_adam, generator=g) * 1e-3)
    gr = torch.randn(n_adam, generator=g)
    gr[torch.rand(n_adam, generator=g) < 0.4] = 0
    d["adam_g"] = w(gr)
    d["adam_perlr"] = w(torch.rand(n_adam, generator=g))
    return d


def inputs_from_golden(g):
    d = {}
    for k in INPUT_KEYS:
        v = g[k]
        d[k] = torch.from_numpy(np.array(v)) if np.ndim(v) else v.item()
    return d


def run_suite(ru, tv, ad, inp, dev):
    """All fourteen ops; `ru` / `tv` / `ad` are module-likes with the reference's pybind names."""
    T = lambda k: inp[k].to(dev).contiguous()
    c = lambda t: t.detach().cpu().numpy()
    out = {}
    ro, rd, lo, hi = T("rays_o"), T("rays_d"), T("xyz_min"), T("xyz_max")
    near, far, stepdist = float(inp["near"]), float(inp["far"]), float(inp["stepdist"])
    pts, mask, ray_id, step_id, N_steps, t_min, t_max = ru.sample_pts_on_rays(ro, rd, lo, hi, near, far, stepdist)
    out.update(rays_pts=c(pts), mask_outbbox=c(mask), ray_id=c(ray_id), step_id=c(step_id), N_steps=c(N_steps),
               t_min=c(t_min), t_max=c(t_max))
    tm2, tx2 = ru.infer_t_minmax(ro, rd, lo, hi, near, far)
    assert np.array_equal(c(tm2), out["t_min"]) and np.array_equal(c(tx2), out["t_max"])
    assert np.array_equal(c(ru.infer_n_samples(t_min, t_max, stepdist)), out["N_steps"])
    start, dirs = ru.infer_ray_start_dir(ro, rd, t_min)
    out.update(rays_start=c(start), rays_dir=c(dirs))
    out["maskcache"] = c(ru.maskcache_lookup(T("world"), pts, T("scale"), T("shift")))
    ndc_pts, ndc_mask = ru.sample_ndc_pts_on_rays(T("ndc_o"), T("ndc_d"), lo, hi, int(inp["ndc_n"]))
    out.update(ndc_pts=c(ndc_pts), ndc_mask=c(ndc_mask))

    shift_a, interval = float(inp["shift_a"]), float(inp["interval"])
    exp_d, alpha = ru.raw2alpha(T("density"), shift_a, interval)
    out.update(exp_d=c(exp_d), alpha=c(alpha))
    # the backward is fed the GIVEN exp_d when the inputs carry one (so that libm differences of the forward do not
    # leak into the check of the backward), else its own
    e_in = inp["exp_d_in"].to(dev) if "exp_d_in" in inp else exp_d
    out["raw2alpha_grad"] = c(ru.raw2alpha_backward(e_in, T("grad_back"), interval))

    n_rays = int(inp["a2w_n_rays"])
    a, rid = T("a2w_alpha"), T("a2w_ray_id")
    w, Tt, last, i_s, i_e = ru.alpha2weight(a, rid, n_rays)
    out.update(weight=c(w), T=c(Tt), alphainv_last=c(last), i_start=c(i_s), i_end=c(i_e))
    out["a2w_grad"] = c(ru.alpha2weight_backward(a, w, Tt, last, i_s, i_e, n_rays, T("a2w_gw"), T("a2w_gl")))

    p = T("tv_param")
    for dense in (0, 1):
        gcopy = T("tv_grad_in").clone()
        tv.total_variation_add_grad(p, gcopy, 0.3, float(inp["tv_wy"]), float(inp["tv_wz"]), bool(dense))
        out["tv_out_dense%d" % dense] = c(gcopy)

    for name in ("adam_upd", "masked_adam_upd", "adam_upd_with_perlr"):
        pp, m, v = T("adam_p").clone(), T("adam_m").clone(), T("adam_v").clone()
        for step in (1, 2, 3):
            args = (pp, T("adam_g"), m, v) + ((T("adam_perlr"),) if name == "adam_upd_with_perlr" else ())
            getattr(ad, name)(*args, step, 0.9, 0.99, 0.1, 1e-8)
        out["adam_out_p_" + name], out["adam_out_m_" + name], out["adam_out_v_" + name] = c(pp), c(m), c(v)
    return out


def compare(got, want, ulps=0, libm_ulps=None, label=""):
    """Integer / boolean outputs bit-exact; float64 outputs within `ulps` float64 units in the last place, the three
    outputs that go through exp() / pow() within `libm_ulps` (CPU libm vs CUDA libdevice).  Returns {key: max ulp
    distance observed}."""
    seen = {}
    libm_ulps = ulps if libm_ulps is None else libm_ulps
    libm_keys = ("exp_d", "alpha", "raw2alpha_grad")
    for k, b in want.items():
        if k in INPUT_KEYS or k not in got:
            continue
        a = got[k]
        assert a.shape == b.shape and a.dtype == b.dtype, (label, k, a.shape, b.shape, a.dtype, b.dtype)
        if k in EXACT_KEYS:
            assert np.array_equal(a, b), (label, k)
            continue
        fin = np.isfinite(b)
        assert np.array_equal(np.isfinite(a), fin) and np.array_equal(a[~fin], b[~fin]), (label, k, "non-finite")
        if k == "alpha" and libm_ulps > ulps:
            # 1 - pow(..): an ulp of pow (~1) is many ulps of a small alpha; state it as an absolute bound
            err = np.abs(a[fin] - b[fin]).max()
            assert err <= libm_ulps * 2.0 ** -52, (label, k, err)
            seen[k] = float(err / 2.0 ** -52)
            continue
        d = ulp_diff64(a[fin], b[fin]).max() if fin.any() else 0.0
        seen[k] = float(d)
        assert d <= (libm_ulps if k in libm_keys else ulps), (label, k, d)
    return seen
