"""GPU parity tests, op by op: the CUDA product (through the C ABI, via the thin torch binding and
once via raw ctypes) against the CPU oracle on identical seeded inputs.

Parity classes:  bit-exact -- N_steps, ray_id, step_id, mask_outbbox, maskcache,
i_start, and (because the expression trees are reproduced with explicit fma/div intrinsics) t_min,
t_max, rays_pts; fp32 tolerance stated per assert for everything else."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import oracle as orc
from tests.util import make_rays, rel_to_max, sorted_ray_ids, to_np, ulp_diff

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def pkg():
    import directvoxgo_b200 as p
    return p


def _box():
    return torch.tensor([-1.0, -0.9, -0.8]), torch.tensor([1.0, 0.9, 0.8])


@pytest.mark.parametrize("n_rays,seed,stepdist", [(1, 0, 0.05), (37, 1, 0.031), (1000, 2, 0.0123), (5000, 3, 0.05)])
def test_sample_pts_on_rays_bit_exact(pkg, n_rays, seed, stepdist):
    lo, hi = _box()
    ro, rd, _, _ = make_rays(n_rays, seed, miss=min(4, n_rays - 1))
    near, far = 0.2, 6.0
    ref = orc.sample_pts_on_rays(ro, rd, lo, hi, near, far, stepdist)
    got = pkg.render_utils_cuda.sample_pts_on_rays(ro.to(DEV), rd.to(DEV), lo.to(DEV), hi.to(DEV), near, far, stepdist)
    names = ["rays_pts", "mask_outbbox", "ray_id", "step_id", "N_steps", "t_min", "t_max"]
    for n, a, b in zip(names, got, ref):
        assert a.shape == b.shape and a.dtype == b.dtype, n
        assert np.array_equal(to_np(a), to_np(b)), n  # bit-exact incl. the float outputs
    # the sub-ops exported on their own
    tmin, tmax = pkg.render_utils_cuda.infer_t_minmax(ro.to(DEV), rd.to(DEV), lo.to(DEV), hi.to(DEV), near, far)
    assert np.array_equal(to_np(tmin), to_np(ref[5])) and np.array_equal(to_np(tmax), to_np(ref[6]))
    ns = pkg.render_utils_cuda.infer_n_samples(tmin, tmax, stepdist)
    assert ns.dtype == torch.int64 and np.array_equal(to_np(ns), to_np(ref[4]))
    st, dr = pkg.render_utils_cuda.infer_ray_start_dir(ro.to(DEV), rd.to(DEV), tmin)
    st_ref, dr_ref = orc.infer_ray_start_dir(ro, rd, ref[5])
    assert np.array_equal(to_np(st), to_np(st_ref)) and np.array_equal(to_np(dr), to_np(dr_ref))


def test_sample_ndc_and_maskcache_bit_exact(pkg):
    lo, hi = _box()
    g = torch.Generator().manual_seed(5)
    o = torch.cat([(torch.rand(300, 2, generator=g) - 0.5) * 2.4, -torch.ones(300, 1)], -1).contiguous()
    d = torch.cat([(torch.rand(300, 2, generator=g) - 0.5) * 0.9, 2 * torch.ones(300, 1)], -1).contiguous()
    pts_ref, m_ref = orc.sample_ndc_pts_on_rays(o, d, lo, hi, 65)
    pts, m = pkg.render_utils_cuda.sample_ndc_pts_on_rays(o.to(DEV), d.to(DEV), lo.to(DEV), hi.to(DEV), 65)
    assert pts.shape == (300, 65, 3) and m.shape == (300, 65) and m.dtype == torch.bool
    assert np.array_equal(to_np(pts), to_np(pts_ref)) and np.array_equal(to_np(m), to_np(m_ref))
    world = torch.rand(31, 17, 23, generator=g) > 0.5
    scale = (torch.tensor([31.0, 17.0, 23.0]) - 1) / (hi - lo)
    shift = -lo * scale
    xyz = pts_ref.reshape(-1, 3)
    xyz = torch.cat([xyz, xyz * 1.7, torch.tensor([[lo[0], lo[1], lo[2]], [hi[0], hi[1], hi[2]]])]).contiguous()
    occ_ref = orc.maskcache_lookup(world, xyz, scale, shift)
    occ = pkg.render_utils_cuda.maskcache_lookup(world.to(DEV), xyz.to(DEV), scale.to(DEV), shift.to(DEV))
    assert occ.dtype == torch.bool and np.array_equal(to_np(occ), to_np(occ_ref))
    # empty input (rays that miss everything) -- reference short-circuit render_utils_kernel.cu:333-335
    e = pkg.render_utils_cuda.maskcache_lookup(world.to(DEV), torch.zeros(0, 3, device=DEV), scale.to(DEV), shift.to(DEV))
    assert e.shape == (0,)


def test_raw2alpha_and_backward(pkg):
    g = torch.Generator().manual_seed(9)
    d = torch.cat([torch.randn(100000, generator=g) * 4, torch.tensor([100.0, -100.0, 0.0, 88.0, 30.0])])
    for shift, interval in [(-4.595, 0.5), (0.0, 1.0), (-13.8, 0.25)]:
        e_ref, a_ref = orc.raw2alpha(d, shift, interval)
        e, a = pkg.render_utils_cuda.raw2alpha(d.to(DEV), shift, interval)
        assert ulp_diff(to_np(e), to_np(e_ref)).max() <= 2          # CUDA expf vs glibc expf
        np.testing.assert_allclose(to_np(a), to_np(a_ref), rtol=0, atol=3e-7)  # 1 - powf(): abs tol
        gb = torch.randn(d.shape, generator=g)
        gr_ref = orc.raw2alpha_backward(e_ref, gb, interval)
        gr = pkg.render_utils_cuda.raw2alpha_backward(e_ref.to(DEV), gb.to(DEV), interval)
        np.testing.assert_allclose(to_np(gr), to_np(gr_ref), rtol=3e-6, atol=1e-30)
    z = pkg.render_utils_cuda.raw2alpha(torch.zeros(0, device=DEV), 0.0, 1.0)
    assert z[0].shape == (0,) and z[1].shape == (0,)


@pytest.mark.parametrize("n_rays,n_pts,seed,amax", [(1, 1, 0, 0.5), (7, 100, 1, 0.9), (300, 40000, 2, 0.3),
                                                    (2000, 500000, 3, 0.05), (64, 30000, 4, 0.999)])
def test_alpha2weight_and_backward(pkg, n_rays, n_pts, seed, amax):
    rid = sorted_ray_ids(n_rays, n_pts, seed)
    g = torch.Generator().manual_seed(seed + 100)
    alpha = torch.rand(n_pts, generator=g) ** 2 * amax
    w_r, T_r, last_r, is_r, ie_r = orc.alpha2weight(alpha, rid, n_rays)
    w, T, last, i_s, i_e = pkg.render_utils_cuda.alpha2weight(alpha.to(DEV), rid.to(DEV), n_rays)
    # the warp replays the reference's per-sample float recurrence: every output is bit-exact, the stop index included
    for name, x, y in (("i_start", i_s, is_r), ("i_end", i_e, ie_r), ("T", T, T_r), ("weights", w, w_r),
                       ("alphainv_last", last, last_r)):
        assert np.array_equal(to_np(x), to_np(y)), name
    gw = torch.randn(n_pts, generator=g)
    gl = torch.randn(n_rays, generator=g)
    g_ref = orc.alpha2weight_backward(alpha, w_r, T_r, last_r, is_r, ie_r, n_rays, gw, gl)
    g_got = pkg.render_utils_cuda.alpha2weight_backward(alpha.to(DEV), w_r.to(DEV), T_r.to(DEV), last_r.to(DEV),
                                                       is_r.to(DEV), ie_r.to(DEV), n_rays, gw.to(DEV), gl.to(DEV))
    assert np.array_equal(to_np(g_got), to_np(g_ref))


def test_alpha2weight_empty(pkg):
    out = pkg.render_utils_cuda.alpha2weight(torch.zeros(0, device=DEV), torch.zeros(0, dtype=torch.int64, device=DEV), 5)
    w, T, last, i_s, i_e = out
    assert w.numel() == 0 and torch.all(last == 1) and torch.all(i_s == 0) and torch.all(i_e == 0)


@pytest.mark.parametrize("C,shape,n", [(1, (9, 8, 7), 5000), (3, (16, 5, 11), 20000), (12, (20, 21, 19), 30000)])
def test_grid_sample_trilinear_fwd_bwd(pkg, C, shape, n):
    g = torch.Generator().manual_seed(C)
    grid = torch.randn(1, C, *shape, generator=g)
    lo, hi = torch.tensor([-1.0, -2.0, 0.5]), torch.tensor([1.5, 1.0, 2.5])
    xyz = (lo + (hi - lo) * (torch.rand(n, 3, generator=g) * 1.2 - 0.1)).contiguous()
    xyz[:8] = torch.stack([lo, hi, lo, hi, (lo + hi) / 2, lo, hi, lo])  # exact corners / faces
    ref = orc.grid_sample_3d(grid, xyz, lo, hi)
    got = pkg.ext.grid_sample_3d(grid.to(DEV), xyz.to(DEV), lo.to(DEV), hi.to(DEV))
    # identical expression tree (explicit fma) -> expect <= 1 ulp; state 1e-6 abs / 1e-5 rel
    np.testing.assert_allclose(to_np(got), to_np(ref), rtol=1e-5, atol=1e-6)
    go = torch.randn(n, C, generator=g)
    gg_ref = torch.zeros_like(grid)
    orc.grid_sample_3d_backward(go, xyz, lo, hi, gg_ref)
    gg = torch.zeros_like(grid, device=DEV)
    pkg.ext.grid_sample_3d_backward(go.to(DEV), xyz.to(DEV), lo.to(DEV), hi.to(DEV), gg)
    assert rel_to_max(gg, gg_ref) < 1e-5   # fp32 atomics, unordered
    # autograd wrapper incl. the C==1 squeeze of lib/dvgo.py:325-326
    from directvoxgo_b200.ops import grid_sample_trilinear
    gr = grid.to(DEV).requires_grad_()
    out = grid_sample_trilinear(gr, xyz.to(DEV), lo.to(DEV), hi.to(DEV))
    assert out.shape == ((n,) if C == 1 else (n, C))
    (out.reshape(n, C) * go.to(DEV)).sum().backward()
    assert rel_to_max(gr.grad, gg_ref) < 1e-5


def test_segment_coo(pkg):
    from directvoxgo_b200.ops import segment_coo
    rid = sorted_ray_ids(500, 100000, 8)
    g = torch.Generator().manual_seed(8)
    src = torch.randn(100000, 3, generator=g)
    ref = torch.zeros(500, 3).index_add(0, rid, src)
    s = src.to(DEV).requires_grad_()
    out = segment_coo(src=s, index=rid.to(DEV), out=torch.zeros(500, 3, device=DEV), reduce="sum")
    np.testing.assert_allclose(to_np(out), to_np(ref), rtol=1e-4, atol=1e-4)
    go = torch.randn(500, 3, generator=g)
    (out * go.to(DEV)).sum().backward()
    assert np.array_equal(to_np(s.grad), to_np(go[rid]))
    src1 = torch.randn(100000, generator=g)
    out1 = segment_coo(src=src1.to(DEV), index=rid.to(DEV), out=torch.zeros(500, device=DEV), reduce="sum")
    np.testing.assert_allclose(to_np(out1), to_np(torch.zeros(500).index_add(0, rid, src1)), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("shape", [(1, 1, 5, 6, 7), (1, 3, 9, 8, 7), (1, 12, 16, 15, 17)])
def test_total_variation(pkg, shape):
    g = torch.Generator().manual_seed(3)
    param = torch.randn(shape, generator=g) * 1.5
    grad = torch.randn(shape, generator=g)
    grad[torch.rand(shape, generator=g) < 0.5] = 0
    for dense in (False, True):
        ref = grad.clone()
        orc.total_variation_add_grad(param, ref, 0.3, 0.7, 1.3, dense)
        got = grad.clone().to(DEV)
        pkg.total_variation_cuda.total_variation_add_grad(param.to(DEV), got, 0.3, 0.7, 1.3, dense)
        assert ulp_diff(to_np(got), to_np(ref)).max() <= 1   # deterministic, same expression tree


@pytest.mark.parametrize("N", [1, 5, 4099, 1 << 20])
def test_adam_variants(pkg, N):
    g = torch.Generator().manual_seed(N)
    p0, m0, v0 = torch.randn(N, generator=g), torch.randn(N, generator=g) * 0.01, torch.rand(N, generator=g) * 1e-3
    gr = torch.randn(N, generator=g)
    gr[torch.rand(N, generator=g) < 0.4] = 0
    perlr = torch.rand(N, generator=g)
    for name in ("adam_upd", "masked_adam_upd", "adam_upd_with_perlr"):
        pr, mr, vr = p0.clone(), m0.clone(), v0.clone()
        pg, mg, vg = p0.clone().to(DEV), m0.clone().to(DEV), v0.clone().to(DEV)
        for step in (1, 2, 3):
            extra_r = (perlr,) if name.endswith("perlr") else ()
            extra_g = (perlr.to(DEV),) if name.endswith("perlr") else ()
            getattr(orc, name)(pr, gr, mr, vr, *extra_r, step, 0.9, 0.99, 0.1, 1e-8)
            getattr(pkg.adam_upd_cuda, name)(pg, gr.to(DEV), mg, vg, *extra_g, step, 0.9, 0.99, 0.1, 1e-8)
        assert ulp_diff(to_np(mg), to_np(mr)).max() <= 1, name
        assert ulp_diff(to_np(vg), to_np(vr)).max() <= 1, name
        np.testing.assert_allclose(to_np(pg), to_np(pr), rtol=2e-6, atol=1e-7, err_msg=name)
        if name == "masked_adam_upd":  # untouched where grad == 0
            z = (gr == 0).numpy()
            assert np.array_equal(to_np(pg)[z], p0.numpy()[z]) and np.array_equal(to_np(mg)[z], m0.numpy()[z])


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_alpha2weight_chunk_boundaries_and_early_stops(pkg, dtype):
    """Crafted rays around the 32-sample chunk boundary of the warp-per-ray kernels: segment lengths 1, 31, 32, 33, 64,
    65, 200 x constant alphas whose transmittance crosses 1e-3 after 6, 31, 32, 33, 34, 65 samples or never -- the early-stop
    index, the fills after it (T = 1, weight = 0) and the backward must equal the oracle's serial loop exactly
    (render_utils_kernel.cu:431-561), in the float32 and in the float64 instantiation."""
    from oracle import oracle_f64 as orc64
    o = orc if dtype == torch.float32 else orc64
    lengths = [1, 31, 32, 33, 64, 65, 200]
    # T after k samples = (1 - a)^k: first k with (1-a)^k < 1e-3 is ceil(ln 1e-3 / ln(1-a))
    alphas = [0.7, 0.2056, 0.1995, 0.1941, 0.1886, 0.1023, 0.0]     # kept samples: 6, 31, 32, 33, 34, 65, all
    ray_id, alpha = [], []
    r = 0
    for n in lengths:
        for a in alphas:
            ray_id += [r] * n
            alpha += [a] * n
            r += 2                      # odd ray ids stay empty
    n_rays = r
    ray_id = torch.tensor(ray_id, dtype=torch.int64)
    alpha = torch.tensor(alpha, dtype=dtype)
    g = torch.Generator().manual_seed(4)
    gw = torch.randn(alpha.numel(), generator=g, dtype=dtype)
    gl = torch.randn(n_rays, generator=g, dtype=dtype)
    ref = o.alpha2weight(alpha, ray_id, n_rays)
    got = pkg.render_utils_cuda.alpha2weight(alpha.to(DEV), ray_id.to(DEV), n_rays)
    stops = (ref[4] - ref[3])[::2].reshape(len(lengths), len(alphas))
    assert stops[-1].tolist() == [6, 31, 32, 33, 34, 65, 200], stops[-1].tolist()     # the crafted stops are really hit
    for name, x, y in zip(("weight", "T", "alphainv_last", "i_start", "i_end"), got, ref):
        assert x.dtype == y.dtype and np.array_equal(to_np(x), to_np(y)), name
    g_ref = o.alpha2weight_backward(alpha, *ref, n_rays, gw, gl)
    g_got = pkg.render_utils_cuda.alpha2weight_backward(alpha.to(DEV), *got, n_rays, gw.to(DEV), gl.to(DEV))
    assert np.array_equal(to_np(g_got), to_np(g_ref))
    seg_end = (ref[3] + torch.bincount(ray_id, minlength=n_rays)).tolist()
    past = torch.cat([torch.arange(int(s), int(e)) for s, e in zip(ref[4].tolist(), seg_end) if e > s])
    assert past.numel() > 100 and float(g_ref[past].abs().max()) == 0.0           # no gradient past the stop (:538)


def test_degenerate_shapes(pkg):
    """Empty ray batch, a ray that misses the box (>= 1 sample, render_utils_kernel.cu:46-47), a grid with size-1
    axes, Adam on a 4-byte-offset (not 16-byte aligned) view with N % 4 != 0 -- all against the oracle."""
    ru, tv, ad = pkg.render_utils_cuda, pkg.total_variation_cuda, pkg.adam_upd_cuda
    lo, hi = _box()
    e3 = torch.zeros(0, 3, device=DEV)
    out = ru.sample_pts_on_rays(e3, e3, lo.to(DEV), hi.to(DEV), 0.2, 6.0, 0.05)
    assert [tuple(t.shape) for t in out] == [(0, 3), (0,), (0,), (0,), (0,), (0,), (0,)]
    assert out[1].dtype == torch.bool and out[2].dtype == torch.int64 and out[4].dtype == torch.int64
    ro = torch.tensor([[5.0, 5.0, 5.0], [0.0, 0.0, -3.0]])
    rd = torch.tensor([[1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])       # ray 0 misses the box; ray 1 has two zero components
    ref = orc.sample_pts_on_rays(ro, rd, lo, hi, 0.2, 6.0, 0.05)
    got = ru.sample_pts_on_rays(ro.to(DEV), rd.to(DEV), lo.to(DEV), hi.to(DEV), 0.2, 6.0, 0.05)
    assert int(ref[4][0]) == 1 and bool(ref[1][0])               # one sample, outside the box
    for a, b in zip(got, ref):
        assert np.array_equal(to_np(a), to_np(b))
    g = torch.Generator().manual_seed(8)
    for shape in [(1, 2, 1, 5, 1), (1, 1, 1, 1, 1), (1, 3, 4, 1, 6)]:
        param = torch.randn(shape, generator=g)
        grad = torch.randn(shape, generator=g)
        grad.view(-1)[::2] = 0
        for dense in (False, True):
            r = grad.clone()
            orc.total_variation_add_grad(param, r, 0.3, 0.7, 1.3, dense)
            x = grad.clone().to(DEV)
            tv.total_variation_add_grad(param.to(DEV), x, 0.3, 0.7, 1.3, dense)
            assert ulp_diff(to_np(x), to_np(r)).max() <= 1, (shape, dense)
    N = 1027
    base = [torch.randn(N + 1, generator=g), torch.randn(N + 1, generator=g), torch.randn(N + 1, generator=g) * 0.01,
            torch.rand(N + 1, generator=g) * 1e-3]
    base[1][torch.rand(N + 1, generator=g) < 0.4] = 0
    cpu = [t.clone()[1:].contiguous() for t in base]
    dev = [t.clone().to(DEV)[1:] for t in base]                   # storage offset 4 bytes: the scalar path
    assert all(t.data_ptr() % 16 == 4 and t.is_contiguous() for t in dev)
    for step in (1, 2):
        orc.masked_adam_upd(cpu[0], cpu[1], cpu[2], cpu[3], step, 0.9, 0.99, 0.1, 1e-8)
        ad.masked_adam_upd(dev[0], dev[1], dev[2], dev[3], step, 0.9, 0.99, 0.1, 1e-8)
    for a, b in zip(dev, cpu):
        assert ulp_diff(to_np(a), to_np(b)).max() <= 1


def test_raw_c_abi_call(pkg):
    """Call the shared library directly (ctypes, raw device pointers, explicit stream) -- the path a
    non-torch host would take."""
    lib = ctypes.CDLL(pkg.LIB_PATH)
    lo, hi = _box()
    ro, rd, _, _ = make_rays(257, 77)
    ro_d, rd_d, lo_d, hi_d = ro.to(DEV), rd.to(DEV), lo.to(DEV), hi.to(DEV)
    t_min = torch.empty(257, device=DEV)
    t_max = torch.empty(257, device=DEV)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        rc = lib.dvgo_infer_t_minmax(ctypes.c_void_p(ro_d.data_ptr()), ctypes.c_void_p(rd_d.data_ptr()),
                                     ctypes.c_void_p(lo_d.data_ptr()), ctypes.c_void_p(hi_d.data_ptr()),
                                     ctypes.c_float(0.2), ctypes.c_float(6.0), ctypes.c_int(257),
                                     ctypes.c_void_p(t_min.data_ptr()), ctypes.c_void_p(t_max.data_ptr()),
                                     ctypes.c_void_p(stream.cuda_stream))
    assert rc == 0
    stream.synchronize()
    ref = orc.infer_t_minmax(ro, rd, lo, hi, 0.2, 6.0)
    assert np.array_equal(to_np(t_min), to_np(ref[0])) and np.array_equal(to_np(t_max), to_np(ref[1]))
