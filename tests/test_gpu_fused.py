"""GPU parity tests of the FUSED path (include/dvgo_b200_fused.h) against the CPU oracle model, the
golden outputs of the reference's own Python, and our own op-by-op path (itself parity-green
against the reference's CUDA kernels)."""
import copy
import os

import numpy as np
import pytest
import torch

from tests.util import rel_to_max, to_np, ulp_diff

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _fine_model(grid, seed=1, dens_scale=3.0, mask_p=0.3, **over):
    from directvoxgo_b200 import synthetic as syn
    from directvoxgo_b200.dvgo import DirectVoxGO
    lo, hi = syn.fine_bbox()
    kw = dict(syn.FINE_MODEL, num_voxels=grid ** 3, num_voxels_base=grid ** 3)
    kw.update(over)
    torch.manual_seed(0)
    m = DirectVoxGO(lo, hi, **kw)
    syn.randomize_grids_(m, seed)
    with torch.no_grad():
        m.density.mul_(dens_scale)
        if mask_p > 0:
            m.mask_cache.mask.copy_(torch.rand(m.mask_cache.mask.shape) > mask_p)
    return m


def test_layout_converters_roundtrip():
    from directvoxgo_b200 import ext
    g = torch.randn(1, 12, 9, 7, 5, device=DEV)
    cl = ext.ncdhw_to_cl(g)
    assert cl.shape == (9, 7, 5, 12)
    assert torch.equal(cl, g[0].permute(1, 2, 3, 0).contiguous())
    assert torch.equal(ext.cl_to_ncdhw(cl), g)


@pytest.mark.parametrize("C,tv,tv_dense,masked,perlr,shape", [
    (12, True, True, True, False, (11, 9, 13)), (12, True, False, True, False, (11, 9, 13)),
    (12, False, False, True, False, (11, 9, 13)), (12, False, False, False, False, (11, 9, 13)),
    (1, True, True, True, False, (11, 9, 13)), (1, False, False, False, True, (11, 9, 13)),
    (3, True, True, False, False, (11, 9, 13)), (9, True, False, True, False, (11, 9, 13)),
    # the row-staged (bulk-copy) sweep: more rows than CTAs x stages (stage reuse), a row that is split into two
    # segments with a z-halo (200 x 12 floats = 9.6 KB), 16 channels, and unmasked Adam
    (12, True, True, True, False, (23, 31, 16)), (12, True, False, True, False, (5, 4, 200)),
    (16, True, True, False, False, (7, 5, 150)), (4, True, True, True, False, (3, 2, 8))])
def test_sweep_matches_oracle_tv_plus_adam(C, tv, tv_dense, masked, perlr, shape):
    """The fused TV+Adam sweep on channel-last buffers == total_variation_add_grad followed by the
    matching Adam kernel on the reference's [1,C,X,Y,Z] layout (oracle), 3 consecutive steps."""
    from directvoxgo_b200 import ext
    from oracle import oracle as orc
    X, Y, Z = shape
    g = torch.Generator().manual_seed(C * 7 + tv)
    p = torch.randn(1, C, X, Y, Z, generator=g) * 1.5
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    pl = torch.rand(1, C, X, Y, Z, generator=g) if perlr else None
    to_cl = lambda t: t[0].permute(1, 2, 3, 0).contiguous().to(DEV)
    pc, pn, mc, vc = to_cl(p), torch.empty_like(to_cl(p)), to_cl(m), to_cl(v)
    plc = to_cl(pl) if perlr else None
    wx, wy, wz = 0.3, 0.7, 1.3
    for step in (1, 2, 3):
        grad = torch.randn(1, C, X, Y, Z, generator=g)
        grad[torch.rand(grad.shape, generator=g) < 0.5] = 0
        gc = to_cl(grad)
        # oracle
        if tv:
            orc.total_variation_add_grad(p, grad, wx, wy, wz, tv_dense)
        if perlr:
            orc.adam_upd_with_perlr(p, grad, m, v, pl, step, 0.9, 0.99, 0.1, 1e-8)
        elif masked:
            orc.masked_adam_upd(p, grad, m, v, step, 0.9, 0.99, 0.1, 1e-8)
        else:
            orc.adam_upd(p, grad, m, v, step, 0.9, 0.99, 0.1, 1e-8)
        # product
        out = pn if tv else pc
        ext.sweep(pc, out, gc, mc, vc, plc, X, Y, Z, C, tv, tv_dense, wx, wy, wz, masked and not perlr, step,
                  0.9, 0.99, 0.1, 1e-8)
        if tv:
            pc, pn = pn, pc
        assert torch.count_nonzero(gc) == 0            # gradient accumulator re-zeroed
        back = lambda t: t.permute(3, 0, 1, 2)[None].cpu()
        assert ulp_diff(to_np(back(mc)), to_np(m)).max() <= 1
        assert ulp_diff(to_np(back(vc)), to_np(v)).max() <= 1
        np.testing.assert_allclose(to_np(back(pc)), to_np(p), rtol=2e-6, atol=2e-7)


def _render_both(m, n_rays, seed):
    from directvoxgo_b200 import synthetic as syn
    from directvoxgo_b200.fused import FusedRenderer
    ro, rd, vd, _ = syn.random_training_rays(n_rays, n_views=20, seed=seed, device=DEV)
    rk = dict(syn.RENDER_KWARGS)
    with torch.no_grad():
        ref = m(ro, rd, vd, global_step=0, render_depth=True, **rk)
    out = FusedRenderer(m, rk, mlp="torch").render(ro, rd, vd, render_depth=True)
    return ref, out


@pytest.mark.parametrize("grid,n_rays,dens,maskp", [(32, 1024, 3.0, 0.3), (48, 4096, 1.0, 0.0), (40, 2048, 6.0, 0.5)])
def test_fused_render_matches_module_forward(grid, n_rays, dens, maskp):
    m = _fine_model(grid, dens_scale=dens, mask_p=maskp).to(DEV)
    ref, out = _render_both(m, n_rays, 3)
    # tolerances: T re-association 2e-6 rel; compositing order (atomics) 1e-5 abs on O(1) sums
    np.testing.assert_allclose(to_np(out["alphainv_last"]), to_np(ref["alphainv_last"]), rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(to_np(out["rgb_marched"]), to_np(ref["rgb_marched"]), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(to_np(out["depth"]), to_np(ref["depth"]), rtol=1e-5, atol=2e-3)


def test_fused_survivor_set_is_the_reference_sample_set():
    """The survivor stream (unordered) must be exactly the reference's M4 sample set: compare the
    multiset of (ray, step) and the per-sample weights."""
    from directvoxgo_b200 import synthetic as syn
    from directvoxgo_b200.fused import FusedRenderer
    m = _fine_model(36, dens_scale=4.0, mask_p=0.4).to(DEV)
    ro, rd, vd, _ = syn.random_training_rays(1500, n_views=20, seed=9, device=DEV)
    rk = dict(syn.RENDER_KWARGS)
    with torch.no_grad():
        ref = m(ro, rd, vd, global_step=0, **rk)
        pts, ray_id, step_id = m.sample_ray(ro, rd, **rk)
    fr = FusedRenderer(m, rk, mlp="torch")
    fr.render(ro, rd, vd)
    ws = fr._workspace(1500, False)
    m4 = int(ws.counters[0])
    assert m4 == ref["ray_id"].numel()
    assert int(ws.counters[1]) == 0
    ray = ws.s_ray[:m4].long()
    step = ws.s_slot[:m4].long() - ws.ray_off[:-1].long()[ray]
    key = ray * 100000 + step
    order = torch.argsort(key)
    assert torch.equal(ray[order], ref["ray_id"])                      # bit-exact set, ray-major order
    np.testing.assert_allclose(to_np(ws.s_weight[:m4][order]), to_np(ref["weights"]), rtol=5e-6, atol=1e-9)


@pytest.mark.parametrize("stage", ["fine", "coarse"])
def test_fused_trainer_vs_reference_python_golden(golden_dir, stage):
    from directvoxgo_b200.fused import FusedTrainer
    from tests.test_gpu_model import RK, _build
    g = np.load(os.path.join(golden_dir, "refpy_%s_small.npz" % stage), allow_pickle=False)
    m = _build(g, stage)
    ro, rd, vd, tgt = (torch.tensor(g[k]).to(DEV) for k in ("rays_o", "rays_d", "viewdirs", "target"))
    cfg = dict(fine=dict(weight_main=1.0, weight_entropy_last=1e-3, weight_rgbper=1e-2, lrate_density=0.1,
                         lrate_k0=0.1, lrate_rgbnet=1e-3, lrate_decay=1e9, skip_zero_grad_fields=["density", "k0"],
                         weight_tv_density=1e-5, weight_tv_k0=1e-5, tv_dense=True),
               coarse=dict(weight_main=1.0, weight_entropy_last=1e-2, weight_rgbper=0.1, lrate_density=0.1,
                           lrate_k0=0.1, lrate_rgbnet=0.0, lrate_decay=1e9, skip_zero_grad_fields=[]))[stage]
    tr = FusedTrainer(m, cfg, RK, mlp="torch")
    grads = {}
    orig = tr._optimise

    def capture(n_global):
        if not grads:
            grads["density"] = tr.g_density.clone()
            grads["k0"] = tr.g_k0.clone()
        orig(n_global)
    tr._optimise = capture
    l0 = float(tr.step(ro, rd, vd, tgt))
    l1 = float(tr.step(ro, rd, vd, tgt))
    assert abs(l0 - float(g["loss0"])) < 2e-6 and abs(l1 - float(g["loss1"])) < 5e-6
    gd = grads["density"].reshape(g["grad_density0"].shape)
    gk = grads["k0"].permute(3, 0, 1, 2)[None]
    assert rel_to_max(gd, g["grad_density0"]) < 1e-4          # atomics: rel 1e-4 of max-abs
    assert rel_to_max(gk, g["grad_k00"]) < 5e-4               # + cuBLAS-vs-MKL rgbnet backward
    tr.sync_to_model()
    for got, ref in ((m.density, g["density2"]), (m.k0, g["k02"])):
        d = np.abs(to_np(got) - ref)
        assert np.quantile(d, 0.999) < 2e-3 and np.median(d) < 1e-5


def test_fused_trainer_vs_module_trainer_bigger():
    """48^3 fine grid, 4096 Blender rays, 3 steps: fused (torch-MLP mode) vs the op-by-op path."""
    from directvoxgo_b200 import synthetic as syn
    from directvoxgo_b200.fused import FusedTrainer
    from directvoxgo_b200.trainer import ModuleTrainer
    m1 = _fine_model(48, dens_scale=2.0, mask_p=0.2).to(DEV)
    m2 = copy.deepcopy(m1)
    cfg, rk = dict(syn.FINE_TRAIN), dict(syn.RENDER_KWARGS)
    t1, t2 = ModuleTrainer(m1, cfg, rk), FusedTrainer(m2, cfg, rk, mlp="torch")
    for it in range(3):
        ro, rd, vd, tgt = syn.random_training_rays(4096, n_views=20, seed=50 + it, device=DEV)
        la, lb = float(t1.step(ro, rd, vd, tgt)), float(t2.step(ro, rd, vd, tgt))
        assert abs(la - lb) < 1e-5 * max(1.0, abs(la)), (it, la, lb)
    t2.sync_to_model()
    for a, b in ((m1.density, m2.density), (m1.k0, m2.k0)):
        d = np.abs(to_np(a) - to_np(b))
        # 3 Adam steps of size lr=0.1 each: m/sqrt(v) amplifies 1e-4-relative gradient differences (atomics)
        assert np.median(d) < 1e-4 and np.quantile(d, 0.999) < 5e-3
    for pa, pb in zip(m1.rgbnet.parameters(), m2.rgbnet.parameters()):
        assert rel_to_max(pa, pb) < 1e-3


def test_fused_dmpigo_render_and_step(golden_dir):
    from directvoxgo_b200.dmpigo import DirectMPIGO
    from directvoxgo_b200.fused import FusedRenderer
    g = np.load(os.path.join(golden_dir, "refpy_dmpigo_small.npz"), allow_pickle=False)
    m = DirectMPIGO(xyz_min=g["xyz_min"], xyz_max=g["xyz_max"], num_voxels=20 * 18 * 16, mpi_depth=16,
                    fast_color_thres=1e-3, rgbnet_dim=9, rgbnet_depth=3, rgbnet_width=64, viewbase_pe=0)
    with torch.no_grad():
        m.density.copy_(torch.tensor(g["density0"]))
        m.k0.copy_(torch.tensor(g["k00"]))
        lin = [x for x in m.rgbnet.modules() if isinstance(x, torch.nn.Linear)]
        for i, l in enumerate(lin):
            l.weight.copy_(torch.tensor(g["rgbnet_w%d" % i]))
            l.bias.copy_(torch.tensor(g["rgbnet_b%d" % i]))
    m = m.to(DEV)
    ro, rd, vd = (torch.tensor(g[k]).to(DEV) for k in ("rays_o", "rays_d", "viewdirs"))
    out = FusedRenderer(m, dict(near=0, far=1, bg=0.0, stepsize=0.5), mlp="torch").render(ro, rd, vd)
    np.testing.assert_allclose(to_np(out["rgb_marched"]), g["out_rgb_marched"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(to_np(out["alphainv_last"]), g["out_alphainv_last"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(to_np(out["depth"]), g["out_depth"], rtol=1e-5, atol=1e-3)
    # the 64-wide rgbnet on the tensor-core kernels (zero-padded to 128): fp16-operand tolerance
    r_tc = FusedRenderer(m, dict(near=0, far=1, bg=0.0, stepsize=0.5))
    assert r_tc.mlp_mode == "tc"
    out_tc = r_tc.render(ro, rd, vd)
    np.testing.assert_allclose(to_np(out_tc["rgb_marched"]), g["out_rgb_marched"], rtol=0, atol=2e-3)
    np.testing.assert_allclose(to_np(out_tc["alphainv_last"]), g["out_alphainv_last"], rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("n_rays", [1, 33, 1000])
def test_fused_trainer_ragged_ray_counts_and_misses(n_rays):
    """Ray counts that are not multiples of the warp / tile size, and rays that miss the box entirely
    (every ray still emits >= 1 sample, render_utils_kernel.cu:46-47, which the bbox mask removes)."""
    from directvoxgo_b200 import synthetic as syn
    from directvoxgo_b200.fused import FusedTrainer
    from directvoxgo_b200.trainer import ModuleTrainer
    m1 = _fine_model(32, dens_scale=2.0, mask_p=0.2).to(DEV)
    m2 = copy.deepcopy(m1)
    cfg, rk = dict(syn.FINE_TRAIN), dict(syn.RENDER_KWARGS)
    ro, rd, vd, tgt = syn.random_training_rays(n_rays, n_views=7, seed=11, device=DEV)
    rd[: max(1, n_rays // 3)] *= -1.0          # these rays point away from the scene
    vd = rd / rd.norm(dim=-1, keepdim=True)
    la = float(ModuleTrainer(m1, cfg, rk).step(ro, rd, vd, tgt))
    for mode in ("torch", "tc"):
        m = copy.deepcopy(m2)
        lb = float(FusedTrainer(m, cfg, rk, mlp=mode).step(ro, rd, vd, tgt))
        assert abs(la - lb) < (1e-5 if mode == "torch" else 2e-3) * max(1.0, abs(la)), (mode, la, lb)


def test_fused_all_rays_miss():
    from directvoxgo_b200 import synthetic as syn
    from directvoxgo_b200.fused import FusedRenderer, FusedTrainer
    m = _fine_model(32).to(DEV)
    ro, rd, vd, tgt = syn.random_training_rays(257, n_views=5, seed=2, device=DEV)
    rd = -rd
    vd = -vd
    out = FusedRenderer(m, dict(syn.RENDER_KWARGS)).render(ro, rd, vd)
    assert torch.all(out["alphainv_last"] == 1) and torch.allclose(out["rgb_marched"], torch.ones_like(out["rgb_marched"]))
    assert torch.all(out["depth"] == 0)
    tr = FusedTrainer(m, dict(syn.FINE_TRAIN), dict(syn.RENDER_KWARGS))
    loss = float(tr.step(ro, rd, vd, tgt))
    ref = float(((1.0 - tgt) ** 2).mean()) + 1e-3 * float(-(torch.tensor(1 - 1e-6).log() * (1 - 1e-6) + torch.tensor(1e-6).log() * 1e-6))
    assert abs(loss - ref) < 1e-5
    assert int(tr._workspace(257, True).counters[0]) == 0


def test_fused_full_size_invariants():
    """BASELINE size (160^3, 8192 rays, ~2.4 M samples): size-independent properties of the fused forward --
    per-ray sum of weights + alphainv_last == 1 (render_utils_kernel.cu:445-457), survivor count equal to the
    op-by-op path's sample count, rgb in [0,1], slot bookkeeping consistent."""
    from directvoxgo_b200 import synthetic as syn
    from directvoxgo_b200.fused import FusedRenderer
    m = _fine_model(160, dens_scale=1.0, mask_p=0.0).to(DEV)
    ro, rd, vd, _ = syn.random_training_rays(8192, n_views=100, seed=1000, device=DEV)
    rk = dict(syn.RENDER_KWARGS)
    fr = FusedRenderer(m, rk)
    out = fr.render(ro, rd, vd)
    ws = fr._workspace(8192, False)
    m4 = int(ws.counters[0])
    assert int(ws.counters[1]) == 0 and 2_000_000 < m4 < 3_000_000
    with torch.no_grad():
        ref = m(ro, rd, vd, global_step=0, **rk)
    # The op-by-op path's alpha2weight is bit-exact with the reference (float T_cum re-rounded per sample); the fused
    # march keeps a double product scan (T rel 5e-6), so a sample whose weight sits within that
    # rounding of the 1e-4 threshold may fall on the other side: a handful out of 2.4 M at most.
    assert abs(m4 - ref["ray_id"].numel()) <= 4
    wsum = torch.zeros(8192, device=DEV).index_add_(0, ws.s_ray[:m4].long(), ws.s_weight[:m4])
    # samples below the weight threshold (1e-4) are dropped from the stream: allow their mass
    n_steps = ws.n_steps[:8192].float()
    assert torch.all((wsum + out["alphainv_last"] - 1).abs() < 1e-4 * n_steps + 1e-4)
    assert float(out["rgb_marched"].min()) >= 0 and float(out["rgb_marched"].max()) <= 1 + 1e-5
    # the per-slot record only exists in a training workspace (a rendering call keeps nothing for march_bwd)
    assert ws.slot_code.numel() == 0
    wt = fr._workspace(8192, True)
    fr._march(wt, ro, rd)
    codes = wt.slot_code[: int(wt.ray_off[8192])]
    assert int(wt.counters[0]) == m4 and int((codes >= 0).sum()) == m4 and int(codes.max()) == m4 - 1
    np.testing.assert_allclose(to_np(out["rgb_marched"]), to_np(ref["rgb_marched"]), rtol=0, atol=2e-3)


@pytest.mark.parametrize("dens,maskp", [(1.0, 0.0), (6.0, 0.4)])
def test_exact_transmittance_mode_is_bit_exact_at_full_size(dens, maskp):
    """scene.exact_transmittance: the fused march replays the reference's per-sample `float T_cum` recurrence
    (render_utils_kernel.cu:447-451) instead of its double product scan.  At BASELINE size (160^3, 8192 rays) the
    survivor SET, the weights and alphainv_last must then EQUAL those of the op-by-op path, whose alpha2weight is
    itself bit-exact with the reference kernel (test_gpu_0_vs_ref.py::test_alpha_ops_vs_reference_kernels) -- no
    tolerance.  Also times both march variants (reported in gpurun_out/exact_T_cost.json, not asserted)."""
    import json
    from directvoxgo_b200 import synthetic as syn
    from directvoxgo_b200.fused import FusedRenderer
    m = _fine_model(160, dens_scale=dens, mask_p=maskp).to(DEV)
    ro, rd, vd, _ = syn.random_training_rays(8192, n_views=100, seed=1000, device=DEV)
    rk = dict(syn.RENDER_KWARGS)
    with torch.no_grad():
        ref = m(ro, rd, vd, global_step=0, **rk)
    times = {}
    for exact in (False, True):
        fr = FusedRenderer(m, rk, mlp="torch", exact_transmittance=exact)
        out = fr.render(ro, rd, vd)
        ws = fr._workspace(8192, False)
        m4 = int(ws.counters[0])
        assert int(ws.counters[1]) == 0
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        wt = fr._workspace(8192, True)
        for _ in range(3):
            fr._march(wt, ro, rd)
        torch.cuda.synchronize()
        ev[0].record()
        for _ in range(10):
            fr._march(wt, ro, rd)
        ev[1].record()
        torch.cuda.synchronize()
        times["exact" if exact else "scan"] = ev[0].elapsed_time(ev[1]) / 10
        if not exact:
            assert abs(m4 - ref["ray_id"].numel()) <= 4
            continue
        assert m4 == ref["ray_id"].numel() and m4 > 100000
        ray = ws.s_ray[:m4].long()
        step = ws.s_slot[:m4].long() - ws.ray_off[:-1].long()[ray]
        order = torch.argsort(ray * 100000 + step)
        assert torch.equal(ray[order], ref["ray_id"])                        # the same sample set
        assert torch.equal(ws.s_weight[:m4][order], ref["weights"])          # the same weights, bit for bit
        assert torch.equal(out["alphainv_last"], ref["alphainv_last"])
    times.update(survivors=m4, what="ray_setup + march_fwd + k0_gather of one 8192-ray training batch at 160^3, ms")
    try:
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "exact_T_cost.json")
        os.makedirs(os.path.dirname(path), exist_ok=True)
        old = json.load(open(path)) if os.path.exists(path) else {}
        old["dens%g_mask%g" % (dens, maskp)] = times
        json.dump(old, open(path, "w"), indent=1)
    except OSError:
        pass
    print("march stage ms:", times)


def test_exact_transmittance_trainer_tracks_the_module_path():
    """FusedTrainer(exact_transmittance=True): three training steps against the op-by-op ModuleTrainer from the same
    state -- same tolerances as the default mode (the gradients still go through atomics)."""
    from directvoxgo_b200 import synthetic as syn
    from directvoxgo_b200.fused import FusedTrainer
    from directvoxgo_b200.trainer import ModuleTrainer
    m1 = _fine_model(40, dens_scale=3.0, mask_p=0.3).to(DEV)
    m2 = copy.deepcopy(m1)
    cfg, rk = dict(syn.FINE_TRAIN), dict(syn.RENDER_KWARGS)
    t1, t2 = ModuleTrainer(m1, cfg, rk), FusedTrainer(m2, cfg, rk, mlp="torch", exact_transmittance=True)
    for it in range(3):
        ro, rd, vd, tgt = syn.random_training_rays(2048, n_views=20, seed=50 + it, device=DEV)
        la, lb = float(t1.step(ro, rd, vd, tgt)), float(t2.step(ro, rd, vd, tgt))
        assert abs(la - lb) < 1e-5 * max(1.0, abs(la)), (it, la, lb)
    t2.sync_to_model()
    for a, b in ((m1.density, m2.density), (m1.k0, m2.k0)):
        d = np.abs(to_np(a) - to_np(b))
        assert np.median(d) < 1e-4 and np.quantile(d, 0.999) < 5e-3


def _composite_check(keys, n_rays, seed):
    """composite_kernel on a crafted survivor stream vs index_add_ (lib/dvgo.py:554-559, 569-576)."""
    from directvoxgo_b200 import ext
    g = torch.Generator().manual_seed(seed)
    keys = torch.as_tensor(keys, dtype=torch.int32)
    n = keys.numel()
    w = torch.rand(n, generator=g)
    rgb = torch.rand(n, 3, generator=g)
    ray_off = (torch.arange(n_rays + 1, dtype=torch.int32) * 1000)
    step = torch.randint(0, 1000, (n,), generator=g, dtype=torch.int32)
    slot = ray_off[keys.long()] + step
    cap = n + 37                                       # capacity > count: the tail must be ignored
    pad = lambda t, fill: torch.cat([t, torch.full((cap - n,) + t.shape[1:], fill, dtype=t.dtype)]).to(DEV)
    counters = torch.tensor([n, 0], dtype=torch.int32, device=DEV)
    rgb_acc = torch.zeros(n_rays, 3, device=DEV)
    depth_acc = torch.zeros(n_rays, device=DEV)
    ext.composite(pad(rgb, 7.0), pad(w, 7.0), pad(keys, 0), pad(slot, 5), ray_off.to(DEV), counters, rgb_acc, depth_acc)
    want_rgb = torch.zeros(n_rays, 3, dtype=torch.float64).index_add_(0, keys.long(), (w[:, None] * rgb).double())
    want_d = torch.zeros(n_rays, dtype=torch.float64).index_add_(0, keys.long(), (w * step.float()).double())
    assert torch.allclose(rgb_acc.cpu().double(), want_rgb, rtol=1e-5, atol=1e-5), \
        float((rgb_acc.cpu().double() - want_rgb).abs().max())
    assert torch.allclose(depth_acc.cpu().double(), want_d, rtol=1e-5, atol=1e-3)


def test_composite_adversarial_streams():
    """Chunks of one ray get their stream positions from an atomicAdd (march_fwd), so equal ray ids need not be
    contiguous: A A B B A A inside one 32-lane window must stay three runs (round-1 bug: key equality merged the two
    A runs and double-counted).  Crafted patterns + random interleavings of short runs."""
    _composite_check([7, 7, 3, 3, 7, 7, 9, 9, 9, 3, 7, 3, 7], 10, 0)
    _composite_check([0, 1] * 64, 2, 1)                                   # single-element runs, period 2
    _composite_check([0, 1, 2, 3] * 40 + [1] * 5, 4, 2)                   # period 4 (shuffle distance 4 matches)
    _composite_check([5] * 30 + [2] * 4 + [5] * 30 + [2] * 70 + [5], 6, 3)  # runs straddling 32-lane boundaries
    _composite_check([4], 5, 4)
    _composite_check(list(range(31, -1, -1)) * 3, 32, 5)
    rng = np.random.default_rng(0)
    for trial in range(8):
        n_rays = int(rng.integers(2, 9))
        runs = rng.integers(1, 20, size=400)
        ids = rng.integers(0, n_rays, size=400)
        keys = np.repeat(ids, runs)
        _composite_check(keys, n_rays, 10 + trial)


def test_sphere_scene_trainer_rgb_matches_module_path():
    """Sparse scene (most chunks hold a few survivors, so interleaved runs are the norm): the fused forward's
    rgb_marched / depth equal the op-by-op module path on the same rays, repeated to catch order nondeterminism."""
    from directvoxgo_b200 import synthetic as syn
    from directvoxgo_b200.dvgo import MaskCache
    from directvoxgo_b200.fused import FusedRenderer
    m = _fine_model(64, seed=2, dens_scale=1.0, mask_p=0.0).to(DEV)
    with torch.no_grad():
        ax = torch.linspace(-1, 1, 64, device=DEV)
        r = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), -1).norm(dim=-1)
        m.density.copy_(torch.where(r < 0.6, 5.0, -5.0)[None, None])
        alpha = torch.nn.functional.max_pool3d(m.activate_density(m.density), 3, 1, 1)[0, 0]
        m.mask_cache = MaskCache(mask=(alpha > m.fast_color_thres), xyz_min=m.xyz_min, xyz_max=m.xyz_max).to(DEV)
    rk = dict(syn.RENDER_KWARGS)
    ro, rd, vd, _ = syn.random_training_rays(4096, n_views=20, seed=9, device=DEV)
    with torch.no_grad():
        want = m(ro, rd, vd, global_step=None, render_depth=True, **rk)
    r = FusedRenderer(m, rk, mlp="torch")
    for _ in range(10):
        out = r.render(ro, rd, vd)
        assert torch.allclose(out["rgb_marched"], want["rgb_marched"], atol=2e-5), \
            float((out["rgb_marched"] - want["rgb_marched"]).abs().max())
        assert torch.allclose(out["depth"], want["depth"], rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("rgbnet_dim,n_rays", [(12, 1500), (9, 700), (4, 333), (16, 1000)])
def test_survivor_tile_producers_match_pack_kernels(rgbnet_dim, n_rays):
    """The X~ tiles written by k0_gather_tiles are byte-identical to k0_gather -> mlp_pack_x, the dZ3 tiles written by
    sample_grad to mlp_pack_dz of its fp32 outputs, and both are zero from the survivor count to the end of the pair."""
    from directvoxgo_b200 import ext, synthetic as syn
    from directvoxgo_b200.dvgo import DirectVoxGO
    from directvoxgo_b200.fused import FusedRenderer
    lo, hi = syn.fine_bbox()
    kw = dict(syn.FINE_MODEL, num_voxels=36 ** 3, num_voxels_base=36 ** 3, rgbnet_dim=rgbnet_dim)
    torch.manual_seed(0)
    m = DirectVoxGO(lo, hi, **kw)
    syn.randomize_grids_(m, 3)
    with torch.no_grad():
        m.density.mul_(3.0)
    m = m.to(DEV)
    rk = dict(syn.RENDER_KWARGS)
    ro, rd, vd, _ = syn.random_training_rays(n_rays, n_views=20, seed=5, device=DEV)
    fr = FusedRenderer(m, rk, mlp="tc")
    fr.fuse_gather = False          # this test checks the stand-alone tile producer (k0_gather_tiles)
    ws = fr._workspace(n_rays, False)
    pe, pe16 = fr._tc_embed(vd)
    C, ps = rgbnet_dim, pe.shape[1]
    P = 3 + 6 * int(fr._viewfreq().numel())
    ws.tiles(C, ps, False)                              # allocated zeroed: padding-only chunks are never written
    K1 = (C + ps + 15) // 16 * 16

    def row_of(t, p):
        return np.array([t[(p // 128) * 128 * K1 + (((p % 128) // 8) * (K1 // 8) * 128 + (c // 8) * 128 + (p % 8) * 16 + (c % 8) * 2) // 2]
                         for c in range(K1)], np.float32)

    m4_prev = 0
    for n_use in (n_rays, n_rays // 2):                 # second pass: fewer survivors over the first pass's stale rows
        fr._march(ws, ro[:n_use].contiguous(), rd[:n_use].contiguous(), (pe, pe16))
        torch.cuda.synchronize()
        got = ws.xt.clone()
        m4 = int(ws.counters[0])
        assert m4 > 256 and m4 != m4_prev
        # the fp32 feature stream of the SAME survivor stream (its order is not reproducible across marches)
        ext.k0_gather(fr.scene, ro[:n_use].contiguous(), rd[:n_use].contiguous(), fr.k0, ws.t_min, ws.ray_off, ws.s_ray,
                      ws.s_slot, ws.counters, ws.feat)
        want = torch.zeros_like(got)
        ext.mlp_pack_x(ws.feat, ws.s_ray, pe, P, ws.counters, want)
        torch.cuda.synchronize()
        used = (m4 + 255) // 256 * 256 * K1 * 2         # bytes of the tiles up to the end of the last pair
        assert torch.equal(got[:used], want[:used])
        if m4_prev == 0:
            assert torch.all(got[used:] == 0)           # nothing written past the pair
        t = got[:used].cpu().numpy().view(np.float16)
        for p in (0, 7, 129, m4 - 1):                   # decode the operand layout on the host
            row = row_of(t, p)
            x = torch.cat([ws.feat[p], pe[int(ws.s_ray[p])]]).half().float().cpu().numpy()
            np.testing.assert_array_equal(row[:C + ps], x)
            assert np.all(row[C + ps:] == 0)
        for p in range(m4, (m4 + 255) // 256 * 256, 37):  # rows between the count and the end of the pair are zero
            assert np.all(row_of(t, p) == 0)
        m4_prev = m4

    # dZ3 tiles from sample_grad
    g = torch.Generator().manual_seed(1)
    cap = ws.cap
    rgb = torch.rand(cap, 3, generator=g).to(DEV)
    w = torch.rand(cap, generator=g).to(DEV)
    G = (torch.randn(n_rays, 3, generator=g) / n_rays).to(DEV)
    tgt = torch.rand(n_rays, 3, generator=g).to(DEV)
    d_rgb, d_w = torch.empty(cap, 3, device=DEV), torch.empty(cap, device=DEV)
    loss = torch.zeros(2, device=DEV)
    dzt = torch.zeros(ext.mlp_dztile_bytes(cap), dtype=torch.uint8, device=DEV)
    dzt.view(-1, 128)[0::2] = 0xCD                      # stale bytes in every written chunk (columns 0..7 of all rows)
    S = 2.0 ** 18
    ext.sample_grad(rgb, w, ws.s_ray, G, tgt, ws.counters, n_rays, 0.01, d_rgb, d_w, loss, dzt, S)
    want = torch.zeros_like(dzt)
    ext.mlp_pack_dz(rgb, d_rgb, S, ws.counters, want)
    torch.cuda.synchronize()
    used = (m4 + 255) // 256 * 256 * 32
    assert torch.equal(dzt[:used], want[:used])
    assert torch.all(dzt[used:].view(-1, 128)[0::2] == 0xCD)       # nothing written past the pair
    z = (d_rgb[:m4] * rgb[:m4] * (1 - rgb[:m4]) * S).half().cpu().numpy()
    t = dzt[:used].cpu().numpy().view(np.float16)
    for p in (0, 200, m4 - 1):
        off = (p // 128) * 2048 + ((p % 128) // 8) * 128 + (p % 8) * 8
        np.testing.assert_array_equal(t[off:off + 3], z[p])
        assert np.all(t[off + 3:off + 8] == 0) and np.all(t[off + 64:off + 72] == 0)


@pytest.mark.parametrize("C", [12, 9, 4])
def test_gather_fused_into_rgbnet_forward_matches_two_kernel_path(C):
    """mlp_fwd_gather_kernel (producer warps build the X~ tile in shared memory) against k0_gather_tiles + mlp_fwd on
    the SAME survivor stream (one march; the stream order differs from run to run): byte-identical tiles over the tile
    pairs the backward kernel reads, hence bit-identical rgb, in the rendering form (no tiles out) and the training
    form (tiles copied out); and the rendered image agrees with the two-kernel renderer (lib/dvgo.py:509, :536-539)."""
    from directvoxgo_b200 import ext, synthetic as syn
    from directvoxgo_b200.fused import FusedRenderer
    m = _fine_model(48, dens_scale=1.0, mask_p=0.3, rgbnet_dim=C).to(DEV)
    ro, rd, vd, _ = syn.random_training_rays(3000, n_views=20, seed=3, device=DEV)
    rk = dict(syn.RENDER_KWARGS)
    fr = FusedRenderer(m, rk)
    assert fr.mlp_mode == "tc"
    imgs = []
    for fuse in (True, False):
        fr.fuse_gather = fuse
        imgs.append(fr.render(ro, rd, vd)["rgb_marched"].clone())
    assert int(fr._workspace(3000, False).counters[1]) == 0
    np.testing.assert_allclose(to_np(imgs[0]), to_np(imgs[1]), rtol=0, atol=1e-5)

    pe = fr._tc_embed(vd)
    ws = fr._workspace(3000, True)
    fr.fuse_gather = True
    fr._march(ws, ro, rd, pe)                      # march only: the gather is left to the forward kernel
    m4 = int(ws.counters[0])
    assert m4 > 1000
    pe_stride = pe[0].shape[1]
    xt_two = torch.zeros_like(ws.tiles(C, pe_stride, True))
    ext.k0_gather_tiles(fr.scene, ro, rd, fr.k0, ws.t_min, ws.ray_off, ws.s_ray, ws.s_slot, ws.counters, ws.s_pos,
                        pe[1], pe_stride, xt_two)
    rgb_two = torch.zeros_like(ws.rgb)
    fr._tc.forward_tiles(xt_two, C, pe_stride, ws.counters, ws.cap, rgb_two)
    xt_fused = torch.full_like(xt_two, 0xAB)       # every byte of the pairs in use must be written by the kernel
    rgb_train, rgb_render = torch.zeros_like(ws.rgb), torch.zeros_like(ws.rgb)
    fr._tc.forward_gather(fr.scene, fr.k0, ws.s_pos, pe[1], C, pe_stride, ws.counters, ws.cap, rgb_train, xt_fused)
    fr._tc.forward_gather(fr.scene, fr.k0, ws.s_pos, pe[1], C, pe_stride, ws.counters, ws.cap, rgb_render, None)
    tile_bytes = xt_two.numel() // ((ws.cap + 127) // 128 + 1)
    used = (m4 + 255) // 256 * 2 * tile_bytes
    assert torch.equal(xt_fused[:used], xt_two[:used])
    assert torch.equal(rgb_train[:m4], rgb_two[:m4]) and torch.equal(rgb_render[:m4], rgb_two[:m4])


def test_host_fed_loop_returns_every_steps_loss_one_call_late():
    """trainer.HostFedLoop (H2D copy of step i under step i-1, loss of step i read after step i+1 is enqueued) yields
    exactly the losses of the serialised copy -> step -> read loop, in order, and leaves the same parameters."""
    import copy
    from directvoxgo_b200 import synthetic as syn
    from directvoxgo_b200.fused import FusedTrainer
    from directvoxgo_b200.trainer import HostFedLoop
    m1 = _fine_model(40, dens_scale=2.0, mask_p=0.2).to(DEV)
    m2 = copy.deepcopy(m1)
    cfg, rk = dict(syn.FINE_TRAIN), dict(syn.RENDER_KWARGS)
    cfg["N_rand"] = 2048
    host = [tuple(t.pin_memory() for t in syn.random_training_rays(2048, n_views=10, seed=900 + b, device="cpu"))
            for b in range(5)]
    t1 = FusedTrainer(m1, cfg, rk, mlp="tc")
    want = [float(t1.step(*[x.to(DEV) for x in b]).item()) for b in host]
    t2 = FusedTrainer(m2, cfg, rk, mlp="tc")
    loop = HostFedLoop(t2, host[0])
    got = []
    for b in host:
        prev = loop.step(b)
        if prev is not None:
            got.append(prev)
    got.append(loop.drain())
    assert len(got) == len(want)
    # same kernels on the same inputs; the grid gradients are accumulated with unordered atomics
    np.testing.assert_allclose(np.array(got), np.array(want), rtol=2e-4, atol=1e-6)
    t1.sync_to_model(); t2.sync_to_model()
    d = np.abs(to_np(m1.density) - to_np(m2.density))
    assert np.quantile(d, 0.999) < 5e-3
