"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/refpy_prep.npz and tests/golden/ref_ckpt_fine_last.tar by
running the reference's OWN Python for the rows either side of the per-iteration path:

  N2  lib/ray_utils.py  get_rays_of_a_view (all flag combinations, with and without NDC)
  N1  lib/ray_utils.py  get_training_rays_in_maskcache_sampling + lib/dvgo.py hit_coarse_geo
  N3  lib/dvgo.py       voxel_count_views, scale_volume_grid; run.py:330-332 occupancy refresh
  N4  run.py:420-437    the `{stage}_last.tar` checkpoint (model_kwargs + state_dict + MaskedAdam state)

As in make_golden_refpy.py the reference's CUDA extensions are served by the CPU oracle; everything else is
the unmodified reference code on torch-CPU.   Run from the repo root:  python -m oracle.make_golden_prep
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

from oracle.make_golden_refpy import OUT, REF, import_reference


def pose(theta, phi, radius):
    """A camera on a sphere looking at the origin (OpenGL convention, like lib/load_blender.py:37-42)."""
    t, p = np.deg2rad(theta), np.deg2rad(phi)
    c = np.array([np.cos(p) * np.sin(t), np.cos(p) * np.cos(t), -np.sin(p)]) * radius
    z = c / np.linalg.norm(c)
    x = np.cross([0, 0, 1.0], z); x /= np.linalg.norm(x)
    y = np.cross(z, x)
    m = np.eye(4, dtype=np.float32)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = x, y, z, c
    return torch.tensor(m)


def main():
    ref_dvgo, _, ref_adam = import_reference()
    sys.path.insert(0, REF)
    import lib.ray_utils as ru
    torch.manual_seed(777)
    np.random.seed(777)
    save = {}

    # ---- N2: rays of a view ---------------------------------------------------------------------------------
    H, W = 23, 31
    K = np.array([[40.5, 0, 0.5 * W], [0, 41.25, 0.5 * H], [0, 0, 1]], dtype=np.float32)
    c2w = pose(35.0, -30.0, 4.0)
    save["view_K"], save["view_c2w"], save["view_HW"] = K, c2w.numpy(), np.array([H, W])
    combos = []
    for ndc in (False, True):
        for inverse_y in (False, True):
            for flip_x, flip_y in ((False, False), (True, False), (False, True)):
                for mode in ("center", "lefttop"):
                    combos.append((ndc, inverse_y, flip_x, flip_y, mode))
    save["view_combos"] = np.array([[int(a), int(b), int(c), int(d), int(m == "center")] for a, b, c, d, m in combos])
    for k, (ndc, inverse_y, flip_x, flip_y, mode) in enumerate(combos):
        ro, rd, vd = ru.get_rays_of_a_view(H, W, K, c2w, ndc, inverse_y, flip_x, flip_y, mode)
        save["view%d_o" % k] = ro.contiguous().numpy().copy()
        save["view%d_d" % k] = rd.numpy().copy()
        save["view%d_v" % k] = vd.numpy().copy()

    # ---- a small coarse-stage model with a blobby occupancy mask ---------------------------------------------------
    lo, hi = np.array([-1.0, -0.9, -0.8], np.float32), np.array([1.0, 0.9, 0.8], np.float32)
    kw = dict(num_voxels=20 ** 3, num_voxels_base=20 ** 3, alpha_init=1e-2, fast_color_thres=1e-4,
              rgbnet_dim=12, rgbnet_direct=True, rgbnet_depth=3, rgbnet_width=64, viewbase_pe=4)
    model = ref_dvgo.DirectVoxGO(xyz_min=lo, xyz_max=hi, **kw)
    with torch.no_grad():
        model.density.copy_(torch.randn(model.density.shape) * 3.0 + 1.0)
        model.k0.copy_(torch.randn(model.k0.shape))
        gx = torch.stack(torch.meshgrid(*[torch.linspace(-1, 1, s) for s in model.mask_cache.mask.shape], indexing="ij"), -1)
        mask = ((gx - torch.tensor([0.2, -0.1, 0.0])).norm(dim=-1) < 0.45) | ((gx + 0.5).norm(dim=-1) < 0.25)
        model.mask_cache.mask.copy_(mask)
    save.update(xyz_min=lo, xyz_max=hi, density0=model.density.detach().numpy().copy(),
                k00=model.k0.detach().numpy().copy(), mask0=model.mask_cache.mask.numpy().copy(),
                act_shift=np.float64(model.act_shift))
    rk = dict(near=0.5, far=6.0, bg=1.0, stepsize=0.5)

    # ---- N1: training-ray preparation ----------------------------------------------------------------------------------
    H2, W2 = 24, 32
    K2 = np.array([[34.0, 0, 0.5 * W2], [0, 34.0, 0.5 * H2], [0, 0, 1]], dtype=np.float32)
    poses = torch.stack([pose(t, p, 3.2) for t, p in ((10.0, -20.0), (130.0, -45.0), (250.0, -10.0))])
    imgs = [torch.rand(H2, W2, 3) for _ in poses]
    HW = np.array([[H2, W2]] * len(poses))
    Ks = np.stack([K2] * len(poses))
    save.update(tr_poses=poses.numpy(), tr_K=K2, tr_HW=HW, tr_imgs=torch.stack(imgs).numpy())
    ro, rd, vd = ru.get_rays_of_a_view(H2, W2, K2, poses[0], False, False, False, False)
    save["hit0"] = model.hit_coarse_geo(rays_o=ro, rays_d=rd, **rk).numpy().copy()
    rgb_tr, ro_tr, rd_tr, vd_tr, imsz = ru.get_training_rays_in_maskcache_sampling(
        rgb_tr_ori=imgs, train_poses=poses, HW=HW, Ks=Ks, ndc=False, inverse_y=False, flip_x=False, flip_y=False,
        model=model, render_kwargs=rk)
    save.update(mc_rgb=rgb_tr.numpy().copy(), mc_o=ro_tr.numpy().copy(), mc_d=rd_tr.numpy().copy(),
                mc_v=vd_tr.numpy().copy(), mc_imsz=np.array([int(n) for n in imsz]))

    # ---- N3: voxel_count_views ------------------------------------------------------------------------------------------
    _, ro_all, rd_all, _, imsz_all = ru.get_training_rays(torch.stack(imgs), poses, HW, Ks, False, False, False, False)
    cnt = model.voxel_count_views(rays_o_tr=ro_all, rays_d_tr=rd_all, imsz=imsz_all, near=rk["near"], far=rk["far"],
                                  stepsize=rk["stepsize"], downrate=1)
    save["count_views"] = cnt.numpy().copy()

    # ---- N3: occupancy refresh (run.py:330-332) ---------------------------------------------------------------------
    with torch.no_grad():
        self_alpha = F.max_pool3d(model.activate_density(model.density), kernel_size=3, padding=1, stride=1)[0, 0]
        refreshed = model.mask_cache.mask & (self_alpha > model.fast_color_thres)
    save["mask_refreshed"] = refreshed.numpy().copy()

    # ---- N4: checkpoint as run.py:420-437 writes it ----------------------------------------------------------------
    groups = [{"params": model.density, "lr": 0.1, "skip_zero_grad": True},
              {"params": model.k0, "lr": 0.1, "skip_zero_grad": True},
              {"params": model.rgbnet.parameters(), "lr": 1e-3, "skip_zero_grad": False}]
    opt = ref_adam.MaskedAdam(groups)
    g = torch.Generator().manual_seed(3)
    n = 64
    o = torch.randn(n, 3, generator=g); o = o / o.norm(dim=-1, keepdim=True) * 2.6
    d = (torch.rand(n, 3, generator=g) - 0.5) * 1.2 - o
    d = (d / d.norm(dim=-1, keepdim=True)).contiguous()
    tgt = torch.rand(n, 3, generator=g)
    ret = model(o.contiguous(), d, d, global_step=0, **rk)
    opt.zero_grad(set_to_none=True)
    F.mse_loss(ret["rgb_marched"], tgt).backward()
    opt.step()
    ret2 = model(o.contiguous(), d, d, global_step=1, **rk)
    save.update(ck_o=o.numpy(), ck_d=d.numpy(), ck_rgb=ret2["rgb_marched"].detach().numpy().copy())
    torch.save({"global_step": 1, "model_kwargs": model.get_kwargs(), "model_state_dict": model.state_dict(),
                "optimizer_state_dict": opt.state_dict()}, os.path.join(OUT, "ref_ckpt_fine_last.tar"))

    # ---- N3: scale_volume_grid (progressive growing; replaces grids and mask) -----------------------------------------
    save.update(prescale_density=model.density.detach().numpy().copy(), prescale_k0=model.k0.detach().numpy().copy())
    model.scale_volume_grid(26 ** 3)
    save.update(scaled_density=model.density.detach().numpy().copy(), scaled_k0=model.k0.detach().numpy().copy(),
                scaled_mask=model.mask_cache.mask.numpy().copy(), scaled_world_size=model.world_size.numpy().copy())
    np.savez_compressed(os.path.join(OUT, "refpy_prep.npz"), **save)
    print("refpy_prep.npz: hit ratio", float(save["hit0"].mean()), "kept rays", save["mc_imsz"],
          "count max", float(save["count_views"].max()), "refreshed", int(save["mask_refreshed"].sum()),
          "scaled", save["scaled_world_size"])


if __name__ == "__main__":
    main()
