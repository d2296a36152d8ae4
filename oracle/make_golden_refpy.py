"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/refpy_*.npz by running the reference's OWN Python
(/root/reference/lib/dvgo.py, lib/dmpigo.py, lib/masked_adam.py: DirectVoxGO.forward, MaskCache,
Raw2Alpha, Alphas2Weights, MaskedAdam.step) in the GPU-less build container.

The reference's three CUDA extensions cannot run here, so `directvoxgo_b200.dropin.install` is
handed the CPU oracle (oracle/oracle.py) for them; everything else -- the sampling / masking
cascade, F.grid_sample, the rgbnet, the compositing, the autograd graph, the optimiser dispatch --
is the unmodified reference code on torch-CPU.  The loss is run.py:377-386 restated (run.py itself
needs mmcv/imageio, absent here).  What these fixtures pin: oracle/model_ref.py (CPU tests) and the
CUDA product's model-level path (GPU tests).  What they do NOT pin is the custom kernels'
arithmetic itself -- that is pinned against the real CUDA kernels by oracle/make_golden_gpu.py.

Run from the repo root:  python -m oracle.make_golden_refpy
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

REF = os.environ.get("DVGO_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def import_reference():
    from oracle import oracle as orc
    # install the oracle-backed modules WITHOUT importing the CUDA product package's kernels
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "_dvgo_dropin", os.path.join(os.path.dirname(__file__), "..", "directvoxgo_b200", "dropin.py"))
    dropin = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(dropin)
    dropin.install(orc.as_modules())
    sys.path.insert(0, REF)
    import lib.dvgo as ref_dvgo
    import lib.dmpigo as ref_dmpigo
    import lib.masked_adam as ref_adam
    return ref_dvgo, ref_dmpigo, ref_adam


def run_loss(ret, target, n, cfg):
    loss = cfg["weight_main"] * F.mse_loss(ret["rgb_marched"], target)  # run.py:377
    if cfg["weight_entropy_last"] > 0:  # run.py:379-382
        pout = ret["alphainv_last"].clamp(1e-6, 1 - 1e-6)
        loss = loss + cfg["weight_entropy_last"] * (-(pout * torch.log(pout) + (1 - pout) * torch.log(1 - pout)).mean())
    if cfg["weight_rgbper"] > 0:  # run.py:383-386
        rgbper = (ret["raw_rgb"] - target[ret["ray_id"]]).pow(2).sum(-1)
        loss = loss + cfg["weight_rgbper"] * (rgbper * ret["weights"].detach()).sum() / n
    return loss


def make_rays(n, seed, extent):
    g = torch.Generator().manual_seed(seed)
    # cameras on a sphere of radius ~2.6 extents looking at the origin with some jitter, plus a
    # few rays that miss the box entirely and a few axis-parallel ones (d component == 0).
    o = torch.randn(n, 3, generator=g)
    o = o / o.norm(dim=-1, keepdim=True) * (2.6 * extent)
    tgt = (torch.rand(n, 3, generator=g) - 0.5) * 1.6 * extent
    d = tgt - o
    d = d / d.norm(dim=-1, keepdim=True) * (0.8 + 0.4 * torch.rand(n, 1, generator=g))
    d[:4] = -d[:4]                      # misses
    d[4, 0] = 0.0; d[5, 1] = 0.0        # exact zeros -> the 1e-6 substitution
    vd = d / d.norm(dim=-1, keepdim=True)
    return o.contiguous(), d.contiguous(), vd.contiguous(), torch.rand(n, 3, generator=g)


def golden_dvgo(ref_dvgo, ref_adam, name, stage):
    torch.manual_seed(777)
    np.random.seed(777)
    extent = 1.0
    lo, hi = np.array([-1.0, -0.9, -0.8], np.float32), np.array([1.0, 0.9, 0.8], np.float32)
    if stage == "fine":
        kw = dict(num_voxels=18 ** 3, num_voxels_base=18 ** 3, alpha_init=1e-2, fast_color_thres=1e-4,
                  rgbnet_dim=12, rgbnet_direct=True, rgbnet_depth=3, rgbnet_width=128, viewbase_pe=4)
        cfg = dict(weight_main=1.0, weight_entropy_last=1e-3, weight_rgbper=1e-2, lrate_density=0.1,
                   lrate_k0=0.1, lrate_rgbnet=1e-3, skip=["density", "k0"], tv=1e-5, tv_dense=True)
        dens_scale, dens_shift = 3.0, 3.0
    else:
        kw = dict(num_voxels=16 ** 3, num_voxels_base=16 ** 3, alpha_init=1e-6, fast_color_thres=1e-7,
                  rgbnet_dim=0)
        cfg = dict(weight_main=1.0, weight_entropy_last=1e-2, weight_rgbper=0.1, lrate_density=0.1,
                   lrate_k0=0.1, lrate_rgbnet=0.0, skip=[], tv=0.0, tv_dense=False)
        dens_scale, dens_shift = 4.0, 12.0
    model = ref_dvgo.DirectVoxGO(xyz_min=lo, xyz_max=hi, **kw)
    with torch.no_grad():
        # big, shifted densities so that all four masks (bbox, occupancy, alpha, weight) and the
        # T<1e-3 early stop are exercised on a tiny grid
        model.density.copy_(torch.randn(model.density.shape) * dens_scale + dens_shift)
        model.k0.copy_(torch.randn(model.k0.shape))
        mask = torch.rand(model.mask_cache.mask.shape) > 0.25
        model.mask_cache.mask.copy_(mask)
    rays_o, rays_d, viewdirs, target = make_rays(96, 123, extent)
    rk = dict(near=0.2, far=6.0, bg=1.0, stepsize=0.5, render_depth=True)

    save = {"xyz_min": lo, "xyz_max": hi, "density0": model.density.detach().numpy().copy(),
            "k00": model.k0.detach().numpy().copy(), "mask": model.mask_cache.mask.numpy().copy(),
            "rays_o": rays_o.numpy(), "rays_d": rays_d.numpy(), "viewdirs": viewdirs.numpy(),
            "target": target.numpy(), "act_shift": np.float64(model.act_shift),
            "voxel_size": model.voxel_size.numpy(), "voxel_size_ratio": model.voxel_size_ratio.numpy(),
            "world_size": model.world_size.numpy()}
    if model.rgbnet is not None:
        lin = [m for m in model.rgbnet.modules() if isinstance(m, torch.nn.Linear)]
        for i, l in enumerate(lin):
            save["rgbnet_w%d" % i] = l.weight.detach().numpy().copy()
            save["rgbnet_b%d" % i] = l.bias.detach().numpy().copy()

    # optimiser groups as lib/utils.py:20-48 builds them
    groups = [{"params": model.density, "lr": cfg["lrate_density"], "skip_zero_grad": "density" in cfg["skip"]},
              {"params": model.k0, "lr": cfg["lrate_k0"], "skip_zero_grad": "k0" in cfg["skip"]}]
    if model.rgbnet is not None:
        groups.append({"params": model.rgbnet.parameters(), "lr": cfg["lrate_rgbnet"], "skip_zero_grad": False})
    opt = ref_adam.MaskedAdam(groups)

    for it in range(2):  # two iterations so that Adam state / step>1 is exercised
        ret = model(rays_o, rays_d, viewdirs, global_step=it, **rk)
        opt.zero_grad(set_to_none=True)
        loss = run_loss(ret, target, len(rays_o), cfg)
        loss.backward()
        if it == 0:
            for k in ("alphainv_last", "weights", "rgb_marched", "raw_alpha", "raw_rgb", "ray_id", "depth"):
                save["out_" + k] = ret[k].detach().numpy().copy()
            save["loss0"] = np.float32(loss.item())
            save["grad_density0"] = model.density.grad.numpy().copy()
            save["grad_k00"] = model.k0.grad.numpy().copy()
            if model.rgbnet is not None:
                for i, l in enumerate(lin):
                    save["grad_rgbnet_w%d" % i] = l.weight.grad.numpy().copy()
                    save["grad_rgbnet_b%d" % i] = l.bias.grad.numpy().copy()
        if cfg["tv"] > 0:  # run.py:389-395
            model.density_total_variation_add_grad(cfg["tv"] / len(rays_o), cfg["tv_dense"])
            model.k0_total_variation_add_grad(cfg["tv"] / len(rays_o), cfg["tv_dense"])
        opt.step()
        save["loss%d" % it] = np.float32(loss.item())
    save["density2"] = model.density.detach().numpy().copy()
    save["k02"] = model.k0.detach().numpy().copy()
    if model.rgbnet is not None:
        for i, l in enumerate(lin):
            save["rgbnet_w%d_2" % i] = l.weight.detach().numpy().copy()
    save["cfg_json"] = np.array(repr({**cfg, **{k: v for k, v in kw.items()}, **rk}))
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name), **save)
    print(name, "M4 =", len(save["out_ray_id"]), "loss", save["loss0"], save["loss1"],
          "i.e. %d rays, %d surviving samples" % (len(rays_o), len(save["out_weights"])))


def golden_dmpigo(ref_dmpigo, name):
    torch.manual_seed(777)
    lo, hi = np.array([-1.2, -1.1, -1.0], np.float32), np.array([1.2, 1.1, 1.0], np.float32)
    model = ref_dmpigo.DirectMPIGO(xyz_min=lo, xyz_max=hi, num_voxels=20 * 18 * 16, mpi_depth=16,
                                   fast_color_thres=1e-3, rgbnet_dim=9, rgbnet_depth=3, rgbnet_width=64,
                                   viewbase_pe=0)
    with torch.no_grad():
        model.density.add_(torch.randn(model.density.shape) * 2.0)
        model.k0.copy_(torch.randn(model.k0.shape))
    g = torch.Generator().manual_seed(5)
    n = 64
    rays_o = torch.cat([(torch.rand(n, 2, generator=g) - 0.5) * 2.6, -torch.ones(n, 1)], -1).contiguous()
    rays_d = torch.cat([(torch.rand(n, 2, generator=g) - 0.5) * 0.8, 2 * torch.ones(n, 1)], -1).contiguous()
    viewdirs = (rays_d / rays_d.norm(dim=-1, keepdim=True)).contiguous()
    rk = dict(near=0, far=1, bg=0.0, stepsize=0.5, render_depth=True)
    ret = model(rays_o, rays_d, viewdirs, global_step=0, **rk)
    loss = F.mse_loss(ret["rgb_marched"], torch.full((n, 3), 0.5))
    loss.backward()
    lin = [m for m in model.rgbnet.modules() if isinstance(m, torch.nn.Linear)]
    save = {"xyz_min": lo, "xyz_max": hi, "density0": model.density.detach().numpy().copy(),
            "k00": model.k0.detach().numpy().copy(), "rays_o": rays_o.numpy(), "rays_d": rays_d.numpy(),
            "viewdirs": viewdirs.numpy(), "world_size": model.world_size.numpy(),
            "grad_density0": model.density.grad.numpy().copy(), "grad_k00": model.k0.grad.numpy().copy(),
            "loss0": np.float32(loss.item())}
    for i, l in enumerate(lin):
        save["rgbnet_w%d" % i] = l.weight.detach().numpy().copy()
        save["rgbnet_b%d" % i] = l.bias.detach().numpy().copy()
    for k in ("alphainv_last", "weights", "rgb_marched", "raw_alpha", "raw_rgb", "ray_id", "depth"):
        save["out_" + k] = ret[k].detach().numpy().copy()
    np.savez_compressed(os.path.join(OUT, name), **save)
    print(name, "M4 =", len(save["out_ray_id"]), "loss", save["loss0"])


if __name__ == "__main__":
    ref_dvgo, ref_dmpigo, ref_adam = import_reference()
    golden_dvgo(ref_dvgo, ref_adam, "refpy_fine_small.npz", "fine")
    golden_dvgo(ref_dvgo, ref_adam, "refpy_coarse_small.npz", "coarse")
    golden_dmpigo(ref_dmpigo, "refpy_dmpigo_small.npz")
