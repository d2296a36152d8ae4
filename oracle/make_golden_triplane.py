"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/refpy_triplane.npz by calling the reference's OWN tri-plane
sampler, `lib/tri_dvgo.py` `DirectVoxGO.grid_sampler2D` (:456-479), on the CPU: forward for both aggregations and the
gradient w.r.t. the three planes.  The module's unrelated imports that are absent here (matplotlib, imageio) are
stubbed; its CUDA extensions are served by the CPU oracle as in make_golden_refpy.py (the sampler itself is pure
ATen).   Run from the repo root:  python -m oracle.make_golden_triplane
"""
import os
import sys
import types

import numpy as np
import torch

from oracle.make_golden_refpy import OUT, import_reference


def stub_missing_imports():
    """Packages the reference imports at module top but never touches on the hot path, absent in this image."""
    import importlib.util

    def missing(name):
        if name in sys.modules:          # already imported, or already stubbed by an earlier call
            return False
        try:
            return importlib.util.find_spec(name) is None
        except (ValueError, ImportError):
            return True
    if missing("matplotlib"):
        mp, pp = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
        pp.step = lambda *a, **k: None
        mp.pyplot = pp
        sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mp, pp
    for n in ("imageio", "mmcv", "lpips"):
        if missing(n):
            sys.modules[n] = types.ModuleType(n)


def import_tri():
    stub_missing_imports()
    import lib.tri_dvgo as tri
    return tri


def main():
    import_reference()
    tri = import_tri()

    g = torch.Generator().manual_seed(21)
    lo, hi = torch.tensor([-1.0, -0.8, -1.2]), torch.tensor([1.1, 0.9, 1.0])
    C = 6
    grids = {"xy": torch.randn(1, C, 13, 17, generator=g), "yz": torch.randn(1, C, 11, 12, generator=g),
             "zx": torch.randn(1, C, 16, 9, generator=g)}
    xyz = lo + (hi - lo) * (torch.rand(1500, 3, generator=g) * 1.3 - 0.15)      # some points outside: zero padding
    go = torch.randn(1500, 3 * C, generator=g)
    save = {"xyz_min": lo.numpy(), "xyz_max": hi.numpy(), "xyz": xyz.numpy(), "grad_out": go.numpy()}
    for k, v in grids.items():
        save["plane_" + k] = v.numpy().copy()
    for agg in ("concat", "sum"):
        fake = types.SimpleNamespace(xyz_min=lo, xyz_max=hi, tri_aggregation=agg, global_cell_decode=False)
        leaf = {k: v.clone().requires_grad_() for k, v in grids.items()}
        out = tri.DirectVoxGO.grid_sampler2D(fake, xyz, leaf)
        save["out_" + agg] = out.detach().numpy().copy()
        (out * go[:, :out.shape[1]]).sum().backward()
        for k in leaf:
            save["grad_%s_%s" % (agg, k)] = leaf[k].grad.numpy().copy()
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "refpy_triplane.npz"), **save)
    print("refpy_triplane.npz", save["out_concat"].shape, save["out_sum"].shape)
    golden_render(tri)


def build_tri_model(tri, device="cpu"):
    """The tri-plane model of lib/tri_dvgo.py:36-257 at a small size (the EDSR encoder / mapping nets are built by the
    constructor but not used by `render`, which takes the three feature planes as an argument)."""
    import contextlib
    import io
    lo, hi = np.array([-1.0, -0.9, -0.8], np.float32), np.array([1.0, 0.9, 0.8], np.float32)
    torch.manual_seed(777)
    with contextlib.redirect_stdout(io.StringIO()):
        m = tri.DirectVoxGO(lo, hi, num_voxels=18 ** 3, num_voxels_base=18 ** 3, alpha_init=1e-2, fast_color_thres=1e-4,
                            rgbnet_dim=4, rgbnet_direct=True, rgbnet_depth=3, rgbnet_width=128, viewbase_pe=4,
                            n_resblocks=1)
    return m.to(device)


def golden_render(tri):
    """`DirectVoxGO.render(feats, rays_o, rays_d, viewdirs, ...)` of lib/tri_dvgo.py:688-809 on the CPU (custom ops
    from the C oracle, everything else the unmodified reference): outputs + gradients w.r.t. the density grid, the
    three feature planes and the rgbnet, for tests/test_gpu_0_vs_ref.py (the same call on the B200 kernels)."""
    from oracle.make_golden_refpy import make_rays
    m = build_tri_model(tri)
    g = torch.Generator().manual_seed(31)
    with torch.no_grad():
        m.density.copy_(torch.randn(m.density.shape, generator=g) * 3.0 + 3.0)
        m.mask_cache.mask.copy_(torch.rand(m.mask_cache.mask.shape, generator=g) > 0.25)
    feats = {"xy": torch.randn(1, 4, 14, 19, generator=g), "yz": torch.randn(1, 4, 12, 13, generator=g),
             "zx": torch.randn(1, 4, 17, 11, generator=g)}
    rays_o, rays_d, viewdirs, target = make_rays(80, 321, 1.0)
    rk = dict(near=0.2, far=6.0, bg=1.0, stepsize=0.5, render_depth=True)
    leaf = {k: v.clone().requires_grad_() for k, v in feats.items()}
    ret = m.render(leaf, rays_o, rays_d, viewdirs, global_step=0, **rk)
    loss = torch.nn.functional.mse_loss(ret["rgb_marched"], target) + 1e-2 * ret["alphainv_last"].mean()
    loss.backward()
    lin = [x for x in m.rgbnet.modules() if isinstance(x, torch.nn.Linear)]
    save = {"density0": m.density.detach().numpy().copy(), "mask": m.mask_cache.mask.numpy().copy(),
            "rays_o": rays_o.numpy(), "rays_d": rays_d.numpy(), "viewdirs": viewdirs.numpy(), "target": target.numpy(),
            "loss": np.float32(loss.item()), "grad_density": m.density.grad.numpy().copy()}
    for k in feats:
        save["plane_" + k] = feats[k].numpy().copy()
        save["grad_plane_" + k] = leaf[k].grad.numpy().copy()
    for i, l in enumerate(lin):
        save["rgbnet_w%d" % i] = l.weight.detach().numpy().copy()
        save["rgbnet_b%d" % i] = l.bias.detach().numpy().copy()
        save["grad_rgbnet_w%d" % i] = l.weight.grad.numpy().copy()
    for k in ("alphainv_last", "weights", "rgb_marched", "raw_alpha", "raw_rgb", "ray_id", "depth"):
        save["out_" + k] = ret[k].detach().numpy().copy()
    np.savez_compressed(os.path.join(OUT, "refpy_triplane_render.npz"), **save)
    print("refpy_triplane_render.npz: M4 =", len(save["out_ray_id"]), "loss", save["loss"])


if __name__ == "__main__":
    main()
