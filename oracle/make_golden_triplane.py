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


def main():
    import_reference()
    mp, pp = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
    pp.step = lambda *a, **k: None
    mp.pyplot = pp
    sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mp, pp
    for n in ("imageio", "cv2", "mmcv"):
        sys.modules.setdefault(n, types.ModuleType(n))
    import lib.tri_dvgo as tri

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


if __name__ == "__main__":
    main()
