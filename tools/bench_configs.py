"""Time the fused training step on the BASELINE.json configurations that are not the bench line (cfg 1 coarse,
cfg 4 DirectMPIGO / LLFF shape, cfg 5 320^3 with 65 536 rays) on one B200, and check each against the op-by-op
module path (drop-in ops + torch autograd) on the same inputs.  Usage: PYTHONPATH=. python tools/bench_configs.py"""
import json

import numpy as np
import torch

from directvoxgo_b200 import synthetic as syn
from directvoxgo_b200 import ray_utils as ru
from directvoxgo_b200.dmpigo import DirectMPIGO
from directvoxgo_b200.dvgo import DirectVoxGO
from directvoxgo_b200.fused import FusedTrainer
from directvoxgo_b200.trainer import ModuleTrainer

DEV = torch.device("cuda", 0)


def time_steps(trainer, batches, warm, steps):
    for i in range(warm):
        trainer.step(*batches[i % len(batches)])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        loss = trainer.step(*batches[i % len(batches)])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, float(loss)


def run(name, make_model, cfg, rk, batches, steps=50, warm=100, check=True):
    n = batches[0][0].shape[0]
    out = {"config": name, "rays_per_step": n}
    if check:   # first-step loss: fused (tensor-core rgbnet) vs module path on identical state
        torch.manual_seed(1)
        l_f = float(FusedTrainer(make_model(), dict(cfg, lrate_decay=1e9), rk).step(*batches[0]))
        torch.manual_seed(1)
        l_m = float(ModuleTrainer(make_model(), dict(cfg, lrate_decay=1e9), rk).step(*batches[0]))
        out.update(loss_fused=l_f, loss_module=l_m, loss_rel_diff=abs(l_f - l_m) / abs(l_m))
    torch.manual_seed(1)
    tr = FusedTrainer(make_model(), cfg, rk)
    ms, loss = time_steps(tr, batches, warm, steps)
    out.update(grid=[tr.X, tr.Y, tr.Z], k0_channels=tr.C, rgbnet=tr.mlp_mode if tr.model.rgbnet is not None else "none",
               ms_per_step=ms, rays_per_s=n / ms * 1e3, last_loss=loss)
    ws = tr._workspace(n, True)
    out["survivors_last_step"] = int(ws.counters[0].item())
    del tr
    torch.cuda.empty_cache()
    print(json.dumps(out), flush=True)
    return out


def main():
    res = []
    # cfg 1: coarse stage, 107x107x88, 3-channel colour grid, no rgbnet, per-lr Adam off (count table not built here)
    lo, hi = syn.coarse_bbox()

    def coarse():
        m = DirectVoxGO(lo, hi, **syn.COARSE_MODEL).to(DEV)
        return syn.randomize_grids_(m, 7)
    cfg1 = dict(N_rand=8192, lrate_density=0.1, lrate_k0=0.1, lrate_decay=20, weight_main=1.0, weight_entropy_last=0.01,
                weight_rgbper=0.1, skip_zero_grad_fields=[])
    b1 = [syn.random_training_rays(8192, n_views=50, seed=s, device=DEV) for s in range(4)]
    res.append(run("cfg1 coarse 107x107x88, C=3, 8192 rays, fwd+bwd+Adam (no TV)", coarse, cfg1, dict(syn.RENDER_KWARGS), b1))

    # cfg 4: DirectMPIGO, LLFF shape: NDC rays of 1008x756 forward-facing views, 256^3 voxels, mpi_depth 128,
    # rgbnet_dim 9, width 64 (configs/llff/llff_default.py), 4096 rays/iter, TV on (anisotropic weights)
    H, W, focal = 756, 1008, 815.0
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]], dtype=np.float32)
    mlo, mhi = np.array([-1.5, -1.3, -1.0], np.float32), np.array([1.5, 1.3, 1.0], np.float32)

    def mpi():
        m = DirectMPIGO(xyz_min=mlo, xyz_max=mhi, num_voxels=256 ** 3, mpi_depth=128, fast_color_thres=1e-3,
                        rgbnet_dim=9, rgbnet_depth=3, rgbnet_width=64, viewbase_pe=0).to(DEV)
        g = torch.Generator().manual_seed(3)
        with torch.no_grad():
            m.density.add_(torch.randn(m.density.shape, generator=g).to(DEV))
            m.k0.copy_(torch.randn(m.k0.shape, generator=g).to(DEV))
        return m
    rng = np.random.RandomState(5)
    b4 = []
    for s in range(4):
        c2w = np.eye(4, dtype=np.float32)
        c2w[:3, 3] = rng.uniform(-0.3, 0.3, 3) * np.array([1, 1, 0.2])
        ro, rd, vd = ru.get_rays_of_a_view(H, W, K, c2w, True, False, False, False)
        idx = torch.tensor(rng.choice(H * W, 4096, replace=False), device=DEV)
        b4.append(tuple(x.reshape(-1, 3)[idx].contiguous() for x in (ro, rd, vd)) + (torch.rand(4096, 3, device=DEV),))
    cfg4 = dict(N_rand=4096, lrate_density=0.1, lrate_k0=0.1, lrate_rgbnet=1e-3, lrate_decay=20, weight_main=1.0,
                weight_entropy_last=0.01, weight_rgbper=0.01, weight_tv_density=1e-5, weight_tv_k0=1e-5, tv_dense=True,
                skip_zero_grad_fields=[])
    res.append(run("cfg4 DirectMPIGO LLFF-shape, 256^3 voxels / mpi_depth 128, C=9, rgbnet 64, 4096 NDC rays, fwd+bwd+TV+Adam",
                   mpi, cfg4, dict(near=0, far=1, bg=1.0, stepsize=1.0), b4))

    # cfg 5: 320^3 density + 12-ch k0 + rgbnet 128, 65 536 rays per iteration on ONE GPU (the ray-sharded runs split this)
    flo, fhi = syn.fine_bbox()

    def big():
        m = DirectVoxGO(flo, fhi, **dict(syn.FINE_MODEL, num_voxels=320 ** 3, num_voxels_base=320 ** 3)).to(DEV)
        return syn.randomize_grids_(m, 9)
    b5 = [syn.random_training_rays(65536, n_views=100, seed=s, device=DEV) for s in range(2)]
    res.append(run("cfg5 320^3, C=12, rgbnet 128, 65536 rays on one GPU, fwd+bwd+TV+MaskedAdam", big, dict(syn.FINE_TRAIN),
                   dict(syn.RENDER_KWARGS), b5, steps=10, warm=5))
    print(json.dumps({"results": res}))


if __name__ == "__main__":
    main()
