"""Measure the rows either side of the per-iteration path on one B200, BASELINE shapes:
800x800 Blender-geometry views, 160^3 fine grid with a sphere occupancy mask (coarse 107x107x88 grid for
voxel_count_views).  Each fused kernel path is timed beside the op-by-op composition of the reference's
algorithm on the same GPU (our drop-in ops + torch, i.e. what lib/ray_utils.py / lib/dvgo.py do).
Usage: PYTHONPATH=. python tools/bench_prep.py [n_views] > gpurun_out/prep_bench.json"""
import json
import sys
import time

import numpy as np
import torch

from directvoxgo_b200 import ext, render_utils_cuda, synthetic as syn
from directvoxgo_b200 import ray_utils as ru
from directvoxgo_b200.dvgo import DirectVoxGO
from directvoxgo_b200.ops import grid_sample_trilinear

DEV = torch.device("cuda", 0)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.time() - t0) / reps, out


def torch_rays(H, W, K, c2w):
    """lib/ray_utils.py:9-47,80-85 as torch ops on the GPU."""
    c = torch.as_tensor(c2w, dtype=torch.float32, device=DEV)
    i, j = torch.meshgrid(torch.linspace(0, W - 1, W, device=DEV), torch.linspace(0, H - 1, H, device=DEV), indexing="ij")
    i, j = i.t() + 0.5, j.t() + 0.5
    dirs = torch.stack([(i - K[0][2]) / K[0][0], -(j - K[1][2]) / K[1][1], -torch.ones_like(i)], -1)
    rays_d = torch.sum(dirs[..., None, :] * c[:3, :3], -1)
    rays_o = c[:3, 3].expand(rays_d.shape)
    return rays_o, rays_d, rays_d / rays_d.norm(dim=-1, keepdim=True)


def main():
    n_views = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    H = W = 800
    K = syn.intrinsics(H, W)
    poses = syn.random_poses(n_views, seed=11)
    rk = dict(near=2.0, far=6.0, stepsize=0.5)
    lo, hi = syn.fine_bbox()
    torch.manual_seed(0)
    model = DirectVoxGO(lo, hi, **dict(syn.FINE_MODEL, num_voxels=160 ** 3, num_voxels_base=160 ** 3)).to(DEV)
    ax = torch.linspace(-1, 1, 160, device=DEV)
    g = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), -1)
    model.mask_cache.mask.copy_(g.norm(dim=-1) < 0.55)
    imgs = [torch.rand(H, W, 3, device=DEV) for _ in range(n_views)]
    HW, Ks = np.array([[H, W]] * n_views), np.stack([K] * n_views)
    out = {"views": n_views, "resolution": "800x800", "grid": "160^3, sphere occupancy"}

    # N2 ray generation
    t_k, _ = timed(lambda: ru.get_rays_of_a_view(H, W, K, poses[0], False, False, False, False), 20)
    t_t, _ = timed(lambda: torch_rays(H, W, K, poses[0]), 20)
    out["rays_of_view_ms"] = {"kernel": t_k * 1e3, "torch_ops": t_t * 1e3, "bytes_out": 3 * H * W * 12,
                              "GBps_kernel": 3 * H * W * 12 / t_k / 1e9}

    # N1 training-ray preparation
    def fused():
        return ru.get_training_rays_in_maskcache_sampling(imgs, poses, HW, Ks, False, False, False, False, model, rk)

    def composed():  # lib/ray_utils.py:146-183 on the drop-in ops (64-row chunks, boolean compaction per view)
        stepdist = float(rk["stepsize"] * model.voxel_size)
        outs, top = [], 0
        for c2w, img in zip(poses, imgs):
            ro, rd, vd = torch_rays(H, W, K, c2w)
            mask = torch.empty(H, W, dtype=torch.bool, device=DEV)
            for i in range(0, H, 64):
                o = ro[i:i + 64].reshape(-1, 3).contiguous()
                d = rd[i:i + 64].reshape(-1, 3).contiguous()
                pts, outside, ray_id = render_utils_cuda.sample_pts_on_rays(o, d, model.xyz_min, model.xyz_max,
                                                                           rk["near"], rk["far"], stepdist)[:3]
                keep = ~outside
                hit = torch.zeros(len(o), dtype=torch.bool, device=DEV)
                hit[ray_id[keep][model.mask_cache(pts[keep])]] = 1
                mask[i:i + 64] = hit.reshape(-1, W)
            outs.append((img[mask], ro[mask], rd[mask], vd[mask]))
            top += int(mask.sum())
        return outs, top

    t_f, res = timed(fused, 2)
    t_c, ref = timed(composed, 1)
    kept = sum(int(n) for n in res[4])
    assert abs(kept - ref[1]) <= 1e-3 * ref[1]      # torch_rays differs from the kernel's rays in the last bit
    out["training_ray_prep"] = {"fused_ms_per_view": t_f / n_views * 1e3, "composed_ms_per_view": t_c / n_views * 1e3,
                                "rays_per_s_fused": n_views * H * W / t_f, "kept_ratio": ref[1] / (n_views * H * W), "kept_fused": kept, "kept_composed": ref[1],
                                "speedup": t_c / t_f}

    # N3 voxel_count_views on the coarse grid (where the reference calls it, run.py:268-276)
    clo, chi = syn.coarse_bbox()
    coarse = DirectVoxGO(clo, chi, num_voxels=1024000, num_voxels_base=1024000, alpha_init=1e-6, fast_color_thres=1e-7).to(DEV)
    _, ro_all, rd_all, _, imsz = ru.get_training_rays(torch.stack(imgs), poses, HW, Ks, False, False, False, False)

    def count_autograd():  # lib/dvgo.py:265-295 on our autograd trilinear op
        n_samples = int(np.linalg.norm(np.array(coarse.density.shape[2:]) + 1) / rk["stepsize"]) + 1
        rng = torch.arange(n_samples, device=DEV)[None].float()
        count = torch.zeros_like(coarse.density.detach())
        for ro_v, rd_v in zip(ro_all.split(imsz), rd_all.split(imsz)):
            ones = torch.ones_like(coarse.density).requires_grad_()
            for ro, rd in zip(ro_v.flatten(0, -2).split(10000), rd_v.flatten(0, -2).split(10000)):
                vec = torch.where(rd == 0, torch.full_like(rd, 1e-6), rd)
                ra, rb = (coarse.xyz_max - ro) / vec, (coarse.xyz_min - ro) / vec
                t_min = torch.minimum(ra, rb).amax(-1).clamp(min=rk["near"], max=rk["far"])
                step = rk["stepsize"] * coarse.voxel_size * rng
                interpx = t_min[..., None] + step / rd.norm(dim=-1, keepdim=True)
                pts = ro[..., None, :] + rd[..., None, :] * interpx[..., None]
                grid_sample_trilinear(ones, pts, coarse.xyz_min, coarse.xyz_max).sum().backward()
            with torch.no_grad():
                count += (ones.grad > 1)
        return count

    t_f, c1 = timed(lambda: coarse.voxel_count_views(ro_all, rd_all, imsz, rk["near"], rk["far"], rk["stepsize"]), 2)
    t_c, c2 = timed(count_autograd, 1)
    d = (c1 - c2).abs()
    out["voxel_count_views"] = {"fused_ms_per_view": t_f / n_views * 1e3, "autograd_ms_per_view": t_c / n_views * 1e3,
                                "speedup": t_c / t_f, "identical_voxels": float((d == 0).float().mean()),
                                "max_diff": float(d.max()), "grid": list(coarse.density.shape[2:])}

    # N3 occupancy refresh and resize at 160^3
    syn.randomize_grids_(model, 3)
    t_k, _ = timed(lambda: model.update_occupancy_cache(), 10)
    import torch.nn.functional as F

    def refresh_torch():
        a = F.max_pool3d(model.activate_density(model.density), kernel_size=3, padding=1, stride=1)[0, 0]
        return model.mask_cache.mask & (a > model.fast_color_thres)
    t_t, _ = timed(refresh_torch, 10)
    out["occupancy_refresh_ms"] = {"kernel": t_k * 1e3, "torch_ops": t_t * 1e3}
    small = torch.randn(1, 12, 127, 127, 127, device=DEV)
    t_k, a = timed(lambda: ext.resize_trilinear(small, 160, 160, 160), 5)
    t_t, b = timed(lambda: F.interpolate(small, size=(160, 160, 160), mode="trilinear", align_corners=True), 5)
    out["resize_127_to_160_12ch_ms"] = {"kernel": t_k * 1e3, "aten": t_t * 1e3, "max_abs_diff": float((a - b).abs().max())}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
