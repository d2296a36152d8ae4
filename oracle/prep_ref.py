"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the rows either side of the per-iteration path:
ray generation (lib/ray_utils.py:9-85), hit_coarse_geo (lib/dvgo.py:412-423), voxel_count_views (:265-295), the
occupancy refresh (run.py:330-332) and the trilinear resize of scale_volume_grid (lib/dvgo.py:236-241).

numpy float32 with one rounding per operation (torch's elementwise semantics; the CUDA kernels in
directvoxgo_b200/csrc/fused_prep.cu spell the same operations with round-to-nearest intrinsics) plus the C oracle
for the reference's custom ops.  Pinned by tests/golden/refpy_prep.npz -- outputs of the reference's OWN Python
(oracle/make_golden_prep.py) -- in tests/test_oracle_golden.py.  The product never imports this file.
"""
import numpy as np
import torch

from . import oracle as orc

f32 = np.float32


def rays_of_view(H, W, K, c2w, ndc=False, inverse_y=False, flip_x=False, flip_y=False, mode="center"):
    """get_rays_of_a_view (lib/ray_utils.py:80-85): rays_o, rays_d, viewdirs, each [H,W,3] float32."""
    c = np.asarray(c2w, dtype=f32)
    i, j = np.meshgrid(np.arange(W, dtype=f32), np.arange(H, dtype=f32), indexing="xy")     # i: column, j: row
    if mode == "center":
        i, j = i + f32(0.5), j + f32(0.5)
    elif mode != "lefttop":
        raise NotImplementedError(mode)
    if flip_x:
        i = i[:, ::-1]
    if flip_y:
        j = j[::-1, :]
    fx, fy, cx, cy = f32(K[0][0]), f32(K[1][1]), f32(K[0][2]), f32(K[1][2])
    a = (i - cx) / fx
    b = (j - cy) / fy
    if inverse_y:
        dirs = [a, b, np.ones_like(a)]
    else:
        dirs = [a, -b, -np.ones_like(a)]
    rays_d = np.stack([(dirs[0] * c[k, 0] + dirs[1] * c[k, 1]) + dirs[2] * c[k, 2] for k in range(3)], -1).astype(f32)
    rays_o = np.broadcast_to(c[:3, 3], rays_d.shape).astype(f32)
    nrm = np.sqrt((rays_d[..., 0] * rays_d[..., 0] + rays_d[..., 1] * rays_d[..., 1]) + rays_d[..., 2] * rays_d[..., 2])
    viewdirs = (rays_d / nrm[..., None]).astype(f32)
    if ndc:     # lib/ray_utils.py:62-79 with near = 1.
        focal = K[0][0]
        sx, sy = f32(-1. / (W / (2. * focal))), f32(-1. / (H / (2. * focal)))
        t = -(f32(1.0) + rays_o[..., 2]) / rays_d[..., 2]
        o = rays_o + t[..., None] * rays_d
        o0 = sx * o[..., 0] / o[..., 2]
        o1 = sy * o[..., 1] / o[..., 2]
        o2 = f32(1.0) + f32(2.0) / o[..., 2]
        d0 = sx * (rays_d[..., 0] / rays_d[..., 2] - o[..., 0] / o[..., 2])
        d1 = sy * (rays_d[..., 1] / rays_d[..., 2] - o[..., 1] / o[..., 2])
        d2 = f32(-2.0) / o[..., 2]
        rays_o, rays_d = np.stack([o0, o1, o2], -1).astype(f32), np.stack([d0, d1, d2], -1).astype(f32)
    return rays_o, rays_d, viewdirs


def hit_coarse_geo(rays_o, rays_d, xyz_min, xyz_max, mask, near, far, stepdist):
    """lib/dvgo.py:412-423 on the C oracle ops: which rays have a sample inside the bbox and in occupied space."""
    ro = torch.as_tensor(np.ascontiguousarray(rays_o, dtype=f32)).reshape(-1, 3)
    rd = torch.as_tensor(np.ascontiguousarray(rays_d, dtype=f32)).reshape(-1, 3)
    lo, hi = torch.as_tensor(xyz_min, dtype=torch.float32), torch.as_tensor(xyz_max, dtype=torch.float32)
    m = torch.as_tensor(mask).bool()
    scale = (torch.tensor(list(m.shape), dtype=torch.float32) - 1) / (hi - lo)
    shift = -lo * scale
    pts, outside, ray_id = orc.sample_pts_on_rays(ro, rd, lo, hi, near, far, stepdist)[:3]
    keep = ~outside
    occ = orc.maskcache_lookup(m, pts[keep].contiguous(), scale, shift)
    hit = torch.zeros(len(ro), dtype=torch.bool)
    hit[ray_id[keep][occ]] = True
    return hit.numpy().reshape(np.asarray(rays_o).shape[:-1])


def voxel_count_views(rays_o_views, rays_d_views, xyz_min, xyz_max, shape, near, far, stepsize, voxel_size):
    """lib/dvgo.py:265-295: per view, scatter the trilinear weights of every (ray, sample) and count acc > 1."""
    lo, hi = np.asarray(xyz_min, f32), np.asarray(xyz_max, f32)
    n_samples = int(np.linalg.norm(np.array(shape) + 1) / stepsize) + 1
    rng = np.arange(n_samples, dtype=f32)[None]
    stepdist = f32(f32(stepsize) * f32(voxel_size))
    count = np.zeros((1, 1, *shape), f32)
    for ro, rd in zip(rays_o_views, rays_d_views):
        ro, rd = np.asarray(ro, f32).reshape(-1, 3), np.asarray(rd, f32).reshape(-1, 3)
        vec = np.where(rd == 0, f32(1e-6), rd)
        ra, rb = (hi - ro) / vec, (lo - ro) / vec
        t_min = np.clip(np.minimum(ra, rb).max(-1), f32(near), f32(far))
        nrm = np.sqrt((rd[:, 0] * rd[:, 0] + rd[:, 1] * rd[:, 1]) + rd[:, 2] * rd[:, 2])
        interp = t_min[:, None] + (stepdist * rng) / nrm[:, None]
        pts = (ro[:, None, :] + rd[:, None, :] * interp[..., None]).astype(f32).reshape(-1, 3)
        acc = torch.zeros(1, 1, *shape)
        orc.grid_sample_3d_backward(torch.ones(len(pts), 1), torch.as_tensor(pts), torch.as_tensor(lo), torch.as_tensor(hi), acc)
        count += (acc.numpy() > 1)
    return count


def alpha_maxpool_mask(density, act_shift, interval, thres, mask_in=None):
    """run.py:330-332: mask & (maxpool3(raw2alpha(density)) > thres); density [X,Y,Z]."""
    d = torch.as_tensor(np.ascontiguousarray(density, dtype=f32))
    alpha = orc.raw2alpha(d.flatten(), float(act_shift), float(interval))[1].reshape(d.shape).numpy()
    X, Y, Z = alpha.shape
    pad = np.full((X + 2, Y + 2, Z + 2), -np.inf, f32)
    pad[1:-1, 1:-1, 1:-1] = alpha
    m = np.full(alpha.shape, -np.inf, f32)
    for a in range(3):
        for b in range(3):
            for c in range(3):
                m = np.maximum(m, pad[a:a + X, b:b + Y, c:c + Z])
    out = m > f32(thres)
    return out if mask_in is None else (out & np.asarray(mask_in, bool))


def resize_trilinear(src, size):
    """F.interpolate(src[None], size, mode='trilinear', align_corners=True)[0] for src [C,X,Y,Z] (ATen's
    upsample_trilinear3d index arithmetic; the blend is evaluated in float64 -- tolerance-level restatement)."""
    src = np.asarray(src, f32)
    C, X, Y, Z = src.shape
    X2, Y2, Z2 = size

    def axis(n_in, n_out):
        r = f32(n_in - 1) / f32(n_out - 1) if n_out > 1 else f32(0)
        f = (r * np.arange(n_out, dtype=f32)).astype(f32)
        i0 = f.astype(np.int64)
        i1 = i0 + (i0 < n_in - 1)
        l1 = (f - i0.astype(f32)).astype(np.float64)
        return i0, i1, 1.0 - l1, l1
    x0, x1, ax0, ax1 = axis(X, X2)
    y0, y1, ay0, ay1 = axis(Y, Y2)
    z0, z1, az0, az1 = axis(Z, Z2)
    s = src.astype(np.float64)

    def g(xi, yi, zi):
        return s[:, xi][:, :, yi][:, :, :, zi]
    wz0, wz1 = az0[None, None, None, :], az1[None, None, None, :]
    wy0, wy1 = ay0[None, None, :, None], ay1[None, None, :, None]
    wx0, wx1 = ax0[None, :, None, None], ax1[None, :, None, None]
    out = wx0 * (wy0 * (wz0 * g(x0, y0, z0) + wz1 * g(x0, y0, z1)) + wy1 * (wz0 * g(x0, y1, z0) + wz1 * g(x0, y1, z1))) + \
        wx1 * (wy0 * (wz0 * g(x1, y0, z0) + wz1 * g(x1, y0, z1)) + wy1 * (wz0 * g(x1, y1, z0) + wz1 * g(x1, y1, z1)))
    return out.astype(f32)
