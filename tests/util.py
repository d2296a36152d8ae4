"""Shared helpers for the parity tests: seeded inputs and ulp-aware comparisons."""
import numpy as np
import torch


def ulp_diff(a, b):
    """Distance in fp32 units-in-the-last-place between two float32 arrays (same shape)."""
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float32)).view(np.int32).astype(np.int64)
    b = np.ascontiguousarray(np.asarray(b, dtype=np.float32)).view(np.int32).astype(np.int64)
    a = np.where(a < 0, np.int64(-2 ** 31) - a, a)
    b = np.where(b < 0, np.int64(-2 ** 31) - b, b)
    return np.abs(a - b)


def to_np(t):
    return t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t)


def rel_to_max(a, b):
    a, b = to_np(a).astype(np.float64), to_np(b).astype(np.float64)
    denom = max(np.abs(b).max(), 1e-30) if b.size else 1.0
    return float(np.abs(a - b).max() / denom) if a.size else 0.0


def make_rays(n, seed, extent=1.0, miss=4, zeros=True):
    """Rays around a box of half-extent ~`extent`: cameras on a sphere looking inwards, a few that
    miss the box, a few with an exactly-zero direction component (the d==0 -> 1e-6 branch)."""
    g = torch.Generator().manual_seed(seed)
    o = torch.randn(n, 3, generator=g)
    o = o / o.norm(dim=-1, keepdim=True) * (2.6 * extent)
    tgt = (torch.rand(n, 3, generator=g) - 0.5) * 1.6 * extent
    d = tgt - o
    d = d / d.norm(dim=-1, keepdim=True) * (0.8 + 0.4 * torch.rand(n, 1, generator=g))
    if miss and n > miss:
        d[:miss] = -d[:miss]
    if zeros and n > 8:
        d[miss, 0] = 0.0
        d[miss + 1, 1] = 0.0
        d[miss + 2, 2] = 0.0
    vd = d / d.norm(dim=-1, keepdim=True)
    return o.contiguous(), d.contiguous(), vd.contiguous(), torch.rand(n, 3, generator=g)


def sorted_ray_ids(n_rays, n_pts, seed, empty_rays=True):
    """A sorted int64 ray_id of length n_pts over n_rays rays with ragged (some empty) segments."""
    g = torch.Generator().manual_seed(seed)
    w = torch.rand(n_rays, generator=g)
    if empty_rays:
        w[torch.rand(n_rays, generator=g) < 0.2] = 0
    if w.sum() == 0:
        w[0] = 1
    ids = torch.multinomial(w, n_pts, replacement=True, generator=g)
    return ids.sort().values.contiguous()
