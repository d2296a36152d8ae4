"""Ray generation and training-ray preparation on the B200 kernels -- rows N1 / N2 of the scope table
(SURVEY.md 8f).  Same function names, arguments and return values as the reference's lib/ray_utils.py, so
run.py:146-200 (`gather_training_rays`) and run.py:82-88 (per-view render rays) run unchanged on top of it.

What differs underneath:
  * `get_rays_of_a_view` is ONE kernel (`dvgo_rays_of_view`) instead of ~25 elementwise torch launches and a
    meshgrid; the arithmetic is the reference's, rounded operation by operation (lib/ray_utils.py:9-85);
  * `get_training_rays_in_maskcache_sampling` never materialises a view's [H,W,3] ray tensors: one kernel
    generates each pixel's ray on the fly and marches it against the occupancy grid (warp per pixel, early
    exit at the first occupied sample), a prefix sum orders the survivors, a second kernel regenerates and
    writes only the surviving rays.  The reference runs its sampler + lookup over all rays in 64-row chunks
    with a host sync per chunk (lib/ray_utils.py:160-165); here there is one sync per training set.

There is no CPU path: tensors are created on the CUDA device (the reference creates them on `c2w.device`).
"""
import time

import numpy as np
import torch

from . import _C

_ext = _C.ext
_MODES = {"lefttop": 0, "center": 1}


def _device_of(t):
    return t.device if (torch.is_tensor(t) and t.is_cuda) else torch.device("cuda", torch.cuda.current_device())


def make_view(H, W, K, c2w, ndc=False, inverse_y=False, flip_x=False, flip_y=False, mode="center"):
    """Host description of one pinhole view for the kernels (`dvgo_view_t`)."""
    if mode not in _MODES:
        raise NotImplementedError(mode)
    c = np.asarray(c2w.detach().cpu().numpy() if torch.is_tensor(c2w) else c2w, dtype=np.float64)
    focal = K[0][0]
    # the reference evaluates these two coefficients in Python / numpy arithmetic (lib/ray_utils.py:68-69)
    sx = -1. / (W / (2. * focal))
    sy = -1. / (H / (2. * focal))
    return _ext.View(int(H), int(W), float(K[0][0]), float(K[1][1]), float(K[0][2]), float(K[1][2]),
                     c[:3, :4].reshape(-1).tolist(), bool(inverse_y), bool(flip_x), bool(flip_y),
                     _MODES[mode], bool(ndc), float(sx), float(sy))


def get_rays(H, W, K, c2w, inverse_y, flip_x, flip_y, mode="center"):
    """lib/ray_utils.py:9-47.  'random' mode (a per-pixel torch.rand jitter) is composed on the device with
    torch ops; 'lefttop' and 'center' are the kernel."""
    dev = _device_of(c2w)
    if mode == "random":
        i, j = torch.meshgrid(torch.linspace(0, W - 1, W, device=dev), torch.linspace(0, H - 1, H, device=dev),
                              indexing="ij")
        i, j = i.t().float(), j.t().float()
        i, j = i + torch.rand_like(i), j + torch.rand_like(j)
        if flip_x:
            i = i.flip((1,))
        if flip_y:
            j = j.flip((0,))
        c = torch.as_tensor(c2w, dtype=torch.float32, device=dev)
        if inverse_y:
            dirs = torch.stack([(i - K[0][2]) / K[0][0], (j - K[1][2]) / K[1][1], torch.ones_like(i)], -1)
        else:
            dirs = torch.stack([(i - K[0][2]) / K[0][0], -(j - K[1][2]) / K[1][1], -torch.ones_like(i)], -1)
        rays_d = torch.sum(dirs[..., None, :] * c[:3, :3], -1)
        return c[:3, 3].expand(rays_d.shape), rays_d
    view = make_view(H, W, K, c2w, False, inverse_y, flip_x, flip_y, mode)
    rays_o = torch.empty(H, W, 3, device=dev)
    rays_d = torch.empty(H, W, 3, device=dev)
    _ext.rays_of_view(view, 0, H * W, rays_o, rays_d, None)
    return rays_o, rays_d


def ndc_rays(H, W, focal, near, rays_o, rays_d):
    """lib/ray_utils.py:62-79 (torch ops; the kernel path applies the same transform inside get_rays_of_a_view)."""
    t = -(near + rays_o[..., 2]) / rays_d[..., 2]
    rays_o = rays_o + t[..., None] * rays_d
    o0 = -1. / (W / (2. * focal)) * rays_o[..., 0] / rays_o[..., 2]
    o1 = -1. / (H / (2. * focal)) * rays_o[..., 1] / rays_o[..., 2]
    o2 = 1. + 2. * near / rays_o[..., 2]
    d0 = -1. / (W / (2. * focal)) * (rays_d[..., 0] / rays_d[..., 2] - rays_o[..., 0] / rays_o[..., 2])
    d1 = -1. / (H / (2. * focal)) * (rays_d[..., 1] / rays_d[..., 2] - rays_o[..., 1] / rays_o[..., 2])
    d2 = -2. * near / rays_o[..., 2]
    return torch.stack([o0, o1, o2], -1), torch.stack([d0, d1, d2], -1)


def get_rays_of_a_view(H, W, K, c2w, ndc, inverse_y, flip_x, flip_y, mode="center"):
    """lib/ray_utils.py:80-85: rays_o, rays_d (NDC-warped when ndc), viewdirs (unit, pre-NDC), each [H,W,3]."""
    if mode == "random":
        rays_o, rays_d = get_rays(H, W, K, c2w, inverse_y, flip_x, flip_y, mode)
        viewdirs = rays_d / rays_d.norm(dim=-1, keepdim=True)
        if ndc:
            rays_o, rays_d = ndc_rays(H, W, K[0][0], 1., rays_o, rays_d)
        return rays_o, rays_d, viewdirs
    dev = _device_of(c2w)
    view = make_view(H, W, K, c2w, ndc, inverse_y, flip_x, flip_y, mode)
    out = [torch.empty(H, W, 3, device=dev) for _ in range(3)]
    _ext.rays_of_view(view, 0, H * W, *out)
    return tuple(out)


@torch.no_grad()
def get_training_rays(rgb_tr, train_poses, HW, Ks, ndc, inverse_y, flip_x, flip_y):
    """lib/ray_utils.py:88-112: every view has the same H, W, K; returns [N,H,W,3] tensors."""
    assert len(np.unique(HW, axis=0)) == 1
    assert len(np.unique(Ks.reshape(len(Ks), -1), axis=0)) == 1
    assert len(rgb_tr) == len(train_poses) and len(rgb_tr) == len(Ks) and len(rgb_tr) == len(HW)
    H, W = HW[0]
    K = Ks[0]
    dev = _device_of(rgb_tr)
    rays_o_tr = torch.empty([len(rgb_tr), H, W, 3], device=dev)
    rays_d_tr = torch.empty([len(rgb_tr), H, W, 3], device=dev)
    viewdirs_tr = torch.empty([len(rgb_tr), H, W, 3], device=dev)
    for i, c2w in enumerate(train_poses):
        view = make_view(H, W, K, c2w, ndc, inverse_y, flip_x, flip_y)
        _ext.rays_of_view(view, 0, H * W, rays_o_tr[i], rays_d_tr[i], viewdirs_tr[i])
    imsz = [1] * len(rgb_tr)
    return rgb_tr, rays_o_tr, rays_d_tr, viewdirs_tr, imsz


@torch.no_grad()
def get_training_rays_flatten(rgb_tr_ori, train_poses, HW, Ks, ndc, inverse_y, flip_x, flip_y):
    """lib/ray_utils.py:115-143: views of different sizes, flattened to [N,3]."""
    assert len(rgb_tr_ori) == len(train_poses) and len(rgb_tr_ori) == len(Ks) and len(rgb_tr_ori) == len(HW)
    dev = _device_of(rgb_tr_ori[0])
    N = sum(im.shape[0] * im.shape[1] for im in rgb_tr_ori)
    rgb_tr = torch.empty([N, 3], device=dev)
    rays_o_tr, rays_d_tr, viewdirs_tr = (torch.empty_like(rgb_tr) for _ in range(3))
    imsz = []
    top = 0
    for c2w, img, (H, W), K in zip(train_poses, rgb_tr_ori, HW, Ks):
        assert img.shape[:2] == (H, W)
        n = H * W
        view = make_view(H, W, K, c2w, ndc, inverse_y, flip_x, flip_y)
        _ext.rays_of_view(view, 0, n, rays_o_tr[top:top + n], rays_d_tr[top:top + n], viewdirs_tr[top:top + n])
        rgb_tr[top:top + n].copy_(img.flatten(0, 1), non_blocking=True)
        imsz.append(n)
        top += n
    assert top == N
    return rgb_tr, rays_o_tr, rays_d_tr, viewdirs_tr, imsz


@torch.no_grad()
def get_training_rays_in_maskcache_sampling(rgb_tr_ori, train_poses, HW, Ks, ndc, inverse_y, flip_x, flip_y,
                                            model, render_kwargs):
    """lib/ray_utils.py:146-183: keep only the rays that touch occupied space (model.hit_coarse_geo).
    Per view: hit mask (rays generated on the fly) -> prefix sum -> gather of the survivors; one host sync at
    the end for the total count.  `imsz` holds one 0-dim tensor per view, as in the reference (`n = mask.sum()`)."""
    assert len(rgb_tr_ori) == len(train_poses) and len(rgb_tr_ori) == len(Ks) and len(rgb_tr_ori) == len(HW)
    t0 = time.time()
    dev = model.density.device
    scene = model.coarse_geo_scene(**render_kwargs)
    N = sum(im.shape[0] * im.shape[1] for im in rgb_tr_ori)
    rgb_tr = torch.empty([N, 3], device=dev)
    rays_o_tr, rays_d_tr, viewdirs_tr = (torch.empty_like(rgb_tr) for _ in range(3))
    top = torch.zeros(1, dtype=torch.int64, device=dev)
    imsz = []
    for c2w, img, (H, W), K in zip(train_poses, rgb_tr_ori, HW, Ks):
        assert img.shape[:2] == (H, W)
        view = make_view(H, W, K, c2w, ndc, inverse_y, flip_x, flip_y)
        hit = _ext.view_hit_coarse_geo(view, scene)
        pos = hit.flatten().cumsum(0)
        img_d = img.to(dev, dtype=torch.float32, non_blocking=True).contiguous()
        _ext.view_gather_rays(view, hit, pos, top, img_d, rgb_tr, rays_o_tr, rays_d_tr, viewdirs_tr)
        top += pos[-1]
        imsz.append(pos[-1])
    total = int(top.item())
    out_dev = rgb_tr_ori[0].device
    res = tuple(x[:total].to(out_dev) for x in (rgb_tr, rays_o_tr, rays_d_tr, viewdirs_tr))
    get_training_rays_in_maskcache_sampling.last_seconds = time.time() - t0
    get_training_rays_in_maskcache_sampling.last_ratio = total / max(N, 1)
    return (*res, imsz)


def batch_indices_generator(N, BS):
    """lib/ray_utils.py:283-290: random permutation slices (numpy RNG, so run.py's seeding applies)."""
    idx, top = torch.LongTensor(np.random.permutation(N)), 0
    while True:
        if top + BS > N:
            idx, top = torch.LongTensor(np.random.permutation(N)), 0
        yield idx[top:top + BS]
        top += BS
