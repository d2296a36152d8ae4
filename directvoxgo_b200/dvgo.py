"""DirectVoxGO dense-grid scene model on the B200 kernels.

Same constructor keywords, parameter / buffer names (state_dict keys) and `forward` result
dictionary as the reference's lib/dvgo.py `DirectVoxGO` (:30-42 ctor, :450-577 forward), so a
training loop written against the reference (run.py:327-398) runs unchanged on this class.
What differs is underneath: every custom op is one of our sm_100a kernels, trilinear sampling and
segment_coo are our own ops instead of ATen / torch_scatter, and the 16 boolean-mask compactions
of the reference's forward are folded into 4 (one index build per mask stage).

The LIIF research branch (`implicit_voxel_feat=True`, lib/dvgo.py:329-410) is out of scope
 and raises NotImplementedError.
"""
import time

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _C, render_utils_cuda, total_variation_cuda
from .ops import Alphas2Weights, Raw2Alpha, grid_sample_trilinear, segment_coo


def _grid_points(xyz_min, xyz_max, shape):
    """World coordinates of every voxel centre-line node, [X,Y,Z,3] (lib/dvgo.py:143-147)."""
    axes = [torch.linspace(float(xyz_min[i]), float(xyz_max[i]), int(shape[i]), device=xyz_min.device)
            for i in range(3)]
    return torch.stack(torch.meshgrid(*axes, indexing="ij"), -1)


class MaskCache(nn.Module):
    """Occupancy grid of known free space (lib/dvgo.py:583-613).

    Built either from a coarse-stage checkpoint (`path`: density max-pooled 3x3x3, softplus alpha
    >= mask_cache_thres, :587-593) or from an explicit boolean `mask` + bbox."""

    def __init__(self, path=None, mask_cache_thres=None, mask=None, xyz_min=None, xyz_max=None):
        super().__init__()
        if path is not None:
            st = torch.load(path, map_location="cpu", weights_only=False)
            self.mask_cache_thres = mask_cache_thres
            kw = st["model_kwargs"]
            dens = F.max_pool3d(st["model_state_dict"]["density"], kernel_size=3, padding=1, stride=1)
            alpha = 1 - torch.exp(-F.softplus(dens + kw["act_shift"]) * kw["voxel_size_ratio"])
            mask = (alpha >= mask_cache_thres)[0, 0]
            xyz_min, xyz_max = kw["xyz_min"], kw["xyz_max"]
        lo = torch.as_tensor(np.asarray(xyz_min.detach().cpu() if torch.is_tensor(xyz_min) else xyz_min),
                             dtype=torch.float32)
        hi = torch.as_tensor(np.asarray(xyz_max.detach().cpu() if torch.is_tensor(xyz_max) else xyz_max),
                             dtype=torch.float32)
        mask = mask.bool()
        lo, hi = lo.to(mask.device), hi.to(mask.device)
        self.register_buffer("mask", mask.contiguous())
        scale = (torch.tensor(list(mask.shape), dtype=torch.float32, device=mask.device) - 1) / (hi - lo)
        self.register_buffer("xyz2ijk_scale", scale)
        self.register_buffer("xyz2ijk_shift", -lo * scale)

    @torch.no_grad()
    def forward(self, xyz):
        shape = xyz.shape[:-1]
        hit = render_utils_cuda.maskcache_lookup(
            self.mask, xyz.reshape(-1, 3).contiguous(), self.xyz2ijk_scale, self.xyz2ijk_shift)
        return hit.reshape(shape)


class DirectVoxGO(nn.Module):
    def __init__(self, xyz_min, xyz_max,
                 num_voxels=0, num_voxels_base=0,
                 alpha_init=None,
                 mask_cache_path=None, mask_cache_thres=1e-3,
                 fast_color_thres=0,
                 rgbnet_dim=0, rgbnet_direct=False, rgbnet_full_implicit=False,
                 rgbnet_depth=3, rgbnet_width=128,
                 viewbase_pe=4,
                 posbase_pe=0,
                 implicit_voxel_feat=False, feat_unfold=False, local_ensemble=True, cell_decode=True,
                 **kwargs):
        super().__init__()
        if implicit_voxel_feat:
            raise NotImplementedError("implicit_voxel_feat (LIIF branch) is outside the B200 hot path")
        self.register_buffer("xyz_min", torch.tensor(np.asarray(xyz_min, dtype=np.float32)))
        self.register_buffer("xyz_max", torch.tensor(np.asarray(xyz_max, dtype=np.float32)))
        self.fast_color_thres = fast_color_thres
        self.posbase_pe = posbase_pe
        self.implicit_voxel_feat = False
        self.feat_unfold, self.local_ensemble, self.cell_decode = feat_unfold, local_ensemble, cell_decode

        # base resolution fixes the meaning of one density unit (lib/dvgo.py:56-57)
        self.num_voxels_base = num_voxels_base
        self.voxel_size_base = ((self.xyz_max - self.xyz_min).prod() / num_voxels_base).pow(1 / 3)
        # density bias so that a zero grid renders alpha_init everywhere (:60-61)
        self.alpha_init = alpha_init
        self.act_shift = np.log(1 / (1 - alpha_init) - 1)
        self._set_grid_resolution(num_voxels)

        self.density = nn.Parameter(torch.zeros([1, 1, *self.world_size]))
        self.rgbnet_kwargs = dict(
            rgbnet_dim=rgbnet_dim, rgbnet_direct=rgbnet_direct, rgbnet_full_implicit=rgbnet_full_implicit,
            rgbnet_depth=rgbnet_depth, rgbnet_width=rgbnet_width, viewbase_pe=viewbase_pe,
            posbase_pe=posbase_pe, implicit_voxel_feat=False, feat_unfold=feat_unfold,
            local_ensemble=local_ensemble, cell_decode=cell_decode)
        self.rgbnet_full_implicit = rgbnet_full_implicit
        if rgbnet_dim <= 0:  # coarse stage: the grid stores colour directly (:83-87)
            self.k0_dim = 3
            self.k0 = nn.Parameter(torch.zeros([1, 3, *self.world_size]))
            self.rgbnet = None
        else:  # fine stage: feature grid + shallow view-dependent MLP (:88-133)
            self.k0_dim = 0 if rgbnet_full_implicit else rgbnet_dim
            self.k0 = nn.Parameter(torch.zeros([1, self.k0_dim, *self.world_size]))
            self.rgbnet_direct = rgbnet_direct
            self.register_buffer("viewfreq", torch.tensor([2.0 ** i for i in range(viewbase_pe)]))
            if posbase_pe > 0:
                self.register_buffer("posfreq", torch.tensor([2.0 ** i for i in range(posbase_pe)]))
            dim0 = 3 + 3 * viewbase_pe * 2
            if rgbnet_full_implicit:
                pass
            elif posbase_pe > 0:
                dim0 += 3 + 3 * posbase_pe * 2
            elif rgbnet_direct:
                dim0 += self.k0_dim
            else:
                dim0 += self.k0_dim - 3
            hidden = [nn.Sequential(nn.Linear(rgbnet_width, rgbnet_width), nn.ReLU(inplace=True))
                      for _ in range(rgbnet_depth - 2)]
            self.rgbnet = nn.Sequential(nn.Linear(dim0, rgbnet_width), nn.ReLU(inplace=True), *hidden,
                                        nn.Linear(rgbnet_width, 3))
            nn.init.constant_(self.rgbnet[-1].bias, 0)

        # occupancy of known free space (:135-153)
        self.mask_cache_path = mask_cache_path
        self.mask_cache_thres = mask_cache_thres
        if mask_cache_path:
            coarse = MaskCache(path=mask_cache_path, mask_cache_thres=mask_cache_thres).to(self.xyz_min.device)
            mask = coarse(_grid_points(self.xyz_min, self.xyz_max, self.density.shape[2:]))
        else:
            mask = torch.ones(list(self.world_size), dtype=torch.bool)
        self.mask_cache = MaskCache(mask=mask, xyz_min=self.xyz_min, xyz_max=self.xyz_max)

    # ------------------------------------------------------------------ resolution / bookkeeping
    def _set_grid_resolution(self, num_voxels):
        self.num_voxels = num_voxels
        extent = self.xyz_max - self.xyz_min
        self.voxel_size = (extent.prod() / num_voxels).pow(1 / 3)
        self.world_size = (extent / self.voxel_size).long()
        self.voxel_size_ratio = self.voxel_size / self.voxel_size_base

    def get_kwargs(self):
        return {
            "xyz_min": self.xyz_min.cpu().numpy(), "xyz_max": self.xyz_max.cpu().numpy(),
            "num_voxels": self.num_voxels, "num_voxels_base": self.num_voxels_base,
            "alpha_init": self.alpha_init, "act_shift": self.act_shift,
            "voxel_size_ratio": self.voxel_size_ratio,
            "mask_cache_path": self.mask_cache_path, "mask_cache_thres": self.mask_cache_thres,
            "fast_color_thres": self.fast_color_thres,
            **self.rgbnet_kwargs,
        }

    @torch.no_grad()
    def maskout_near_cam_vox(self, cam_o, near):
        pts = _grid_points(self.xyz_min, self.xyz_max, self.density.shape[2:])
        nearest = torch.stack([(pts.unsqueeze(-2) - co).pow(2).sum(-1).sqrt().amin(-1)
                               for co in cam_o.split(100)]).amin(0)
        self.density[nearest[None, None] <= near] = -100

    @torch.no_grad()
    def scale_volume_grid(self, num_voxels):
        """Progressive growing (lib/dvgo.py:228-263): trilinear up-sampling of both grids (`dvgo_resize_trilinear`,
        ATen's align_corners=True arithmetic) and a fresh occupancy mask = maxpool3(alpha) > fast_color_thres
        (`dvgo_alpha_maxpool_mask`), AND the coarse mask when there is one."""
        self._set_grid_resolution(num_voxels)
        size = tuple(int(s) for s in self.world_size)
        self.density = nn.Parameter(_C.ext.resize_trilinear(self.density.data.contiguous(), *size))
        if self.k0_dim > 0:
            self.k0 = nn.Parameter(_C.ext.resize_trilinear(self.k0.data.contiguous(), *size))
        else:
            self.k0 = nn.Parameter(torch.zeros([1, self.k0_dim, *size], device=self.density.device))
        mask = _C.ext.alpha_maxpool_mask(self.density.data, float(self.act_shift), float(self.voxel_size_ratio),
                                         float(self.fast_color_thres), None)
        if self.mask_cache_path:
            coarse = MaskCache(path=self.mask_cache_path, mask_cache_thres=self.mask_cache_thres).to(self.xyz_min.device)
            mask &= coarse(_grid_points(self.xyz_min, self.xyz_max, size))
        self.mask_cache = MaskCache(mask=mask, xyz_min=self.xyz_min, xyz_max=self.xyz_max)

    @torch.no_grad()
    def update_occupancy_cache(self):
        """The periodic occupancy refresh of the training loop (run.py:330-332), in place:
        mask &= maxpool3(alpha(density)) > fast_color_thres."""
        m = self.mask_cache.mask
        assert m.shape == self.density.shape[2:], "the refresh needs the mask at the density resolution (run.py:331)"
        m.copy_(_C.ext.alpha_maxpool_mask(self.density.data, float(self.act_shift), float(self.voxel_size_ratio),
                                          float(self.fast_color_thres), m))

    @torch.no_grad()
    def voxel_count_views(self, rays_o_tr, rays_d_tr, imsz, near, far, stepsize, downrate=1, irregular_shape=False):
        """How many training views see each voxel (lib/dvgo.py:265-295) -- the scatter-only use of the trilinear
        backward: grad of sum(sample(ones)) w.r.t. `ones`, thresholded at > 1 per view.  One kernel per view
        scatters the trilinear weights of every (ray, sample) straight into a per-view accumulator (no point
        tensor, no autograd graph), a second folds `acc > 1` into the count and re-zeroes it."""
        t0 = time.time()
        n_samples = int(np.linalg.norm(np.array(self.density.shape[2:]) + 1) / stepsize) + 1
        device = self.density.device
        stepdist = float(stepsize * self.voxel_size)
        count = torch.zeros_like(self.density.detach())
        acc = torch.zeros(tuple(self.density.shape[2:]), device=device)
        for ro_v, rd_v in zip(rays_o_tr.split(imsz), rays_d_tr.split(imsz)):
            if not irregular_shape:
                ro_v, rd_v = ro_v[::downrate, ::downrate], rd_v[::downrate, ::downrate]
            ro = ro_v.to(device).reshape(-1, 3).contiguous()
            rd = rd_v.to(device).reshape(-1, 3).contiguous()
            _C.ext.voxel_count_scatter(ro, rd, self.xyz_min, self.xyz_max, float(near), float(far), stepdist,
                                       n_samples, acc)
            _C.ext.voxel_count_commit(acc, count)
        self._last_count_seconds = time.time() - t0
        return count

    # ------------------------------------------------------------------ regularisers / activations
    def density_total_variation_add_grad(self, weight, dense_mode):
        w = weight * self.world_size.max() / 128  # lib/dvgo.py:297-300
        total_variation_cuda.total_variation_add_grad(self.density, self.density.grad, w, w, w, dense_mode)

    def k0_total_variation_add_grad(self, weight, dense_mode):
        w = weight * self.world_size.max() / 128  # lib/dvgo.py:302-305
        total_variation_cuda.total_variation_add_grad(self.k0, self.k0.grad, w, w, w, dense_mode)

    def activate_density(self, density, interval=None):
        interval = self.voxel_size_ratio if interval is None else interval
        return Raw2Alpha.apply(density.flatten(), float(self.act_shift), float(interval)).reshape(density.shape)

    def grid_sampler(self, xyz, *grids, mode=None, align_corners=True, is_k0=False, stepsize=None):
        """Trilinear lookup of each grid at world points `xyz` (lib/dvgo.py:312-328)."""
        if not align_corners or (mode not in (None, "bilinear")):
            raise NotImplementedError("only bilinear / align_corners=True (the reference's setting)")
        out = [grid_sample_trilinear(g, xyz, self.xyz_min, self.xyz_max) for g in grids]
        return out[0] if len(out) == 1 else out

    # ------------------------------------------------------------------ ray sampling
    def coarse_geo_scene(self, near, far, stepsize, **render_kwargs):
        """Scene constants for the occupancy-only kernels (hit test, training-ray preparation)."""
        mc = self.mask_cache
        X, Y, Z = (int(v) for v in self.density.shape[2:])
        return _C.ext.Scene(X, Y, Z, max(int(self.k0.shape[1]), 1), self.xyz_min, self.xyz_max, mc.mask,
                            mc.xyz2ijk_scale, mc.xyz2ijk_shift, float(near), float(far),
                            float(stepsize * self.voxel_size), float(self.act_shift),
                            float(stepsize * self.voxel_size_ratio), float(self.fast_color_thres), False, 0)

    @torch.no_grad()
    def hit_coarse_geo(self, rays_o, rays_d, near, far, stepsize, **render_kwargs):
        """Which rays touch occupied space at all (lib/dvgo.py:412-423): one warp per ray walks the samples and
        stops at the first one that is inside the bbox and occupied (`dvgo_hit_coarse_geo`)."""
        shape = rays_o.shape[:-1]
        dev = self.density.device
        rays_o = rays_o.reshape(-1, 3).to(dev).contiguous()
        rays_d = rays_d.reshape(-1, 3).to(dev).contiguous()
        hit = _C.ext.hit_coarse_geo(self.coarse_geo_scene(near, far, stepsize), rays_o, rays_d)
        return hit.reshape(shape)

    def sample_ray(self, rays_o, rays_d, near, far, stepsize, is_train=0, **render_kwargs):
        """Points on rays inside the bbox, near to far: (ray_pts [M,3], ray_id [M], step_id [M])
        (lib/dvgo.py:425-448)."""
        stepdist = float(stepsize * self.voxel_size)
        pts, outside, ray_id, step_id, _, _, _ = render_utils_cuda.sample_pts_on_rays(
            rays_o.contiguous(), rays_d.contiguous(), self.xyz_min, self.xyz_max, near, far, stepdist)
        keep = (~outside).nonzero(as_tuple=True)[0]
        return pts[keep], ray_id[keep], step_id[keep]

    # ------------------------------------------------------------------ forward
    def forward(self, rays_o, rays_d, viewdirs, global_step=None, **render_kwargs):
        """Volume rendering of N rays; returns the reference's dict (lib/dvgo.py:560-576):
        alphainv_last [N], weights [M], rgb_marched [N,3], raw_alpha [M], raw_rgb [M,3], ray_id [M]
        (+ depth [N] when render_kwargs['render_depth'])."""
        assert rays_o.dim() == 2 and rays_o.shape[-1] == 3, "Only suuport point queries in [N, 3] format"
        N = len(rays_o)
        pts, ray_id, step_id = self.sample_ray(rays_o=rays_o, rays_d=rays_d,
                                               is_train=global_step is not None, **render_kwargs)
        interval = render_kwargs["stepsize"] * self.voxel_size_ratio

        if self.mask_cache is not None:  # skip known free space (:469-473)
            keep = self.mask_cache(pts).nonzero(as_tuple=True)[0]
            pts, ray_id, step_id = pts[keep], ray_id[keep], step_id[keep]

        density = self.grid_sampler(pts, self.density)
        alpha = self.activate_density(density, interval)
        if self.fast_color_thres > 0:  # (:478-484)
            keep = (alpha > self.fast_color_thres).nonzero(as_tuple=True)[0]
            pts, ray_id, step_id, alpha = pts[keep], ray_id[keep], step_id[keep], alpha[keep]

        weights, alphainv_last = Alphas2Weights.apply(alpha, ray_id, N)
        if self.fast_color_thres > 0:  # (:488-494)
            keep = (weights > self.fast_color_thres).nonzero(as_tuple=True)[0]
            pts, ray_id, step_id = pts[keep], ray_id[keep], step_id[keep]
            alpha, weights = alpha[keep], weights[keep]

        if not self.rgbnet_full_implicit:
            k0 = self.grid_sampler(pts, self.k0, is_k0=True)
        if self.rgbnet is None:
            rgb = torch.sigmoid(k0)  # no view dependence (:512-514)
        else:
            emb = (viewdirs.unsqueeze(-1) * self.viewfreq).flatten(-2)
            emb = torch.cat([viewdirs, emb.sin(), emb.cos()], -1).flatten(0, -2)[ray_id]
            if self.posbase_pe > 0:
                pe = (pts.unsqueeze(-1) * self.posfreq).flatten(-2)
                feat = torch.cat([pts, pe.sin(), pe.cos(), emb], -1)
                rgb = torch.sigmoid(self.rgbnet(feat))
            elif self.rgbnet_direct:
                rgb = torch.sigmoid(self.rgbnet(torch.cat([k0, emb], -1)))
            else:
                rgb = torch.sigmoid(self.rgbnet(torch.cat([k0[:, 3:], emb], -1)) + k0[:, :3])

        # composite along each ray (:554-559)
        rgb_marched = segment_coo(src=weights.unsqueeze(-1) * rgb, index=ray_id,
                                  out=torch.zeros([N, 3], device=rgb.device), reduce="sum")
        rgb_marched = rgb_marched + alphainv_last.unsqueeze(-1) * render_kwargs["bg"]
        ret = {"alphainv_last": alphainv_last, "weights": weights, "rgb_marched": rgb_marched,
               "raw_alpha": alpha, "raw_rgb": rgb, "ray_id": ray_id}
        if render_kwargs.get("render_depth", False):
            with torch.no_grad():
                ret["depth"] = segment_coo(src=weights * step_id, index=ray_id,
                                           out=torch.zeros([N], device=rgb.device), reduce="sum")
        return ret
