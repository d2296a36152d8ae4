"""DirectMPIGO: forward-facing scenes as a multiplane image in NDC space, on the B200 kernels.

Constructor keywords, parameter names and `forward` dictionary follow the reference's
lib/dmpigo.py (`DirectMPIGO`, :17-25 ctor, :200-283 forward).  Differences from DirectVoxGO:
fixed-count NDC sampler (`sample_ndc_pts_on_rays`, :173-198), act_shift = 0 (:30), density
initialised per depth plane so that every plane contributes equally (:37-44), anisotropic TV
weights (:147-157), rgbnet input is [k0, viewdir-PE] with no diffuse split (:234-242).
"""
import functools

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import render_utils_cuda, total_variation_cuda
from .dvgo import MaskCache, _grid_points
from .ops import Alphas2Weights, Raw2Alpha, grid_sample_trilinear, segment_coo


@functools.lru_cache(maxsize=128)
def _full_ids(n_rays, n_steps, device):
    ray_id = torch.arange(n_rays, device=device).view(-1, 1).expand(n_rays, n_steps).flatten()
    step_id = torch.arange(n_steps, device=device).view(1, -1).expand(n_rays, n_steps).flatten()
    return ray_id, step_id


class DirectMPIGO(nn.Module):
    def __init__(self, xyz_min, xyz_max,
                 num_voxels=0, mpi_depth=0,
                 mask_cache_path=None, mask_cache_thres=1e-3,
                 fast_color_thres=0,
                 rgbnet_dim=0,
                 rgbnet_depth=3, rgbnet_width=128,
                 viewbase_pe=0,
                 **kwargs):
        super().__init__()
        self.register_buffer("xyz_min", torch.tensor(np.asarray(xyz_min, dtype=np.float32)))
        self.register_buffer("xyz_max", torch.tensor(np.asarray(xyz_max, dtype=np.float32)))
        self.fast_color_thres = fast_color_thres
        self.act_shift = 0
        self._set_grid_resolution(num_voxels, mpi_depth)

        self.density = nn.Parameter(torch.zeros([1, 1, *self.world_size]))
        with torch.no_grad():  # equal-contribution initialisation along depth (:37-44)
            g = np.full([mpi_depth], 1. / mpi_depth - 1e-6)
            p = [1 - g[0]] + [(1 - g[:i + 1].sum()) / (1 - g[:i].sum()) for i in range(1, len(g))]
            for i, pi in enumerate(p):
                self.density[..., i].fill_(np.log(pi ** (-1 / self.voxel_size_ratio) - 1))
            self.density[..., -1].fill_(10)

        self.rgbnet_kwargs = dict(rgbnet_dim=rgbnet_dim, rgbnet_depth=rgbnet_depth,
                                  rgbnet_width=rgbnet_width, viewbase_pe=viewbase_pe)
        if rgbnet_dim <= 0:
            self.k0_dim = 3
            self.k0 = nn.Parameter(torch.zeros([1, 3, *self.world_size]))
            self.rgbnet = None
        else:
            self.k0_dim = rgbnet_dim
            self.k0 = nn.Parameter(torch.zeros([1, rgbnet_dim, *self.world_size]))
            self.register_buffer("viewfreq", torch.tensor([2.0 ** i for i in range(viewbase_pe)]))
            dim0 = (3 + 3 * viewbase_pe * 2) + rgbnet_dim
            hidden = [nn.Sequential(nn.Linear(rgbnet_width, rgbnet_width), nn.ReLU(inplace=True))
                      for _ in range(rgbnet_depth - 2)]
            self.rgbnet = nn.Sequential(nn.Linear(dim0, rgbnet_width), nn.ReLU(inplace=True), *hidden,
                                        nn.Linear(rgbnet_width, 3))
            nn.init.constant_(self.rgbnet[-1].bias, 0)

        self.mask_cache_path = mask_cache_path
        self.mask_cache_thres = mask_cache_thres
        if mask_cache_path:
            coarse = MaskCache(path=mask_cache_path, mask_cache_thres=mask_cache_thres).to(self.xyz_min.device)
            mask = coarse(_grid_points(self.xyz_min, self.xyz_max, self.density.shape[2:]))
        else:
            mask = torch.ones(list(self.world_size), dtype=torch.bool)
        self.mask_cache = MaskCache(mask=mask, xyz_min=self.xyz_min, xyz_max=self.xyz_max)

    def _set_grid_resolution(self, num_voxels, mpi_depth):
        self.num_voxels, self.mpi_depth = num_voxels, mpi_depth
        extent_xy = (self.xyz_max - self.xyz_min)[:2]
        r = (num_voxels / mpi_depth / extent_xy.prod()).sqrt()
        self.world_size = torch.zeros(3, dtype=torch.long)
        self.world_size[:2] = (extent_xy * r).long()
        self.world_size[2] = mpi_depth
        self.voxel_size_ratio = 256. / mpi_depth

    def get_kwargs(self):
        return {
            "xyz_min": self.xyz_min.cpu().numpy(), "xyz_max": self.xyz_max.cpu().numpy(),
            "num_voxels": self.num_voxels, "mpi_depth": self.mpi_depth,
            "act_shift": self.act_shift, "voxel_size_ratio": self.voxel_size_ratio,
            "mask_cache_path": self.mask_cache_path, "mask_cache_thres": self.mask_cache_thres,
            "fast_color_thres": self.fast_color_thres,
            **self.rgbnet_kwargs,
        }

    @torch.no_grad()
    def scale_volume_grid(self, num_voxels, mpi_depth):
        self._set_grid_resolution(num_voxels, mpi_depth)
        size = tuple(int(s) for s in self.world_size)
        self.density = nn.Parameter(F.interpolate(self.density.data, size=size, mode="trilinear", align_corners=True))
        self.k0 = nn.Parameter(F.interpolate(self.k0.data, size=size, mode="trilinear", align_corners=True))
        alpha = F.max_pool3d(self.activate_density(self.density), kernel_size=3, padding=1, stride=1)[0, 0]
        self.mask_cache = MaskCache(mask=(alpha > self.fast_color_thres),
                                    xyz_min=self.xyz_min, xyz_max=self.xyz_max)

    def _tv_weights(self, weight):
        wxy = float(weight * self.world_size[:2].max() / 128)
        wz = float(weight * self.mpi_depth / 128)
        return wxy, wxy, wz

    def density_total_variation_add_grad(self, weight, dense_mode):
        total_variation_cuda.total_variation_add_grad(
            self.density, self.density.grad, *self._tv_weights(weight), dense_mode)

    def k0_total_variation_add_grad(self, weight, dense_mode):
        total_variation_cuda.total_variation_add_grad(
            self.k0, self.k0.grad, *self._tv_weights(weight), dense_mode)

    def activate_density(self, density, interval=None):
        interval = self.voxel_size_ratio if interval is None else interval
        return Raw2Alpha.apply(density.flatten(), 0., float(interval)).reshape(density.shape)

    def grid_sampler(self, xyz, grid):
        return grid_sample_trilinear(grid, xyz, self.xyz_min, self.xyz_max)

    def sample_ray(self, rays_o, rays_d, near, far, stepsize, is_train=False, **render_kwargs):
        """Fixed-count NDC sampling, t in [0,1] (lib/dmpigo.py:173-198)."""
        assert near == 0 and far == 1
        n_samples = int((self.mpi_depth - 1) / stepsize) + 1
        pts, outside = render_utils_cuda.sample_ndc_pts_on_rays(
            rays_o.contiguous(), rays_d.contiguous(), self.xyz_min, self.xyz_max, n_samples)
        inside = ~outside
        ray_id, step_id = _full_ids(inside.shape[0], inside.shape[1], pts.device)
        if bool(inside.all()):
            return pts.reshape(-1, 3), ray_id, step_id
        keep = inside.flatten().nonzero(as_tuple=True)[0]
        return pts.reshape(-1, 3)[keep], ray_id[keep], step_id[keep]

    def forward(self, rays_o, rays_d, viewdirs, global_step=None, **render_kwargs):
        assert rays_o.dim() == 2 and rays_o.shape[-1] == 3, "Only suuport point queries in [N, 3] format"
        N = len(rays_o)
        pts, ray_id, step_id = self.sample_ray(rays_o=rays_o, rays_d=rays_d,
                                               is_train=global_step is not None, **render_kwargs)
        interval = render_kwargs["stepsize"] * self.voxel_size_ratio

        if self.mask_cache is not None:
            keep = self.mask_cache(pts).nonzero(as_tuple=True)[0]
            pts, ray_id, step_id = pts[keep], ray_id[keep], step_id[keep]

        density = self.grid_sampler(pts, self.density)
        alpha = self.activate_density(density, interval)
        if self.fast_color_thres > 0:
            keep = (alpha > self.fast_color_thres).nonzero(as_tuple=True)[0]
            pts, ray_id, step_id, alpha = pts[keep], ray_id[keep], step_id[keep], alpha[keep]

        weights, alphainv_last = Alphas2Weights.apply(alpha, ray_id, N)
        if self.fast_color_thres > 0:
            keep = (weights > self.fast_color_thres).nonzero(as_tuple=True)[0]
            pts, ray_id, step_id = pts[keep], ray_id[keep], step_id[keep]
            alpha, weights = alpha[keep], weights[keep]

        vox_emb = self.grid_sampler(pts, self.k0)
        if self.rgbnet is None:
            rgb = torch.sigmoid(vox_emb)
        else:
            emb = (viewdirs.unsqueeze(-1) * self.viewfreq).flatten(-2)
            emb = torch.cat([viewdirs, emb.sin(), emb.cos()], -1)[ray_id]
            rgb = torch.sigmoid(self.rgbnet(torch.cat([vox_emb, emb], -1)))

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
