"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's model-level hot path.

`RefDVGO` restates, with torch-CPU ops plus the C oracle for the reference's custom CUDA ops,
  * DirectVoxGO.forward            lib/dvgo.py:450-577  (sample_ray :425-448, the 4-mask cascade,
                                   grid_sampler :312-328 via the real ATen F.grid_sample on CPU,
                                   rgbnet :524-541, segment_coo compositing :554-576)
  * the training loss              run.py:377-386
  * TV + MaskedAdam                run.py:389-397, lib/dvgo.py:297-305, lib/masked_adam.py:39-71
It is what the reference would do on a CPU if its custom ops had a CPU path (they do not:
CHECK_CUDA at lib/cuda/render_utils.cpp:40), in the spirit of the authors' pre-CUDA "native
pytorch" version (IMPROVING_LOG.md:12,37).  Used as the parity checker and, timed on the host
cores, as bench.py's `cpu_baseline` / `--impl reference` arm.  The product never imports it.

Pinned by tests/golden/refpy_*.npz: outputs of the reference's OWN lib/dvgo.py + lib/masked_adam.py
run in the build container (oracle/make_golden_refpy.py), which this file must reproduce.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import oracle as orc


def set_ops(ns):
    """Swap the provider of the reference's custom ops (default: the C oracle on the CPU).  The GPU tests pass a
    namespace built from oracle/_ref -- the reference's OWN CUDA kernels -- to run and time the reference's op
    sequence on the same B200 (tests/test_gpu_0_vs_ref.py).  Returns the previous provider."""
    global orc
    prev, orc = orc, ns
    return prev


class _Raw2Alpha(torch.autograd.Function):  # lib/dvgo.py:618-642
    @staticmethod
    def forward(ctx, density, shift, interval):
        e, a = orc.raw2alpha(density, shift, interval)
        ctx.save_for_backward(e)
        ctx.interval = interval
        return a

    @staticmethod
    def backward(ctx, g):
        return orc.raw2alpha_backward(ctx.saved_tensors[0], g.contiguous(), ctx.interval), None, None


class _Alphas2Weights(torch.autograd.Function):  # lib/dvgo.py:644-660
    @staticmethod
    def forward(ctx, alpha, ray_id, N):
        w, T, last, i_s, i_e = orc.alpha2weight(alpha, ray_id, N)
        ctx.save_for_backward(alpha, w, T, last, i_s, i_e)
        ctx.n_rays = N
        return w, last

    @staticmethod
    def backward(ctx, gw, gl):
        alpha, w, T, last, i_s, i_e = ctx.saved_tensors
        return orc.alpha2weight_backward(alpha, w, T, last, i_s, i_e, ctx.n_rays,
                                         gw.contiguous(), gl.contiguous()), None, None


def grid_sampler(xyz, grid, xyz_min, xyz_max):
    """lib/dvgo.py:312-328 verbatim in meaning: ATen grid_sample on ind_norm, [C,P] -> [P,C]."""
    shape = xyz.shape[:-1]
    ind = ((xyz.reshape(1, 1, 1, -1, 3) - xyz_min) / (xyz_max - xyz_min)).flip((-1,)) * 2 - 1
    out = F.grid_sample(grid, ind, mode="bilinear", align_corners=True)
    out = out.reshape(grid.shape[1], -1).T.reshape(*shape, grid.shape[1])
    return out.squeeze(-1) if grid.shape[1] == 1 else out


class RefDVGO:
    """Fine- or coarse-stage DirectVoxGO state on the CPU: density [1,1,X,Y,Z], k0 [1,C,X,Y,Z],
    optional rgbnet weights (list of (W,b) for Linear layers, ReLU between), all-true or given mask."""

    def __init__(self, xyz_min, xyz_max, density, k0, rgbnet=None, mask=None, act_shift=0.0,
                 voxel_size_ratio=1.0, voxel_size=None, fast_color_thres=0.0, viewbase_pe=4,
                 rgbnet_direct=True):
        self.xyz_min = torch.as_tensor(xyz_min, dtype=torch.float32)
        self.xyz_max = torch.as_tensor(xyz_max, dtype=torch.float32)
        self.density = density.detach().clone().float().requires_grad_()
        self.k0 = k0.detach().clone().float().requires_grad_()
        self.rgbnet = None if rgbnet is None else [
            (W.detach().clone().float().requires_grad_(), b.detach().clone().float().requires_grad_())
            for W, b in rgbnet]
        X, Y, Z = self.density.shape[2:]
        self.mask = torch.ones(X, Y, Z, dtype=torch.bool) if mask is None else mask.bool().clone()
        self.xyz2ijk_scale = (torch.tensor([X, Y, Z], dtype=torch.float32) - 1) / (self.xyz_max - self.xyz_min)
        self.xyz2ijk_shift = -self.xyz_min * self.xyz2ijk_scale  # lib/dvgo.py:600-602
        self.act_shift = float(act_shift)
        self.voxel_size_ratio = float(voxel_size_ratio)
        if voxel_size is None:
            voxel_size = ((self.xyz_max - self.xyz_min).prod() / (X * Y * Z)).pow(1 / 3)
        self.voxel_size = float(voxel_size)
        self.fast_color_thres = fast_color_thres
        self.viewfreq = torch.tensor([2.0 ** i for i in range(viewbase_pe)])
        self.rgbnet_direct = rgbnet_direct
        self.opt_state = {}

    def to(self, device):
        """Move the state (GPU tests only: reference CUDA kernels from oracle/_ref as the op provider)."""
        for n in ("xyz_min", "xyz_max", "mask", "xyz2ijk_scale", "xyz2ijk_shift", "viewfreq"):
            setattr(self, n, getattr(self, n).to(device))
        self.density = self.density.detach().to(device).requires_grad_()
        self.k0 = self.k0.detach().to(device).requires_grad_()
        if self.rgbnet is not None:
            self.rgbnet = [(W.detach().to(device).requires_grad_(), b.detach().to(device).requires_grad_())
                           for W, b in self.rgbnet]
        return self

    def params(self):
        ps = {"density": self.density, "k0": self.k0}
        if self.rgbnet is not None:
            for i, (W, b) in enumerate(self.rgbnet):
                ps["rgbnet.%d.weight" % i] = W
                ps["rgbnet.%d.bias" % i] = b
        return ps

    def mlp(self, x):
        for i, (W, b) in enumerate(self.rgbnet):
            x = F.linear(x, W, b)
            if i + 1 < len(self.rgbnet):
                x = F.relu(x)
        return x

    def forward(self, rays_o, rays_d, viewdirs, near, far, stepsize, bg, render_depth=False):
        N = len(rays_o)
        stepdist = stepsize * self.voxel_size
        pts, outside, ray_id, step_id, N_steps, t_min, t_max = orc.sample_pts_on_rays(
            rays_o, rays_d, self.xyz_min, self.xyz_max, near, far, stepdist)
        keep = ~outside  # lib/dvgo.py:444-447
        pts, ray_id, step_id = pts[keep], ray_id[keep], step_id[keep]
        interval = stepsize * self.voxel_size_ratio
        m = orc.maskcache_lookup(self.mask, pts, self.xyz2ijk_scale, self.xyz2ijk_shift)  # :469-473
        pts, ray_id, step_id = pts[m], ray_id[m], step_id[m]
        density = grid_sampler(pts, self.density, self.xyz_min, self.xyz_max)  # :476
        alpha = _Raw2Alpha.apply(density.flatten(), self.act_shift, interval)  # :477
        if self.fast_color_thres > 0:  # :478-484
            m = alpha > self.fast_color_thres
            pts, ray_id, step_id, alpha = pts[m], ray_id[m], step_id[m], alpha[m]
        weights, alphainv_last = _Alphas2Weights.apply(alpha, ray_id, N)  # :487
        if self.fast_color_thres > 0:  # :488-494
            m = weights > self.fast_color_thres
            weights, alpha, pts, ray_id, step_id = weights[m], alpha[m], pts[m], ray_id[m], step_id[m]
        k0 = grid_sampler(pts, self.k0, self.xyz_min, self.xyz_max)  # :509
        if self.rgbnet is None:
            rgb = torch.sigmoid(k0)  # :512-514
        else:
            emb = (viewdirs.unsqueeze(-1) * self.viewfreq).flatten(-2)  # :524-526
            emb = torch.cat([viewdirs, emb.sin(), emb.cos()], -1)[ray_id]
            if self.rgbnet_direct:
                rgb = torch.sigmoid(self.mlp(torch.cat([k0, emb], -1)))  # :536-539
            else:
                rgb = torch.sigmoid(self.mlp(torch.cat([k0[:, 3:], emb], -1)) + k0[:, :3])  # :541
        rgb_marched = torch.zeros(N, 3, device=rays_o.device).index_add(0, ray_id, weights.unsqueeze(-1) * rgb)  # :554-558
        rgb_marched = rgb_marched + alphainv_last.unsqueeze(-1) * bg  # :559
        ret = {"alphainv_last": alphainv_last, "weights": weights, "rgb_marched": rgb_marched,
               "raw_alpha": alpha, "raw_rgb": rgb, "ray_id": ray_id, "step_id": step_id,
               "N_steps": N_steps, "t_min": t_min, "t_max": t_max}
        if render_depth:
            with torch.no_grad():  # :569-576
                ret["depth"] = torch.zeros(N, device=rays_o.device).index_add(0, ray_id, weights * step_id)
        return ret

    @staticmethod
    def loss(ret, target, n_rays, weight_main=1.0, weight_entropy_last=0.0, weight_rgbper=0.0, n_global=None):
        """run.py:377-386.  With n_global (ray-sharded data parallel) every mean runs over the GLOBAL batch:
        this rank's share of the loss is returned and n_rays is ignored."""
        share = 1.0 if n_global is None else len(target) / n_global
        if n_global is not None:
            n_rays = n_global
        loss = weight_main * F.mse_loss(ret["rgb_marched"], target) * share
        if weight_entropy_last > 0:
            pout = ret["alphainv_last"].clamp(1e-6, 1 - 1e-6)
            ent = -(pout * torch.log(pout) + (1 - pout) * torch.log(1 - pout)).mean() * share
            loss = loss + weight_entropy_last * ent
        if weight_rgbper > 0:
            rgbper = (ret["raw_rgb"] - target[ret["ray_id"]]).pow(2).sum(-1)
            loss = loss + weight_rgbper * (rgbper * ret["weights"].detach()).sum() / n_rays
        return loss

    def zero_grad(self):
        for p in self.params().values():
            p.grad = None

    def tv_add_grad(self, weight_density, weight_k0, n_rays, dense_mode):
        """run.py:389-395 + lib/dvgo.py:297-305."""
        wmax = float(max(self.density.shape[2:]))
        for p, wt in ((self.density, weight_density), (self.k0, weight_k0)):
            if wt > 0 and p.grad is not None:
                w = wt / n_rays * wmax / 128
                orc.total_variation_add_grad(p, p.grad, w, w, w, dense_mode)

    def adam_step(self, lrs, skip_zero_grad=("density", "k0"), betas=(0.9, 0.99), eps=1e-8, per_lr=None):
        """lib/masked_adam.py:39-71.  lrs: {'density':..,'k0':..,'rgbnet':..}."""
        for name, p in self.params().items():
            if p.grad is None:
                continue
            group = name.split(".")[0]
            st = self.opt_state.setdefault(name, {"step": 0, "exp_avg": torch.zeros_like(p),
                                                  "exp_avg_sq": torch.zeros_like(p)})
            st["step"] += 1
            args = (p, p.grad.contiguous(), st["exp_avg"], st["exp_avg_sq"])
            tail = (st["step"], betas[0], betas[1], lrs[group], eps)
            if per_lr is not None and p.shape == per_lr.shape:
                orc.adam_upd_with_perlr(*args, per_lr, *tail)
            elif group in skip_zero_grad:
                orc.masked_adam_upd(*args, *tail)
            else:
                orc.adam_upd(*args, *tail)

    def train_step(self, rays_o, rays_d, viewdirs, target, render_kwargs, cfg):
        """One iteration of run.py:372-397 (fwd, loss, bwd, TV, MaskedAdam).  Returns the loss."""
        ret = self.forward(rays_o, rays_d, viewdirs, render_kwargs["near"], render_kwargs["far"],
                           render_kwargs["stepsize"], render_kwargs["bg"])
        self.zero_grad()
        loss = self.loss(ret, target, len(rays_o), cfg.get("weight_main", 1.0),
                         cfg.get("weight_entropy_last", 0.0), cfg.get("weight_rgbper", 0.0))
        loss.backward()
        if cfg.get("weight_tv_density", 0) > 0 or cfg.get("weight_tv_k0", 0) > 0:
            self.tv_add_grad(cfg.get("weight_tv_density", 0), cfg.get("weight_tv_k0", 0), len(rays_o),
                             cfg.get("tv_dense", True))
        self.adam_step({"density": cfg["lrate_density"], "k0": cfg["lrate_k0"],
                        "rgbnet": cfg.get("lrate_rgbnet", 0.0)},
                       skip_zero_grad=tuple(cfg.get("skip_zero_grad_fields", ())))
        return float(loss.detach()), ret

    @classmethod
    def from_module(cls, model):
        """Snapshot a directvoxgo_b200.dvgo.DirectVoxGO (or the reference's lib.dvgo.DirectVoxGO)."""
        rgbnet = None
        if model.rgbnet is not None:
            lin = [m for m in model.rgbnet.modules() if isinstance(m, torch.nn.Linear)]
            rgbnet = [(l.weight.detach().cpu(), l.bias.detach().cpu()) for l in lin]
        return cls(model.xyz_min.cpu(), model.xyz_max.cpu(), model.density.detach().cpu(),
                   model.k0.detach().cpu(), rgbnet, model.mask_cache.mask.cpu(), float(model.act_shift),
                   float(model.voxel_size_ratio), float(model.voxel_size), model.fast_color_thres,
                   viewbase_pe=len(model.viewfreq) if rgbnet is not None else 0,
                   rgbnet_direct=getattr(model, "rgbnet_direct", True))
