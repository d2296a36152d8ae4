"""Training-step drivers for the hot path (what run.py:372-397 does once per iteration).

`ModuleTrainer`  -- the op-by-op drop-in path: DirectVoxGO module on our kernels + torch autograd +
                    TV + MaskedAdam, i.e. exactly the call sequence of the reference's loop body.
`FusedTrainer`   -- (fused.py) the B200 fast path with its own grid layout; same step() contract.

Both expose
    step(rays_o, rays_d, viewdirs, target) -> loss tensor (0-dim, on device, no sync)
and are what bench.py times and what the multi-GPU wrapper shards.
"""
import torch
import torch.nn.functional as F

from .masked_adam import create_optimizer_or_freeze_model


def training_loss(ret, target, n_rays, weight_main=1.0, weight_entropy_last=0.0, weight_rgbper=0.0):
    """Photometric + background-entropy + per-point colour loss (run.py:377-386)."""
    loss = weight_main * F.mse_loss(ret["rgb_marched"], target)
    if weight_entropy_last > 0:
        pout = ret["alphainv_last"].clamp(1e-6, 1 - 1e-6)
        entropy = -(pout * torch.log(pout) + (1 - pout) * torch.log(1 - pout)).mean()
        loss = loss + weight_entropy_last * entropy
    if weight_rgbper > 0:
        rgbper = (ret["raw_rgb"] - target[ret["ray_id"]]).pow(2).sum(-1)
        loss = loss + weight_rgbper * (rgbper * ret["weights"].detach()).sum() / n_rays
    return loss


def tv_schedule(cfg, global_step):
    """(tv on?, dense mode?) at this iteration: the gate of run.py:389-395 when cfg carries the reference's
    tv_before / tv_after / tv_every / tv_dense_before keys, else the simple `tv_dense` switch of the benchmarks."""
    if "tv_before" in cfg or "tv_after" in cfg:
        on = (global_step < cfg.get("tv_before", 0) and global_step > cfg.get("tv_after", 0)
              and global_step % cfg.get("tv_every", 1) == 0)
        return on, global_step < cfg.get("tv_dense_before", 0)
    return True, bool(cfg.get("tv_dense", True))


class ModuleTrainer:
    def __init__(self, model, cfg_train, render_kwargs, dist_group=None, world_size=1, global_step=0):
        self.model = model
        self.cfg = dict(cfg_train)
        self.rk = dict(render_kwargs)
        self.opt = create_optimizer_or_freeze_model(model, self.cfg, global_step=global_step)
        self.global_step = global_step
        self.world_size = world_size
        self.dist_group = dist_group
        self.decay = 0.1 ** (1.0 / (self.cfg.get("lrate_decay", 20) * 1000))   # run.py:401-403

    def _allreduce_grads(self):
        import torch.distributed as dist
        for p in self.model.parameters():
            if p.grad is not None:
                dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=self.dist_group)

    def step(self, rays_o, rays_d, viewdirs, target):
        cfg, m = self.cfg, self.model
        self.global_step += 1
        n_global = len(rays_o) * self.world_size  # loss normalisers use the global batch
        ret = m(rays_o, rays_d, viewdirs, global_step=self.global_step, **self.rk)
        self.opt.zero_grad(set_to_none=True)
        loss = training_loss(ret, target, len(rays_o), cfg.get("weight_main", 1.0),
                             cfg.get("weight_entropy_last", 0.0), cfg.get("weight_rgbper", 0.0))
        if self.world_size > 1:
            loss = loss / self.world_size
        loss.backward()
        if self.world_size > 1:
            self._allreduce_grads()
        tv_on, dense = tv_schedule(cfg, self.global_step)
        if tv_on and cfg.get("weight_tv_density", 0) > 0:
            m.density_total_variation_add_grad(cfg["weight_tv_density"] / n_global, dense)
        if tv_on and cfg.get("weight_tv_k0", 0) > 0:
            m.k0_total_variation_add_grad(cfg["weight_tv_k0"] / n_global, dense)
        self.opt.step()
        for group in self.opt.param_groups:     # run.py:401-406: per-iteration exponential lr decay
            group["lr"] = group["lr"] * self.decay
        return loss.detach()


class HostFedLoop:
    """Drives a trainer from HOST batches (pinned memory) without idling the GPU between steps.

    The reference keeps its training rays on the device and reads the loss only at its logging cadence
    (run.py:353-370, :409-417); a caller that streams ray batches from the host and wants every step's loss would
    otherwise serialise copy -> step -> read.  Here the host -> device copy of step i runs on a copy stream under the
    kernels of step i - 1 (two staging sets), and the loss of step i is read one call later, through a pinned buffer
    and an event, after step i + 1 has been enqueued:

        loop = HostFedLoop(trainer, example_batch)
        for batch in host_batches:            # tuples of pinned CPU tensors (rays_o, rays_d, viewdirs, target)
            prev = loop.step(batch)           # loss of the PREVIOUS step (None on the first call)
        last = loop.drain()                   # loss of the final step

    Every step still pays its own H2D copy and its own 4-byte D2H read; only the order of the waits changes.
    """

    def __init__(self, trainer, example_batch, device=None):
        self.trainer = trainer
        self.device = torch.device(device) if device is not None else trainer.device
        self.staging = [[torch.empty_like(x, device=self.device) for x in example_batch] for _ in range(2)]
        self.loss_host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.copied = [torch.cuda.Event() for _ in range(2)]
        self.done = [None, None]
        self.loss_ready = [torch.cuda.Event() for _ in range(2)]
        self.i = 0

    def step(self, host_batch):
        b = self.i % 2
        cur = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self.copy_stream):
            if self.done[b] is not None:          # step i - 2 has finished reading this staging set
                self.copy_stream.wait_event(self.done[b])
            for d, h in zip(self.staging[b], host_batch):
                d.copy_(h, non_blocking=True)
            self.copied[b].record(self.copy_stream)
        cur.wait_event(self.copied[b])
        loss = self.trainer.step(*self.staging[b])
        if self.done[b] is None:
            self.done[b] = torch.cuda.Event()
        self.done[b].record(cur)
        self.loss_host[b].copy_(loss.reshape(()), non_blocking=True)
        self.loss_ready[b].record(cur)
        prev = None
        if self.i > 0:
            self.loss_ready[1 - b].synchronize()
            prev = float(self.loss_host[1 - b])
        self.i += 1
        return prev

    def drain(self):
        """Loss of the last step enqueued (waits for it)."""
        if self.i == 0:
            return None
        b = (self.i - 1) % 2
        self.loss_ready[b].synchronize()
        return float(self.loss_host[b])
