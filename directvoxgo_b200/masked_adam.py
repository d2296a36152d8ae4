"""MaskedAdam on the B200 kernels -- same optimiser contract as the reference's lib/masked_adam.py:17-71.

Three update rules, chosen per parameter tensor exactly as the reference chooses them (:60-71):
  * a per-voxel learning-rate table has been set and has this parameter's shape -> `adam_upd_with_perlr`
  * else the group's `skip_zero_grad` flag is set                               -> `masked_adam_upd`
  * else                                                                        -> `adam_upd`
State keys (`step`, `exp_avg`, `exp_avg_sq`) and param-group keys (`lr`, `betas`, `eps`,
`skip_zero_grad`) are the reference's, so optimiser state_dicts are interchangeable (run.py:420-437).
Note this is not torch.optim.Adam: eps is added to the un-corrected sqrt(v) (adam_upd_kernel.cu:21,72).
"""
import torch

from . import adam_upd_cuda


def _validate(lr, betas, eps):
    if lr < 0.0:
        raise ValueError("Invalid learning rate: {}".format(lr))
    if eps < 0.0:
        raise ValueError("Invalid epsilon value: {}".format(eps))
    for i, b in enumerate(betas):
        if not 0.0 <= b < 1.0:
            raise ValueError("Invalid beta parameter at index {}: {}".format(i, b))


class MaskedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.99), eps=1e-8):
        _validate(lr, betas, eps)
        self.per_lr = None
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    def set_pervoxel_lr(self, count):
        first = self.param_groups[0]["params"][0]
        assert first.shape == count.shape
        self.per_lr = (count.float() / count.max()).contiguous()

    def _state_for(self, p):
        st = self.state[p]
        if not st:
            st["step"] = 0
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    @torch.no_grad()
    def step(self):
        for group in self.param_groups:
            b1, b2 = group["betas"]
            lr, eps = group["lr"], group["eps"]
            masked = group.get("skip_zero_grad", False)
            for p in group["params"]:
                g = p.grad
                if g is None:
                    continue
                st = self._state_for(p)
                st["step"] += 1
                args = (p, g, st["exp_avg"], st["exp_avg_sq"])
                tail = (st["step"], b1, b2, lr, eps)
                if self.per_lr is not None and p.shape == self.per_lr.shape:
                    adam_upd_cuda.adam_upd_with_perlr(*args, self.per_lr, *tail)
                elif masked:
                    adam_upd_cuda.masked_adam_upd(*args, *tail)
                else:
                    adam_upd_cuda.adam_upd(*args, *tail)


def create_optimizer_or_freeze_model(model, cfg_train, global_step):
    """Param-group contract of the reference's lib/utils.py:20-48: one group per `lrate_<attr>` key of
    cfg_train whose attribute exists on the model; lr decayed by 0.1**(global_step/(lrate_decay*1000));
    lr <= 0 freezes the attribute; `skip_zero_grad` set for names in cfg_train.skip_zero_grad_fields.
    `cfg_train` may be a dict or any object with attribute access."""
    get = (lambda k, d=None: cfg_train.get(k, d)) if isinstance(cfg_train, dict) else \
        (lambda k, d=None: getattr(cfg_train, k, d))
    keys = list(cfg_train.keys()) if hasattr(cfg_train, "keys") else [k for k in dir(cfg_train)]
    decay = 0.1 ** (global_step / (get("lrate_decay") * 1000))
    skip_fields = get("skip_zero_grad_fields", []) or []
    groups = []
    for key in keys:
        if not key.startswith("lrate_") or key == "lrate_decay":
            continue
        name = key[len("lrate_"):]
        target = getattr(model, name, None)
        if target is None:
            continue
        lr = get(key) * decay
        if lr > 0:
            params = target.parameters() if isinstance(target, torch.nn.Module) else target
            groups.append({"params": params, "lr": lr, "skip_zero_grad": name in skip_fields})
        elif isinstance(target, torch.nn.Module):
            for q in target.parameters():
                q.requires_grad = False
        else:
            target.requires_grad = False
    return MaskedAdam(groups)
