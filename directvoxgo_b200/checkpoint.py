"""Checkpoint format compatibility -- row N4 of the scope table.

The reference's `{stage}_last.tar` (run.py:420-437) is a torch.save of
    {'global_step', 'model_kwargs': model.get_kwargs(), 'model_state_dict', 'optimizer_state_dict'}
with NCDHW grids and MaskedAdam state (`step`, `exp_avg`, `exp_avg_sq` per parameter, groups in the order of
the `lrate_*` keys, lib/utils.py:20-48).  Our modules keep the reference's constructor keywords and state_dict
keys, so the same file round-trips: `load_model` / `load_checkpoint` mirror lib/utils.py:53-79 and
`save_checkpoint` writes what run.py writes.  The fused trainer keeps its state in other layouts
(channel-last k0, flat rgbnet buffers); `FusedTrainer.optimizer_state_dict()` / `.load_optimizer_state_dict()`
convert at this boundary.
"""
import torch


def _load(path):
    # checkpoints hold numpy arrays in model_kwargs (xyz_min / xyz_max), hence weights_only=False
    return torch.load(path, map_location="cpu", weights_only=False)


def load_checkpoint(model, optimizer, ckpt_path, no_reload_optimizer):
    """lib/utils.py:53-60."""
    ckpt = _load(ckpt_path)
    start = ckpt["global_step"]
    model.load_state_dict(ckpt["model_state_dict"])
    if not no_reload_optimizer:
        optimizer.load_state_dict(ckpt["optimizer_state_dict"])
    return model, optimizer, start


def load_model(model_class, ckpt_path):
    """lib/utils.py:63-79."""
    ckpt = _load(ckpt_path)
    model = model_class(**ckpt["model_kwargs"])
    model.load_state_dict(ckpt["model_state_dict"])
    return model


def save_checkpoint(path, model, optimizer, global_step):
    """run.py:420-437.  `optimizer` is a MaskedAdam or a FusedTrainer (anything with the reference-format
    `state_dict()` / `optimizer_state_dict()`)."""
    if hasattr(optimizer, "optimizer_state_dict"):
        optimizer.sync_to_model()
        opt_sd = optimizer.optimizer_state_dict()
    else:
        opt_sd = optimizer.state_dict()
    torch.save({"global_step": global_step, "model_kwargs": model.get_kwargs(),
                "model_state_dict": model.state_dict(), "optimizer_state_dict": opt_sd}, path)
