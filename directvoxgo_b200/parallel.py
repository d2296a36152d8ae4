"""Multi-GPU plumbing for the two modes the hot path shards in.  One process per GPU.

Training  : rays are independent given replicated grids.  Each rank marches its own slice of the batch,
            loss normalisers use the GLOBAL ray count, grid / rgbnet gradients are summed with one
            all-reduce per buffer (NCCL over NVLink on the B200 box, gloo in the CPU tests), then every
            rank applies the identical TV + Adam sweep to its replica.
Rendering : whole views are independent: view i -> rank i mod world, no communication.
"""
import torch


def shard_bounds(n, rank, world):
    """Contiguous, balanced slice [lo, hi) of n items for `rank` (first n % world ranks get one extra)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_rays(tensors, rank, world):
    """Slice every [N, ...] tensor of a ray batch to this rank's contiguous share."""
    n = tensors[0].shape[0]
    lo, hi = shard_bounds(n, rank, world)
    return tuple(t[lo:hi] for t in tensors)


def shard_views(n_views, rank, world):
    """Indices of the views this rank renders (round-robin keeps per-rank work balanced)."""
    return list(range(rank, n_views, world))


def allreduce_sum_(tensors, group=None):
    """In-place sum over ranks of each tensor (gradient accumulators)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return tensors
    for t in tensors:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return tensors


def global_loss(local_loss, group=None):
    """Sum of the per-rank loss shares (each rank's share is already divided by the global ray count)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_loss
    out = local_loss.clone()
    dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out
