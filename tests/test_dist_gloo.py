"""CPU tests (gloo, world_size 2) of the multi-GPU plumbing: ray sharding + gradient all-reduce must
reproduce the single-process full-batch step; view sharding must be a disjoint cover."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.util import make_rays


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _small_ref(seed=0):
    from oracle.model_ref import RefDVGO
    g = torch.Generator().manual_seed(seed)
    lo, hi = np.array([-1.0, -0.9, -0.8], np.float32), np.array([1.0, 0.9, 0.8], np.float32)
    dens = torch.randn(1, 1, 14, 13, 12, generator=g) * 3 + 2
    k0 = torch.randn(1, 12, 14, 13, 12, generator=g)
    net = [(torch.randn(128, 39, generator=g) * 0.1, torch.randn(128, generator=g) * 0.1),
           (torch.randn(128, 128, generator=g) * 0.1, torch.randn(128, generator=g) * 0.1),
           (torch.randn(3, 128, generator=g) * 0.1, torch.zeros(3))]
    return RefDVGO(lo, hi, dens, k0, net, None, act_shift=-4.595, voxel_size_ratio=1.0, fast_color_thres=1e-4)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from directvoxgo_b200_parallel_import import parallel  # set up by the parent (no CUDA extension needed)
    torch.set_num_threads(1)
    m = _small_ref()
    ro, rd, vd, tgt = make_rays(64, 5)
    n_global = len(ro)
    sro, srd, svd, stgt = parallel.shard_rays((ro, rd, vd, tgt), rank, world)
    ret = m.forward(sro.contiguous(), srd.contiguous(), svd.contiguous(), 0.2, 6.0, 0.5, 1.0)
    loss = m.loss(ret, stgt, len(sro), 1.0, 1e-3, 1e-2, n_global=n_global)
    loss.backward()
    grads = [p.grad for p in m.params().values()]
    parallel.allreduce_sum_(grads)
    total = parallel.global_loss(loss.detach())
    if rank == 0:
        out["loss"] = float(total)
        out["grads"] = [g.clone() for g in grads]
    dist.barrier()
    dist.destroy_process_group()


def _load_parallel_without_extension():
    """directvoxgo_b200/__init__ requires the CUDA extension; parallel.py itself is pure torch."""
    import importlib.util
    import sys
    import types
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("_dvgo_parallel", os.path.join(root, "directvoxgo_b200", "parallel.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    holder = types.ModuleType("directvoxgo_b200_parallel_import")
    holder.parallel = mod
    sys.modules["directvoxgo_b200_parallel_import"] = holder
    return mod


def test_shard_helpers():
    parallel = _load_parallel_without_extension()
    for n, w in [(8192, 1), (8192, 8), (10, 4), (3, 8), (0, 2)]:
        spans = [parallel.shard_bounds(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    views = [parallel.shard_views(200, r, 8) for r in range(8)]
    assert sorted(sum(views, [])) == list(range(200)) and len(views[0]) == 25


def _entry(rank, world, port, out):
    _load_parallel_without_extension()
    _worker(rank, world, port, out)


def test_ray_sharded_step_matches_full_batch():
    _load_parallel_without_extension()
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_entry, args=(world, port, out), nprocs=world, join=True)
    # single-process full batch
    m = _small_ref()
    ro, rd, vd, tgt = make_rays(64, 5)
    ret = m.forward(ro, rd, vd, 0.2, 6.0, 0.5, 1.0)
    loss = m.loss(ret, tgt, len(ro), 1.0, 1e-3, 1e-2)
    loss.backward()
    assert abs(out["loss"] - float(loss)) < 1e-6
    for g_dp, p in zip(out["grads"], m.params().values()):
        np.testing.assert_allclose(g_dp.numpy(), p.grad.numpy(), rtol=1e-4, atol=1e-7)
