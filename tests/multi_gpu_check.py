"""Run under torchrun on >= 2 GPUs: the ray-sharded FusedTrainer must reproduce the single-GPU step on the full batch
in all three gradient-exchange modes: "peer" (one kernel = reduce-scatter + sweep + all-gather over NVLink peer
memory), NCCL reduce-scatter -> slab-sharded sweep -> all-gather, and the plain NCCL all-reduce variant.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/multi_gpu_check.py
"""
import copy
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import datetime
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
    from directvoxgo_b200 import synthetic as syn
    from directvoxgo_b200.dvgo import DirectVoxGO
    from directvoxgo_b200.fused import FusedTrainer
    from directvoxgo_b200.parallel import shard_rays

    lo, hi = syn.fine_bbox()
    kw = dict(syn.FINE_MODEL, num_voxels=48 ** 3, num_voxels_base=48 ** 3)
    torch.manual_seed(0)
    base = DirectVoxGO(lo, hi, **kw)
    syn.randomize_grids_(base, 1)
    with torch.no_grad():
        base.density.mul_(2.0)
    base = base.to(dev)
    cfg, rk = dict(syn.FINE_TRAIN), dict(syn.RENDER_KWARGS)
    ok = True
    # mlp='torch': exact fp32 rgbnet, NCCL for its gradient; mlp='tc': tensor-core rgbnet, flag barriers + peer Adam
    for exchange, shard, mlp, ltol in (("auto", True, "torch", 2e-5), ("peer_p2p", True, "torch", 2e-5),
                                       ("nccl", True, "torch", 2e-5), ("nccl", False, "torch", 2e-5),
                                       ("auto", True, "tc", 1e-3), ("peer_p2p", True, "tc", 1e-3), ("peer", True, "tc", 1e-3)):
        m_dp, m_one = copy.deepcopy(base), copy.deepcopy(base)
        try:
            t_dp = FusedTrainer(m_dp, cfg, rk, world_size=world, mlp=mlp, shard_sweep=shard, exchange=exchange)
        except RuntimeError as e:      # "peer" (multicast required) on a system without NVLS
            if rank == 0:
                print("exchange=%s mlp=%s unavailable: %s" % (exchange, mlp, str(e)[:80]))
            continue
        if rank == 0:
            print("requested exchange=%s shard_sweep=%s mlp=%s -> running %s (multicast %s, flag barriers %s)" % (
                exchange, shard, mlp, t_dp.exchange, getattr(t_dp, "multicast", False), t_dp._flags is not None))
        t_one = FusedTrainer(m_one, cfg, rk, world_size=1, mlp=mlp)
        for it in range(3):
            batch = syn.random_training_rays(4096, n_views=20, seed=90 + it, device=dev)
            mine = tuple(t.contiguous() for t in shard_rays(batch, rank, world))
            l_dp = t_dp.step(*mine)
            dist.all_reduce(l_dp)
            l_one = t_one.step(*batch)
            if abs(float(l_dp) - float(l_one)) > ltol * max(1.0, abs(float(l_one))):
                ok = False
                print("rank", rank, t_dp.exchange, "shard", shard, "it", it, "loss mismatch", float(l_dp), float(l_one))
        t_dp.sync_to_model(); t_one.sync_to_model()
        for name in ("density", "k0"):
            d = (getattr(m_dp, name) - getattr(m_one, name)).abs()
            med, q = float(d.median()), float(torch.quantile(d.flatten()[:4_000_000], 0.999))
            if not (med < (1e-4 if mlp == "torch" else 2e-3) and q < (5e-3 if mlp == "torch" else 0.25)):
                ok = False
            if rank == 0:
                print("%s shard_sweep=%s %s: median |d| %.2e, 99.9%% %.2e" % (t_dp.exchange, shard, name, med, q))
        # every rank must hold identical parameters after the gather
        chk = t_dp.k0.double().sum()
        lo_, hi_ = chk.clone(), chk.clone()
        dist.all_reduce(lo_, op=dist.ReduceOp.MIN); dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
        if float(hi_ - lo_) != 0.0:
            ok = False
            print("replicas diverged", float(lo_), float(hi_))
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if float(flag) == 1.0 else "FAIL")
    dist.destroy_process_group()
    sys.exit(0 if float(flag) == 1.0 else 1)


if __name__ == "__main__":
    main()
