"""Time the fused TV + MaskedAdam sweep alone (k0 [160^3,12] channel-last and density [160^3]) with CUDA events.
Usage: PYTHONPATH=. python tools/sweep_bench.py [grid]"""
import sys
import torch
from directvoxgo_b200 import ext

S = int(sys.argv[1]) if len(sys.argv) > 1 else 160
dev = "cuda"
for C in (12, 1):
    shape = (S, S, S, C) if C > 1 else (S, S, S)
    p = torch.randn(shape, device=dev)
    q = torch.empty_like(p)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for tv, dense, masked in ((True, True, True), (True, False, True), (False, False, True), (False, False, False)):
        g0 = torch.randn(shape, device=dev)
        if not dense:
            g0[torch.rand(shape, device=dev) < 0.9] = 0      # sparse scene: 10 % of the cells touched
        ts = []
        for it in range(12):
            g = g0.clone()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ext.sweep(p, q if tv else p, g, m, v, None, S, S, S, C, tv, dense, 1e-6, 1e-6, 1e-6, masked, it + 1, 0.9, 0.99,
                      0.1, 1e-8)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
            if tv:
                p, q = q, p
        t = sorted(ts[2:])[len(ts[2:]) // 2]
        gb = p.numel() * 32 / 1e9
        print("C=%2d tv=%d dense=%d masked=%d: %.4f ms  (%.0f GB/s on the dense 32 B/elem figure)" % (C, tv, dense, masked, t, gb / t * 1e3))
