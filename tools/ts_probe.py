"""TMEM-sourced A operand probe (tcgen05.mma with [a_tmem]): correctness of the fp16 packing assumption and the cost
per MMA against the shared-memory-A form.   PYTHONPATH=. python tools/ts_probe.py"""
import torch
from directvoxgo_b200 import ext

torch.manual_seed(0)
for N, K, b_mn in ((16, 128, False), (16, 128, True), (128, 128, False), (128, 48, False), (48, 128, True), (64, 16, False)):
    A = torch.randn(128, K, device="cuda")
    B = torch.randn(K, N, device="cuda") if b_mn else torch.randn(N, K, device="cuda")
    D, cyc = ext.tc_ts_probe(A, B, N, K, b_mn, 200)
    torch.cuda.synchronize()
    want = A.half().float() @ (B.half().float() if b_mn else B.half().float().T)
    err = float((D - want).abs().max())
    ks = K // 16
    print("N=%3d K=%3d b_mn=%d: max|D - A B^T| = %.3e (|want| max %.1f)  TS %.1f cyc/MMA, SS %.1f cyc/MMA" % (
        N, K, b_mn, err, float(want.abs().max()), float(cyc[0]) / (200 * ks), float(cyc[1]) / (200 * ks)))
    if err > 5e-2:
        # which A element does the tensor core read for (row, k)?  unit-vector probe
        A1 = torch.zeros(128, K, device="cuda"); A1[:, 0] = 1; A1[:, 1] = 2; A1[:, 2] = 4; A1[:, 3] = 8
        B1 = torch.zeros_like(B)
        for k in range(min(K, 16)):
            (B1.__setitem__((k, k % N), 1.0) if b_mn else B1.__setitem__((k % N, k), 1.0))
        D1, _ = ext.tc_ts_probe(A1, B1, N, K, b_mn, 0)
        print("   unit probe row0:", D1[0, :16].tolist())
