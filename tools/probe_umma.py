"""Debug aid: reveal the shared-memory address map the tensor core applies to a B operand for a given
(major, LBO, SBO).  Run on the GPU box: python tools/probe_umma.py"""
import torch
from directvoxgo_b200 import ext

dev = "cuda"
K = 16
A = torch.zeros(128, K, device=dev)
for k in range(K):
    A[k, k] = 1.0                       # D[k][n] = B_hw(n, k)
ramp = torch.arange(2048, dtype=torch.float32, device=dev)
for b_mn in (0, 1):
    for (lbo, sbo) in ((128, 512), (512, 128), (128, 256), (256, 128), (128, 1024), (1024, 128)):
        for N in (16, 32):
            D = ext.tc_probe(A, ramp, N, K, bool(b_mn), lbo, sbo, 0)
            torch.cuda.synchronize()
            w = D[:K, :N].t().long()    # [n][k] -> word index read
            print("b_mn=%d lbo=%4d sbo=%4d N=%d" % (b_mn, lbo, sbo, N))
            for n in range(N):
                print("   n=%2d:" % n, " ".join("%4d" % int(x) for x in w[n]))
