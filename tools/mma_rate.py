"""tcgen05.mma issue/throughput probe: cycles per M=128 MMA (K=16, fp16) for different N, operand orientations and
shared-memory layouts.  PYTHONPATH=. python tools/mma_rate.py"""
import torch
from directvoxgo_b200 import ext

def run(name, ctas, N, ksteps, a_mn, b_mn, a, b, layout=0, n_accum=1, reps=200):
    out = ext.tc_rate(ctas, N, ksteps, reps, a_mn, b_mn, *a, *b, layout, n_accum)
    torch.cuda.synchronize()
    c = out.double()
    print("%-58s N=%3d ctas=%3d: %7.1f cycles/MMA (max over CTAs %7.1f)" % (name, N, ctas, float(c.mean()) / (reps * ksteps),
                                                                           float(c.max()) / (reps * ksteps)))

gs = lambda cols: cols // 8 * 128
for ctas in (1, 148):
    for N in (16, 48, 64, 128, 144, 256):
        # no swizzle, both K-major, K = 128: LBO 128, SBO gs(128), k-step 256
        run("no-swizzle A K-major, B K-major", ctas, N, 8, False, False, (128, gs(128), 256), (128, gs(128), 256))
    for N in (48, 128):
        run("no-swizzle A K-major, B MN-major", ctas, N, 8, False, True, (128, gs(128), 256), (gs(N if N >= 64 else 64), 128, 2 * gs(N if N >= 64 else 64)))
        run("no-swizzle A MN-major, B MN-major", ctas, N, 8, True, True, (gs(128), 128, 2 * gs(128)), (gs(128), 128, 2 * gs(128)))
    for N in (128, 256):
        # 128-byte swizzle, K-major: rows of 64 halves (128 B), 8-row atoms of 1024 B; K = 64 per atom -> 4 k-steps of 32 B
        run("SWIZZLE_128B A K-major, B K-major (K=64)", ctas, N, 4, False, False, (16, 1024, 32), (16, 1024, 32), layout=2)
        run("no-swizzle, 2 accumulators alternating", ctas, N if N <= 128 else 128, 8, False, False, (128, gs(128), 256), (128, gs(128), 256), n_accum=2)
