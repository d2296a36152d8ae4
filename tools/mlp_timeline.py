"""Kernel-author tooling: in-kernel clock64 timelines of mlp_fwd_kernel / mlp_bwd_kernel (CTA 0): cycles between the
phase boundaries of a few steady-state tiles.   PYTHONPATH=. python tools/mlp_timeline.py"""
import torch
from directvoxgo_b200 import ext
from directvoxgo_b200.fused_mlp import TensorCoreMLP
from tests.test_gpu_mlp import _make_mlp, _stream

M = 2_400_000
net = _make_mlp(1, 39)
feat, pe, s_ray, counters, cap = _stream(M, 8192, 12, 27, 1, cap_extra=0)
tc = TensorCoreMLP(net, "cuda")
pe_pad = tc.pad_embedding(pe)
rgb = torch.zeros(cap, 3, device="cuda")


def show(who, t, names, first_tile=2, n_show=3):
    t = [x for x in t if x]
    n = len(names)
    print(who, "(%d stamps)" % len(t))
    for tile in range(first_tile, first_tile + n_show):
        seg = t[tile * n:(tile + 1) * n + 1]
        if len(seg) < n + 1:
            break
        d = [b - a for a, b in zip(seg[:-1], seg[1:])]
        print("  tile %d: total %5d cyc | " % (tile, seg[-1] - seg[0]) + ", ".join("%s %d" % (a, b) for a, b in zip(names, d)))


for _ in range(2):
    tl = ext.mlp_fwd_timeline(feat, s_ray, pe_pad, 27, counters, tc.params, 128, rgb)
torch.cuda.synchronize()
fwd = ["wait L1 + refill X", "e1->TMEM+sync", "issue L2", "rgb epilogue of the previous tile", "wait L2", "e2->TMEM+sync", "issue L3, L1(next)", "loop"]
t0 = tl[0:64].cpu().tolist()
print("forward prologue (entry -> first loop top): %d cycles" % (t0[1] - t0[0]))
show("forward, thread 0", t0[1:], fwd)
show("forward, thread 255", tl[65:128].cpu().tolist(), fwd)

d_rgb = (torch.randn(cap, 3, device="cuda") / (3 * 8192)).contiguous()
tcb = TensorCoreMLP(net, "cuda", train=True)
d_feat = torch.zeros(cap, 12, device="cuda")
for _ in range(2):
    tl = ext.mlp_bwd_timeline(feat, s_ray, pe_pad, 27, counters, tcb.params, 128, rgb, d_rgb, 2.0 ** 21, d_feat, tcb.grad_flat)
torch.cuda.synchronize()
epi = ["loop top"] + [x for ph in ("relu1", "relu2", "mask2", "mask1") for x in
                       ("wait A", ph + " A", "wait B", ph + " B")] + ["wait A", "dx A + wait B", "dx B"]
t0 = tl[0:64].cpu().tolist()
show("backward, epilogue thread 0", t0, epi, first_tile=1)
t1 = [x for x in tl[64:128].cpu().tolist() if x]
print("backward, issuer: cycles between consecutive batch issues (both contexts):", [b - a for a, b in zip(t1[10:-1], t1[11:])][:30])
