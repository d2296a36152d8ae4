"""Kernel-author tooling: in-kernel timeline of mlp_fwd_kernel (CTA 0): cycles between phase boundaries."""
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
for _ in range(2):
    tl = ext.mlp_fwd_timeline(feat, s_ray, pe_pad, 27, counters, tc.params, 128, rgb)
torch.cuda.synchronize()
names = ["stage_x", "sync1", "issue L1", "wait L1", "epi1", "sync2", "issue L2", "wait L2", "epi2+L3", "exchange sync",
         "sigmoid+store"]
for who, off in (("thread 0", 0), ("thread 255", 64)):
    t = tl[off:off + 64].cpu().tolist()
    t = [x for x in t if x]
    per_tile = len(names) + 1
    print(who)
    for tile in range(1, min(4, len(t) // per_tile)):
        seg = t[tile * per_tile:(tile + 1) * per_tile + 1]
        d = [b - a for a, b in zip(seg[:-1], seg[1:])]
        print("  tile %d: total %d cyc | " % (tile, seg[-1] - seg[0] if len(seg) > per_tile else sum(d)) +
              ", ".join("%s %d" % (n, x) for n, x in zip(names, d)))

# ---- backward ----
d_rgb = (torch.randn(cap, 3, device="cuda") / (3 * 8192)).contiguous()
tcb = TensorCoreMLP(net, "cuda", train=True)
d_feat = torch.zeros(cap, 12, device="cuda")
for _ in range(2):
    tl = ext.mlp_bwd_timeline(feat, s_ray, pe_pad, 27, counters, tcb.params, 128, rgb, d_rgb, 2.0 ** 21, d_feat, tcb.grad_flat)
torch.cuda.synchronize()
epi = ["stage A+B"] + [x for ph in ("relu1", "relu2", "mask2", "mask1") for x in
                       ("wait A", ph + " A", "wait B", ph + " B")] + ["wait A", "dx A + wait B", "dx B"]
iss = [x for ph in ("L1", "L2", "dW3+dH2", "dW2+db2+dH1", "dW1+dX") for x in
       ("acq A", ph + " A", "acq B", ph + " B")]
for who, off, names in (("epilogue thread 0", 0, epi), ("issuer", 64, iss)):
    t = [x for x in tl[off:off + 64].cpu().tolist() if x]
    n = len(names) + 1
    print(who, "(2nd pair)")
    seg = t[n:2 * n]
    if len(seg) == n:
        d = [b - a for a, b in zip(seg[:-1], seg[1:])]
        print("  pair total %d cyc | " % (seg[-1] - seg[0]) + ", ".join("%s %d" % (a, b) for a, b in zip(names, d)))
