"""Kernel-author tooling: merged absolute clock64 timeline of mlp_bwd_kernel's epilogue thread 0 and its MMA issuer
(CTA 0) for one steady-state tile pair: who waits for whom, and for how long.
PYTHONPATH=. python tools/mlp_bwd_merged_timeline.py"""
import torch
from directvoxgo_b200 import ext
from directvoxgo_b200.fused_mlp import TensorCoreMLP
from tests.test_gpu_mlp import _make_mlp, _stream

M = 2_400_000
net = _make_mlp(1, 39)
feat, pe, s_ray, counters, cap = _stream(M, 8192, 12, 27, 1, cap_extra=0)
tcb = TensorCoreMLP(net, "cuda", train=True)
pe_pad = tcb.pad_embedding(pe)
rgb = torch.rand(cap, 3, device="cuda")
d_rgb = (torch.randn(cap, 3, device="cuda") / (3 * 8192)).contiguous()
d_feat = torch.zeros(cap, 12, device="cuda")
for _ in range(2):
    tl = ext.mlp_bwd_timeline(feat, s_ray, pe_pad, 27, counters, tcb.params, 128, rgb, d_rgb, 2.0 ** 21, d_feat, tcb.grad_flat)
torch.cuda.synchronize()
e = [x for x in tl[0:64].cpu().tolist() if x]
i = [x for x in tl[64:128].cpu().tolist() if x]
epi_names = ["top"] + [x for ph in ("relu1", "relu2", "mask2", "mask1", "dx") for c in "AB" for x in ("%s %s: batch done seen" % (ph, c), "%s %s: published" % (ph, c))]
# issuer stamps per pair: top, then per batch: acquired, issued  (10 batches)
iss_names = ["top"] + [x for ph in ("L1", "L2", "dW3/dH2", "dW2/dH1", "dW1/dX") for c in "AB" for x in ("%s %s: acquired" % (ph, c), "%s %s: issued" % (ph, c))]
ne, ni = 20, 21
ev = []
pair = 1
es = e[pair * ne:(pair + 1) * ne + 1]
t0 = es[0]
# epilogue stamps: top, (wait done, publish) x ... ; the dx stage has stamps after wait A, after (dx A + wait B)?, see kernel
for k, t in enumerate(es):
    ev.append((t - t0, "EPI  stamp %d" % k))
for k, t in enumerate(i):
    if t0 - 500 <= t <= es[-1] + 500:
        ev.append((t - t0, "ISS  stamp %d (%s)" % (k, iss_names[k % ni] if k % ni < len(iss_names) else "")))
for t, n in sorted(ev):
    print("%7d  %s" % (t, n))
