import torch, numpy as np
from tests.test_gpu_mlp import _make_mlp, _stream
from directvoxgo_b200.fused_mlp import TensorCoreMLP
DEV = "cuda"
M, n_global, C, P = 300, 8192, 12, 27
net = _make_mlp(M + 5, C + P)
feat, pe, s_ray, counters, cap = _stream(M, 300, C, P, M + 1)
g = torch.Generator().manual_seed(M)
d_rgb = (torch.randn(cap, 3, generator=g) * (1.0 / (3 * n_global))).to(DEV)
tc = TensorCoreMLP(net, DEV, train=True)
rgb = torch.zeros(cap, 3, device=DEV)
d_feat = torch.full((cap, C), 3.0, device=DEV)
tc.forward(feat, s_ray, pe, counters, rgb)
tc.backward(feat, s_ray, pe, counters, rgb, d_rgb, d_feat, n_global)
torch.cuda.synchronize()
x = torch.cat([feat[:M], pe[s_ray[:M].long()]], -1).requires_grad_()
out = torch.sigmoid(net(x)); out.backward(d_rgb[:M])
ref = x.grad[:, :C]
err = (d_feat[:M] - ref).abs()
mx = ref.abs().max()
print("max ref", float(mx), "max err", float(err.max()))
rowerr = (err.max(dim=1).values / mx).cpu().numpy()
print("rows with err>1e-2:", np.nonzero(rowerr > 1e-2)[0][:60], "count", int((rowerr > 1e-2).sum()))
colerr = (err.max(dim=0).values / mx).cpu().numpy()
print("col err", np.round(colerr, 4))
lin = [m for m in net.modules() if isinstance(m, torch.nn.Linear)]
got = tc.unflatten(tc.grad_flat)
for name, gt, p in zip(["W1", "b1", "W2", "b2", "W3", "b3"], got, [t for l in lin for t in (l.weight, l.bias)]):
    e = (gt - p.grad).abs().max() / p.grad.abs().max()
    print(name, "rel err", float(e))
# ratio analysis on a bad row
bad = int(np.argmax(rowerr))
print("bad row", bad, "got", d_feat[bad].cpu().numpy(), "\nref", ref[bad].cpu().numpy())
