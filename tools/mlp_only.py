"""Launch mlp_fwd_kernel / mlp_bwd_kernel on a 2.4 M-survivor stream (for ncu).  PYTHONPATH=. python tools/mlp_only.py"""
import torch
from directvoxgo_b200 import ext
from directvoxgo_b200.fused_mlp import TensorCoreMLP
from tests.test_gpu_mlp import _make_mlp, _stream

M = 2_400_000
net = _make_mlp(1, 39)
feat, pe, s_ray, counters, cap = _stream(M, 8192, 12, 27, 1, cap_extra=0)
tc = TensorCoreMLP(net, "cuda", train=True)
pe_pad = tc.pad_embedding(pe)
rgb = torch.zeros(cap, 3, device="cuda")
d_rgb = (torch.randn(cap, 3, device="cuda") / (3 * 8192)).contiguous()
d_feat = torch.zeros(cap, 12, device="cuda")
# the kernels' inputs: survivor tiles, built once here (the fused step's producers write them directly)
xt = torch.zeros(ext.mlp_xtile_bytes(cap, 12, pe_pad.shape[1]), dtype=torch.uint8, device="cuda")
dzt = torch.zeros(ext.mlp_dztile_bytes(cap), dtype=torch.uint8, device="cuda")
ext.mlp_pack_x(feat, s_ray, pe_pad, 27, counters, xt)
tc.forward_tiles(xt, 12, pe_pad.shape[1], counters, cap, rgb)
ext.mlp_pack_dz(rgb, d_rgb, tc.grad_scale(8192), counters, dzt)
for _ in range(3):
    tc.forward_tiles(xt, 12, pe_pad.shape[1], counters, cap, rgb)
    tc.backward_tiles(xt, dzt, 12, pe_pad.shape[1], counters, cap, d_feat, 8192)
torch.cuda.synchronize()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record()
for _ in range(5):
    tc.forward_tiles(xt, 12, pe_pad.shape[1], counters, cap, rgb)
e1.record()
for _ in range(5):
    tc.backward_tiles(xt, dzt, 12, pe_pad.shape[1], counters, cap, d_feat, 8192)
e2.record()
torch.cuda.synchronize()
print("mlp_fwd %.4f ms, mlp_bwd %.4f ms (M = %d; includes the weight pack / gradient zeroing launches)" % (
    e0.elapsed_time(e1) / 5, e1.elapsed_time(e2) / 5, M))
