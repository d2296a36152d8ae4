"""Host side of the tensor-core rgbnet (csrc/fused_mlp.cu): flat fp32 master parameters, gradient and
Adam buffers for the 39 -> 128 -> 128 -> 3 MLP of lib/dvgo.py:123-131, and the forward / backward /
optimiser-step calls the FusedTrainer / FusedRenderer make."""
import math

import torch

from . import adam_upd_cuda, ext


class TensorCoreMLP:
    """Hidden widths below 128 (e.g. the LLFF configs' rgbnet_width=64, configs/llff/llff_default.py:30) run on
    the same 128-wide kernels with the weights zero-padded: a padded unit has z = 0, ReLU output 0 and ReLU
    derivative 0, so it contributes nothing forward or backward, its gradients are exactly 0 and Adam leaves it
    at 0 -- the padded network IS the narrow one."""
    WIDTH = 128

    def __init__(self, rgbnet, device, train=False):
        lin = [m for m in rgbnet.modules() if isinstance(m, torch.nn.Linear)]
        if len(lin) != 3 or lin[0].out_features > self.WIDTH or lin[1].in_features != lin[0].out_features or \
                lin[1].out_features != lin[0].out_features or lin[2].out_features != 3:
            raise NotImplementedError("tensor-core rgbnet: depth 3, width <= 128, 3 outputs (the configs' defaults)")
        self.d_in = lin[0].in_features
        self.width = lin[0].out_features
        self.lin = lin
        W, d = self.WIDTH, self.d_in
        self._sizes = [W * d, W, W * W, W, 3 * W, 3]
        self.params = torch.zeros(sum(self._sizes), device=device, dtype=torch.float32)
        self.refresh_from_module()
        self.train = train
        if train:
            self.grad_flat = torch.zeros_like(self.params)
            self.exp_avg = torch.zeros_like(self.params)
            self.exp_avg_sq = torch.zeros_like(self.params)

    def unflatten(self, flat):
        """Views of a flat padded buffer with the shapes of [W1, b1, W2, b2, W3, b3] of the real (narrow) network."""
        W, d, w = self.WIDTH, self.d_in, self.width
        o = [0]
        for n in self._sizes:
            o.append(o[-1] + n)
        return [flat[o[0]:o[1]].view(W, d)[:w], flat[o[1]:o[2]][:w], flat[o[2]:o[3]].view(W, W)[:w, :w],
                flat[o[3]:o[4]][:w], flat[o[4]:o[5]].view(3, W)[:, :w], flat[o[5]:o[6]]]

    @torch.no_grad()
    def refresh_from_module(self):
        for dst, src in zip(self.unflatten(self.params), [t for l in self.lin for t in (l.weight, l.bias)]):
            dst.copy_(src.detach())

    def pad_embedding(self, pe):
        """[N,P] view embedding -> [N,P_pad] table: column P = 1 (carries b1 through the first GEMM), then
        zeros up to a multiple of 4 floats (16-byte loads in the kernel)."""
        n, P = pe.shape
        stride = (P + 1 + 3) // 4 * 4
        out = torch.zeros(n, stride, dtype=pe.dtype, device=pe.device)
        out[:, :P] = pe
        out[:, P] = 1.0
        self._P = P
        return out

    def embed(self, viewdirs, viewfreq, C=None):
        """The padded view-embedding table of `pad_embedding(view_embedding(...))` in one kernel.  With C (the k0
        channel count) also the rays' share of the X~ rows as fp16 ([N, K1]; what k0_gather_tiles copies): (pe, rows16)."""
        P = 3 + 6 * int(viewfreq.numel())
        self._P = P
        out = ext.view_embedding(viewdirs.contiguous(), viewfreq, (P + 1 + 3) // 4 * 4, -1 if C is None else C)
        return out[0] if C is None else (out[0], out[1])

    def pack(self, C, pe_stride):
        """fp32 master weights -> the fp16 operand tiles the kernels bulk-copy (one tiny kernel; called by forward(),
        reused by backward() of the same step: the masters only change in adam_step)."""
        key = (C, pe_stride)
        if getattr(self, "_wpack_key", None) != key:
            self._wpack = torch.empty(ext.mlp_wpack_bytes(C, pe_stride), dtype=torch.uint8, device=self.params.device)
            self._wpack_key = key
        ext.mlp_pack(self.params, C, self.d_in - C, pe_stride, self.WIDTH, self._wpack)
        return self._wpack

    def grad_scale(self, n_global):
        """Power of two applied to the backward's fp16 operands: d_rgb <= ~2/(3 n_global), so 256 n_global brings the
        largest of them to O(100); removed exactly in the fp32 epilogues."""
        return 2.0 ** math.floor(math.log2(256.0 * n_global))

    def forward(self, feat, s_ray, pe_pad, counters, rgb):
        """From fp32 streams (the tiles are built into a temporary first)."""
        wp = self.pack(feat.shape[1], pe_pad.shape[1])
        ext.mlp_fwd(feat, s_ray, pe_pad, self.d_in - feat.shape[1], counters, self.params, self.WIDTH, rgb, wp)

    def backward(self, feat, s_ray, pe_pad, counters, rgb, d_rgb, d_feat, n_global):
        """From fp32 streams.  Accumulates weight gradients into self.grad_flat (zeroed here) and writes d_feat."""
        ext.zero_(self.grad_flat)
        wp = self._wpack if getattr(self, "_wpack_key", None) == (feat.shape[1], pe_pad.shape[1]) else \
            self.pack(feat.shape[1], pe_pad.shape[1])
        ext.mlp_bwd(feat, s_ray, pe_pad, self.d_in - feat.shape[1], counters, self.params, self.WIDTH, rgb, d_rgb,
                    self.grad_scale(n_global), d_feat, self.grad_flat, wp)

    # the fused step's calls: survivor tiles written by their producers (k0_gather_tiles / sample_grad)
    def forward_tiles(self, xt, C, pe_stride, counters, cap, rgb, wp=None):
        """wp: weight tiles packed earlier by pack() (a renderer packs once per view); default: pack now."""
        if wp is None:
            wp = self.pack(C, pe_stride)
        ext.mlp_fwd_tiles(xt, C, self.d_in - C, pe_stride, counters, cap, wp, rgb)

    def forward_gather(self, scene, k0_cl, s_pos, pe16, C, pe_stride, counters, cap, rgb, xt=None, wp=None):
        """k0 gather + forward in one kernel (mlp_fwd_gather_kernel): the X~ tiles are built in shared memory by
        producer warps; xt (training) receives a copy of every tile for backward_tiles()."""
        if wp is None:
            wp = self.pack(C, pe_stride)
        ext.mlp_fwd_gather(scene, k0_cl, s_pos, pe16, self.d_in - C, pe_stride, counters, cap, wp, rgb, xt)

    def backward_tiles(self, xt, dzt, C, pe_stride, counters, cap, d_feat, n_global):
        ext.zero_(self.grad_flat)
        wp = self._wpack if getattr(self, "_wpack_key", None) == (C, pe_stride) else self.pack(C, pe_stride)
        ext.mlp_bwd_tiles(xt, dzt, C, self.d_in - C, pe_stride, counters, cap, wp, self.grad_scale(n_global), d_feat,
                          self.grad_flat)

    def adam_step(self, step, beta1, beta2, lr, eps):
        adam_upd_cuda.adam_upd(self.params, self.grad_flat, self.exp_avg, self.exp_avg_sq, step, beta1, beta2, lr, eps)

    @torch.no_grad()
    def sync_to_module(self, rgbnet=None):
        for dst, src in zip([t for l in self.lin for t in (l.weight, l.bias)], self.unflatten(self.params)):
            dst.data.copy_(src)
