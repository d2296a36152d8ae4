"""Autograd glue over the native ops -- the Python half of the drop-in boundary.

Mirrors the reference's L2 layer: `Raw2Alpha`, `Alphas2Weights` (lib/dvgo.py:618-660) plus the two
third-party ops the reference pulls in by name, re-implemented on our own kernels:
`grid_sample_trilinear` (F.grid_sample call site lib/dvgo.py:312-328) and `segment_coo`
(torch_scatter, call sites lib/dvgo.py:554-558, 571-575).
"""
import torch

from . import ext, render_utils_cuda


class Raw2Alpha(torch.autograd.Function):
    """alpha = 1 - (1 + exp(density + shift)) ** (-interval)   (lib/dvgo.py:618-642)."""

    @staticmethod
    def forward(ctx, density, shift, interval):
        exp_d, alpha = render_utils_cuda.raw2alpha(density, shift, interval)
        if density.requires_grad:
            ctx.save_for_backward(exp_d)
            ctx.interval = interval
        return alpha

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_alpha):
        (exp_d,) = ctx.saved_tensors
        g = render_utils_cuda.raw2alpha_backward(exp_d, grad_alpha.contiguous(), ctx.interval)
        return g, None, None


class Alphas2Weights(torch.autograd.Function):
    """Per-ray exclusive transmittance product (lib/dvgo.py:644-660)."""

    @staticmethod
    def forward(ctx, alpha, ray_id, N):
        weights, T, alphainv_last, i_start, i_end = render_utils_cuda.alpha2weight(alpha, ray_id, N)
        if alpha.requires_grad:
            ctx.save_for_backward(alpha, weights, T, alphainv_last, i_start, i_end)
            ctx.n_rays = N
        return weights, alphainv_last

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_weights, grad_last):
        alpha, weights, T, alphainv_last, i_start, i_end = ctx.saved_tensors
        g = render_utils_cuda.alpha2weight_backward(
            alpha, weights, T, alphainv_last, i_start, i_end, ctx.n_rays,
            grad_weights.contiguous(), grad_last.contiguous())
        return g, None, None


class _GridSampleTrilinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, grid, xyz, xyz_min, xyz_max):
        out = ext.grid_sample_3d(grid, xyz, xyz_min, xyz_max)
        if grid.requires_grad:
            ctx.save_for_backward(xyz, xyz_min, xyz_max)
            ctx.grid_shape = grid.shape
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        xyz, xyz_min, xyz_max = ctx.saved_tensors
        grad_grid = torch.zeros(ctx.grid_shape, dtype=grad_out.dtype, device=grad_out.device)
        ext.grid_sample_3d_backward(grad_out.contiguous(), xyz, xyz_min, xyz_max, grad_grid)
        return grad_grid, None, None, None


class _GridSamplePlane(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plane, xyz, xyz_min, xyz_max, axis_w, axis_h):
        out = ext.grid_sample_2d(plane, xyz, xyz_min, xyz_max, axis_w, axis_h)
        if plane.requires_grad:
            ctx.save_for_backward(xyz, xyz_min, xyz_max)
            ctx.meta = (plane.shape, axis_w, axis_h)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        xyz, xyz_min, xyz_max = ctx.saved_tensors
        shape, axis_w, axis_h = ctx.meta
        grad_plane = torch.zeros(shape, dtype=grad_out.dtype, device=grad_out.device)
        ext.grid_sample_2d_backward(grad_out.contiguous(), xyz, xyz_min, xyz_max, axis_w, axis_h, grad_plane)
        return grad_plane, None, None, None, None, None


class _GridSampleNorm(torch.autograd.Function):
    """F.grid_sample(input, grid, 'bilinear', 'zeros', align_corners=True) for batch 1 with the literal arguments."""

    @staticmethod
    def forward(ctx, inp, pts):
        out = ext.grid_sample_norm(inp, pts)
        if inp.requires_grad:
            ctx.save_for_backward(pts)
            ctx.in_shape = inp.shape
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        (pts,) = ctx.saved_tensors
        grad_in = torch.zeros(ctx.in_shape, dtype=grad_out.dtype, device=grad_out.device)
        ext.grid_sample_norm_backward(grad_out.contiguous(), pts, grad_in)
        return grad_in, None


def make_grid_sample(real):
    """A stand-in for torch.nn.functional.grid_sample that serves the call shapes of the unmodified reference
    (lib/dvgo.py:321, lib/dmpigo.py:169, lib/tri_dvgo.py:462-464,618: batch 1, CUDA fp32, bilinear, zero padding,
    align_corners=True, coordinates without grad) from our kernels and hands everything else to `real`."""

    def grid_sample(input, grid, mode="bilinear", padding_mode="zeros", align_corners=None):
        nd = input.dim() - 2
        if (input.is_cuda and input.dtype == torch.float32 and grid.dtype == torch.float32 and mode == "bilinear"
                and padding_mode == "zeros" and align_corners is True and nd in (2, 3) and input.shape[0] == 1
                and grid.shape[0] == 1 and grid.shape[-1] == nd and not grid.requires_grad and input.is_contiguous()):
            pts = grid.reshape(-1, nd).contiguous()
            out = _GridSampleNorm.apply(input, pts)                       # [P, C]
            return out.T.reshape(1, input.shape[1], *grid.shape[1:-1])   # the [1,C,...] ATen returns (a view)
        return real(input, grid, mode=mode, padding_mode=padding_mode, align_corners=align_corners)

    grid_sample._dvgo_real = real
    return grid_sample


# which world axis indexes the W / H dimension of each plane: the reference samples plane 'ab' with
# ind_norm[..., [i, j]] where ind_norm = (z_n, y_n, x_n) (lib/tri_dvgo.py:460-464)
TRIPLANE_AXES = {"xy": (2, 1), "yz": (1, 0), "zx": (0, 2)}


def grid_sample_triplane(grids, xyz, xyz_min, xyz_max, aggregation="concat"):
    """Tri-plane feature lookup, the reference's `grid_sampler2D` (lib/tri_dvgo.py:456-471):
    grids = {'xy','yz','zx'} of [1,C,H,W] planes, xyz [P,3] world coords -> [P,3C] ('concat') or [P,C] ('sum').
    One kernel per plane: ind_norm arithmetic + ATen's bilinear blend (align_corners=True, zero padding)."""
    x = xyz.reshape(-1, 3).contiguous()
    feats = [_GridSamplePlane.apply(grids[k], x, xyz_min, xyz_max, *TRIPLANE_AXES[k]) for k in ("xy", "yz", "zx")]
    if aggregation == "concat":
        return torch.cat(feats, dim=-1)
    if aggregation == "sum":
        return feats[0] + feats[1] + feats[2]
    raise NotImplementedError(aggregation)


def grid_sample_trilinear(grid, xyz, xyz_min, xyz_max):
    """DenseGrid sampling: grid [1,C,X,Y,Z], xyz [...,3] world coords -> [...,C] ([...] if C==1).

    Equivalent to the reference's `grid_sampler` (lib/dvgo.py:312-328): the ind_norm arithmetic,
    ATen's align_corners=True un-normalisation, the 8-corner blend with zero padding and the
    [C,P]->[P,C] transpose are one kernel; backward scatters into a fresh zero grid-gradient.
    """
    shape = xyz.shape[:-1]
    out = _GridSampleTrilinear.apply(grid, xyz.reshape(-1, 3).contiguous(), xyz_min, xyz_max)
    out = out.reshape(*shape, grid.shape[1])
    return out.squeeze(-1) if grid.shape[1] == 1 else out


class _SegmentCooSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, index, out):
        ext.segment_coo_sum(src.contiguous(), index, out)
        ctx.save_for_backward(index)
        ctx.mark_dirty(out)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        (index,) = ctx.saved_tensors
        return ext.gather_rows(grad_out.contiguous(), index), None, None


def segment_coo(src, index, out=None, dim_size=None, reduce="sum"):
    """torch_scatter.segment_coo for the one mode DirectVoxGO uses: sorted `index` along dim 0,
    reduce='sum', accumulating into `out` (lib/dvgo.py:554-558)."""
    if reduce not in ("sum", "add"):
        raise NotImplementedError("segment_coo shim: only reduce='sum' is implemented (DirectVoxGO's use)")
    if out is None:
        if dim_size is None:
            dim_size = int(index[-1].item()) + 1 if index.numel() else 0
        out = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    else:
        # the reference passes torch.zeros(...) created under the CUDA default tensor type
        out = out.to(device=src.device, dtype=src.dtype)
    return _SegmentCooSum.apply(src, index, out)


def scatter_add(src, index, dim=0, out=None, dim_size=None):
    """torch_scatter.scatter_add (imported but unused by lib/dmpigo.py:11); unsorted index -> ATen."""
    if out is None:
        if dim_size is None:
            dim_size = int(index.max().item()) + 1 if index.numel() else 0
        shape = list(src.shape)
        shape[dim] = dim_size
        out = torch.zeros(shape, dtype=src.dtype, device=src.device)
    return out.index_add_(dim, index, src)
