"""GPU tests of the tensor-core (tcgen05, TF32) rgbnet kernels."""
import numpy as np
import pytest
import torch

from tests.util import rel_to_max, to_np

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _h(x):
    """Round to FP16 and back (what the kernels do when staging GEMM operands)."""
    return x.half().float()


@pytest.mark.parametrize("N,K,a_mn,b_mn", [
    (128, 16, 0, 0), (128, 48, 0, 0), (128, 128, 0, 0),    # forward layers: A, B K-major
    (16, 128, 0, 0), (16, 16, 0, 1), (128, 16, 0, 1),      # K-major A, MN-major B  (dH = dZ * W)
    (128, 128, 0, 1), (16, 128, 0, 1),
    (128, 128, 1, 1), (48, 128, 1, 1), (16, 128, 1, 1),    # MN-major A and B       (dW = dZ^T * H)
    (64, 64, 1, 0)])
def test_tcgen05_descriptor_orientations(N, K, a_mn, b_mn):
    from directvoxgo_b200 import ext
    g = torch.Generator().manual_seed(N * 1000 + K * 10 + a_mn * 2 + b_mn)
    A = torch.randn(128, K, generator=g).to(DEV)      # logical A [M=128, K]
    B = torch.randn(N, K, generator=g).to(DEV)        # logical B [N, K]
    a_in = A.t().contiguous() if a_mn else A.contiguous()
    b_in = B.t().contiguous() if b_mn else B.contiguous()
    D = ext.tc_selftest(a_in, b_in, N, K, bool(a_mn), bool(b_mn))
    torch.cuda.synchronize()
    ref = (_h(A).double() @ _h(B).double().t()).float()
    # fp16-rounded inputs, fp32 accumulation: agreement to fp32 summation error
    err = rel_to_max(D, ref)
    assert err < 2e-6, (N, K, a_mn, b_mn, err)
