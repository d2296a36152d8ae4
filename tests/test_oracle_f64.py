"""CPU test that PINS the float64 oracle (oracle/dvgo_oracle_f64.c) against the outputs of the reference's own
DOUBLE kernels recorded on a B200 (tests/golden/ref_gpu_ops_f64.npz, `python -m oracle.make_golden_gpu <out> f64`).
Integer / boolean outputs and every float64 output that does not go through exp() / pow() are bit-exact; the
three that do (glibc here, CUDA libdevice there) are within 8 ulp."""
import os

import numpy as np
import torch

from oracle import oracle_f64 as o
from tests import util_f64 as u


def test_f64_oracle_reproduces_reference_double_kernels(golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_gpu_ops_f64.npz"), allow_pickle=False)
    inp = u.inputs_from_golden(g)
    assert inp["rays_o"].dtype == torch.float64
    assert not np.array_equal(g["rays_o"], g["rays_o"].astype(np.float32)), "inputs must not be float32 values"
    inp["exp_d_in"] = torch.from_numpy(g["exp_d"])     # backward checked on the recorded exp_d
    got = u.run_suite(o, o, o, inp, "cpu")
    seen = u.compare(got, {k: g[k] for k in g.files}, ulps=0, libm_ulps=8, label="oracle_f64 vs fixture")
    print({k: v for k, v in seen.items() if v})
    # the float temporaries of the reference's double instantiation are really there: the sampled points are
    # float32 values held in float64 tensors (render_utils_kernel.cu:179-184)
    assert np.array_equal(got["rays_pts"], got["rays_pts"].astype(np.float32).astype(np.float64))
    assert np.array_equal(got["T"], got["T"].astype(np.float32).astype(np.float64))


def test_f64_oracle_agrees_with_f32_oracle_on_float_inputs():
    """On float32-representable inputs the two instantiations differ only by roundings (the double one subtracts and
    divides in double before narrowing to the float temporaries; N_steps is the ceil of a double quotient) -- a
    sanity check that the two restatements describe the same algorithm."""
    from oracle import oracle as o32
    from tests.util import make_rays, ulp_diff
    ro, rd, _, _ = make_rays(300, 9)
    lo, hi = torch.tensor([-1.0, -0.9, -0.8]), torch.tensor([1.0, 0.9, 0.8])
    a = o32.sample_pts_on_rays(ro, rd, lo, hi, 0.2, 6.0, 0.03)
    b = o.sample_pts_on_rays(ro.double(), rd.double(), lo.double(), hi.double(), 0.2, 6.0, 0.03)
    t32 = b[5].numpy().astype(np.float32)
    assert np.array_equal(t32.astype(np.float64), b[5].numpy())               # t_min is a float value in a double tensor
    assert ulp_diff(a[5].numpy(), t32).max() <= 1                             # ... within one float ulp of the f32 clip
    assert (a[4] - b[4]).abs().max() <= 1 and (a[4] == b[4]).float().mean() > 0.98
