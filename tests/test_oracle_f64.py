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


def test_f64_oracle_invariants():
    """Size-independent properties of the restated double instantiation (render_utils_kernel.cu:431-505): the weights
    of a ray telescope to 1 - alphainv_last up to the float rounding of T_cum; T is non-increasing inside the kept
    range; samples past the early stop keep the fills (T = 1, weight = 0); the backward equals the analytic gradient
    of sum_i gw_i w_i + gl * alphainv_last to float precision on rays without an early stop."""
    from tests.util import sorted_ray_ids
    n_rays, n_pts = 200, 30000
    rid = sorted_ray_ids(n_rays, n_pts, 21)
    g = torch.Generator().manual_seed(22)
    alpha = (torch.rand(n_pts, generator=g, dtype=torch.float64) ** 3 * 0.5)
    w, T, last, i_s, i_e = o.alpha2weight(alpha, rid, n_rays)
    seg = torch.bincount(rid, minlength=n_rays)
    wsum = torch.zeros(n_rays, dtype=torch.float64).index_add_(0, rid, w)
    assert float((wsum + last - 1).abs().max()) < 5e-5          # float T_cum, up to ~600 samples per ray
    for r in range(n_rays):
        s, e, end = int(i_s[r]), int(i_e[r]), int(i_s[r]) + int(seg[r])
        if seg[r] == 0:
            assert s == 0 and e == 0 and float(last[r]) == 1.0
            continue
        assert torch.all(T[s:e][1:] <= T[s:e][:-1]) and float(T[s]) == 1.0
        assert torch.all(T[e:end] == 1.0) and torch.all(w[e:end] == 0.0)
        assert e == end or float(last[r]) < 1e-3
    # analytic gradient in double on a stop-free problem (small alphas): d/d alpha_k of sum gw_i T_i alpha_i + gl T_end
    n_rays, per = 50, 20
    rid = torch.arange(n_rays).repeat_interleave(per)
    alpha = (0.01 + 0.1 * torch.rand(n_rays * per, generator=g, dtype=torch.float64)).requires_grad_(True)
    gw = torch.randn(n_rays * per, generator=g, dtype=torch.float64)
    gl = torch.randn(n_rays, generator=g, dtype=torch.float64)
    f = (1 - alpha + 1e-10).view(n_rays, per)
    Tt = torch.cat([torch.ones(n_rays, 1, dtype=torch.float64), torch.cumprod(f, 1)], 1)
    loss = (gw.view(n_rays, per) * Tt[:, :-1] * alpha.view(n_rays, per)).sum() + (gl * Tt[:, -1]).sum()
    loss.backward()
    a = alpha.detach()
    w, T, last, i_s, i_e = o.alpha2weight(a, rid, n_rays)
    got = o.alpha2weight_backward(a, w, T, last, i_s, i_e, n_rays, gw, gl)
    assert float((got - alpha.grad).abs().max()) < 2e-6 * float(alpha.grad.abs().max())
