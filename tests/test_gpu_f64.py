"""GPU parity tests of the float64 instantiation (include/dvgo_b200_f64.h, directvoxgo_b200/csrc/f64_ops.cu): the
reference dispatches its three extensions over AT_DISPATCH_FLOATING_TYPES (lib/cuda/render_utils_kernel.cu:86 ...),
so the same Python calls accept float64 tensors.  Checked against
  * the reference's OWN double kernels, live on this GPU (oracle/_ref) -- same inputs, full op surface;
  * the committed fixture recorded from those kernels (tests/golden/ref_gpu_ops_f64.npz);
  * the CPU oracle (oracle/dvgo_oracle_f64.c);
  * torch.autograd.gradcheck of the autograd glue, which is what a double instantiation is for.
Stated tolerance: integer / boolean outputs bit-exact; float64 outputs <= 2 ulp (double) against the reference
kernels (observed maxima are written to gpurun_out/f64_ulps.json), exp/pow outputs <= 8 ulp against the CPU libm."""
import json
import os

import numpy as np
import pytest
import torch

from tests import util_f64 as u

pytestmark = pytest.mark.gpu
DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def pkg():
    import directvoxgo_b200 as p
    return p


def _ours(pkg, inp):
    return u.run_suite(pkg.render_utils_cuda, pkg.total_variation_cuda, pkg.adam_upd_cuda, inp, DEV)


def _note(name, seen):
    path = os.path.join(ROOT, "gpurun_out", "f64_ulps.json")
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        old = json.load(open(path)) if os.path.exists(path) else {}
        old[name] = seen
        json.dump(old, open(path, "w"), indent=1, sort_keys=True)
    except OSError:
        pass
    print(name, "max ulp(double) per output:", {k: v for k, v in seen.items() if v})


@pytest.mark.parametrize("seed", [3, 17])
def test_f64_ops_vs_reference_double_kernels(pkg, ref_gpu, seed):
    inp = u.make_inputs(seed)
    want = u.run_suite(ref_gpu.render_utils_cuda, ref_gpu.total_variation_cuda, ref_gpu.adam_upd_cuda, inp, DEV)
    got = _ours(pkg, inp)
    assert want["rays_pts"].dtype == np.float64 and len(want["ray_id"]) > 20000
    seg = np.bincount(inp["a2w_ray_id"].numpy(), minlength=int(inp["a2w_n_rays"]))
    assert ((want["i_end"] - want["i_start"]) < seg).sum() > 10, "inputs must exercise the early stop of alpha2weight"
    _note("vs_reference_kernels_seed%d" % seed, u.compare(got, want, ulps=2, label="ours vs oracle/_ref"))


def test_f64_ops_vs_committed_reference_fixture(pkg, golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_gpu_ops_f64.npz"), allow_pickle=False)
    got = _ours(pkg, u.inputs_from_golden(g))
    _note("vs_golden_fixture", u.compare(got, {k: g[k] for k in g.files}, ulps=2, label="ours vs fixture"))


def test_f64_ops_vs_cpu_oracle(pkg):
    from oracle import oracle_f64 as o
    inp = u.make_inputs(29, n_rays=256, n_pts=20000, a2w_rays=300, n_adam=20011)
    got = _ours(pkg, inp)
    inp_cpu = dict(inp, exp_d_in=torch.from_numpy(got["exp_d"]))   # backward checked on the same exp_d
    want = u.run_suite(o, o, o, inp_cpu, "cpu")
    _note("vs_cpu_oracle", u.compare(got, want, ulps=0, libm_ulps=8, label="ours vs oracle_f64"))


def test_f64_gradcheck_of_the_autograd_glue(pkg):
    from directvoxgo_b200.ops import Alphas2Weights, Raw2Alpha
    g = torch.Generator().manual_seed(5)
    d = (torch.randn(64, generator=g, dtype=torch.float64) * 2).to(DEV).requires_grad_(True)
    # raw2alpha is double throughout: the default finite-difference step works
    assert torch.autograd.gradcheck(lambda x: Raw2Alpha.apply(x, -4.0, 0.5), (d,), eps=1e-6, atol=1e-7, rtol=1e-5)
    # alpha2weight keeps its running transmittance in float (render_utils_kernel.cu:447,522), so its outputs carry
    # ~1e-7 of rounding noise; the weights are multilinear in alpha, so a wide central difference has no truncation error
    n_rays = 6
    ray_id = torch.arange(n_rays).repeat_interleave(12).to(DEV)
    a = (0.02 + 0.18 * torch.rand(ray_id.numel(), generator=g, dtype=torch.float64)).to(DEV).requires_grad_(True)
    assert torch.autograd.gradcheck(lambda x: Alphas2Weights.apply(x, ray_id, n_rays), (a,), eps=1e-2, atol=2e-4,
                                    rtol=1e-3)


def test_f64_dtype_errors_and_empty_inputs(pkg):
    ru, tv, ad = pkg.render_utils_cuda, pkg.total_variation_cuda, pkg.adam_upd_cuda
    d64 = torch.zeros(8, dtype=torch.float64, device=DEV)
    d32 = torch.zeros(8, dtype=torch.float32, device=DEV)
    with pytest.raises(RuntimeError, match="float32 or float64"):
        ru.raw2alpha(d32.half(), 0.0, 1.0)
    with pytest.raises(RuntimeError, match="must have the dtype of"):
        ru.raw2alpha_backward(d64, d32, 1.0)
    with pytest.raises(RuntimeError, match="must have the dtype of"):
        ad.adam_upd(d64, d64, d32, d64, 1, 0.9, 0.99, 0.1, 1e-8)
    with pytest.raises(RuntimeError, match="must have the dtype of"):
        tv.total_variation_add_grad(d64.view(1, 1, 2, 2, 2), d32.view(1, 1, 2, 2, 2), 1.0, 1.0, 1.0, True)
    # empty inputs: shapes and dtypes of the reference's early returns (render_utils_kernel.cu:377-379,483-485)
    e = torch.zeros(0, dtype=torch.float64, device=DEV)
    ex, al = ru.raw2alpha(e, 0.0, 1.0)
    assert ex.dtype == torch.float64 and ex.numel() == 0 and al.numel() == 0
    w, T, last, i_s, i_e = ru.alpha2weight(e, torch.zeros(0, dtype=torch.int64, device=DEV), 5)
    assert last.dtype == torch.float64 and torch.equal(last, torch.ones(5, dtype=torch.float64, device=DEV))
    assert torch.equal(i_s, torch.zeros(5, dtype=torch.int64, device=DEV)) and torch.equal(i_e, i_s)
    pts, mask, rid, sid, ns, tmin, tmax = ru.sample_pts_on_rays(torch.zeros(0, 3, dtype=torch.float64, device=DEV),
                                                                torch.zeros(0, 3, dtype=torch.float64, device=DEV),
                                                                d64[:3].contiguous(), d64[:3].contiguous() + 1, 0.1, 1.0, 0.1)
    assert pts.shape == (0, 3) and pts.dtype == torch.float64 and rid.numel() == 0
