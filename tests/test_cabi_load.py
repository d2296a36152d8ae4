"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/*.h declares, and the torch binding refuses non-CUDA tensors loudly (no CPU fallback)."""
import ctypes
import glob
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = []
    for h in sorted(glob.glob(os.path.join(ROOT, "include", "*.h"))):
        src = open(h).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names += re.findall(r"\b(dvgo_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_header_declares_the_hot_path():
    names = declared_symbols()
    for must in ["dvgo_infer_t_minmax", "dvgo_infer_n_samples", "dvgo_infer_ray_start_dir",
                 "dvgo_sample_pts_count", "dvgo_sample_pts_fill", "dvgo_sample_ndc_pts_on_rays",
                 "dvgo_maskcache_lookup", "dvgo_raw2alpha", "dvgo_raw2alpha_backward",
                 "dvgo_alpha2weight", "dvgo_alpha2weight_backward", "dvgo_grid_sample_3d",
                 "dvgo_grid_sample_3d_backward", "dvgo_segment_coo_sum",
                 "dvgo_total_variation_add_grad", "dvgo_adam_upd", "dvgo_masked_adam_upd",
                 "dvgo_adam_upd_with_perlr"]:
        assert must in names


def test_library_exports_every_declared_symbol():
    import directvoxgo_b200 as pkg
    lib = ctypes.CDLL(pkg.LIB_PATH)
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, "declared in include/*.h but not exported: %s" % missing
    lib.dvgo_abi_version.restype = ctypes.c_int
    assert lib.dvgo_abi_version() == 2      # DVGO_ABI_VERSION of include/dvgo_b200.h
    lib.dvgo_build_arch.restype = ctypes.c_char_p
    assert lib.dvgo_build_arch() == b"sm_100a"


def test_invalid_arguments_return_einval_without_touching_the_gpu():
    import directvoxgo_b200 as pkg
    lib = ctypes.CDLL(pkg.LIB_PATH)
    assert lib.dvgo_infer_t_minmax(None, None, None, None, ctypes.c_float(0), ctypes.c_float(1),
                                   ctypes.c_int(-1), None, None, None) == -1
    assert lib.dvgo_infer_t_minmax(None, None, None, None, ctypes.c_float(0), ctypes.c_float(1),
                                   ctypes.c_int(4), None, None, None) == -1
    # empty inputs short-circuit to success (reference: render_utils_kernel.cu:333-335,377-379)
    assert lib.dvgo_raw2alpha(None, ctypes.c_float(0), ctypes.c_float(1), ctypes.c_int64(0),
                              None, None, None) == 0
    assert lib.dvgo_maskcache_lookup(None, None, None, None, 1, 1, 1, ctypes.c_int64(0), None, None) == 0


def test_binding_surface_matches_reference_pybind_tables():
    import directvoxgo_b200 as pkg
    # lib/cuda/render_utils.cpp:144-155
    for n in ["infer_t_minmax", "infer_n_samples", "infer_ray_start_dir", "sample_pts_on_rays",
              "sample_ndc_pts_on_rays", "maskcache_lookup", "raw2alpha", "raw2alpha_backward",
              "alpha2weight", "alpha2weight_backward"]:
        assert callable(getattr(pkg.render_utils_cuda, n))
    assert callable(pkg.total_variation_cuda.total_variation_add_grad)  # total_variation.cpp:22-24
    for n in ["adam_upd", "masked_adam_upd", "adam_upd_with_perlr"]:     # adam_upd.cpp:79-86
        assert callable(getattr(pkg.adam_upd_cuda, n))


def test_no_cpu_fallback():
    import directvoxgo_b200 as pkg
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        pkg.render_utils_cuda.raw2alpha(torch.zeros(4), 0.0, 1.0)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        pkg.adam_upd_cuda.adam_upd(torch.zeros(4), torch.zeros(4), torch.zeros(4), torch.zeros(4),
                                   1, 0.9, 0.99, 0.1, 1e-8)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        pkg.total_variation_cuda.total_variation_add_grad(torch.zeros(1, 1, 2, 2, 2), torch.zeros(1, 1, 2, 2, 2),
                                                          1.0, 1.0, 1.0, True)


def test_product_does_not_import_the_oracle():
    pkg_dir = os.path.join(ROOT, "directvoxgo_b200")
    for path in glob.glob(os.path.join(pkg_dir, "**", "*.py"), recursive=True):
        src = open(path).read()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), path
        assert "libdvgo_oracle" not in src, path
    for path in glob.glob(os.path.join(pkg_dir, "csrc", "*")):
        assert "oracle/" not in open(path).read().replace("see oracle/dvgo_oracle.c", ""), path


def test_tensor_core_mlp_host_layout_zero_pads_narrow_widths():
    """Host logic of the tensor-core rgbnet wrapper (no kernel call): a 64-wide rgbnet is laid out as a 128-wide one
    with zero padding, the views returned by `unflatten` address exactly the real entries, and the module round-trips."""
    from directvoxgo_b200.fused_mlp import TensorCoreMLP
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(12, 64), torch.nn.ReLU(),
                              torch.nn.Sequential(torch.nn.Linear(64, 64), torch.nn.ReLU()), torch.nn.Linear(64, 3))
    tc = TensorCoreMLP(net, "cpu", train=True)
    W = TensorCoreMLP.WIDTH
    assert tc.params.numel() == W * 12 + W + W * W + W + 3 * W + 3
    real = tc.unflatten(tc.params)
    for v, p in zip(real, net.parameters()):
        assert v.shape == p.shape and torch.equal(v, p.detach())
    mask = torch.ones_like(tc.params, dtype=torch.bool)
    for v in tc.unflatten(mask):
        v.fill_(False)
    assert int(mask.sum()) == tc.params.numel() - sum(p.numel() for p in net.parameters())
    assert torch.all(tc.params[mask] == 0)
    with torch.no_grad():
        for v in tc.unflatten(tc.params):
            v.add_(1.0)
    tc.sync_to_module()
    for v, p in zip(tc.unflatten(tc.params), net.parameters()):
        assert torch.equal(v, p.detach())
    with pytest.raises(NotImplementedError):
        TensorCoreMLP(torch.nn.Sequential(torch.nn.Linear(12, 256), torch.nn.ReLU(),
                                          torch.nn.Sequential(torch.nn.Linear(256, 256), torch.nn.ReLU()),
                                          torch.nn.Linear(256, 3)), "cpu")


def test_ray_sharding_helpers():
    from directvoxgo_b200.parallel import shard_bounds, shard_views
    for n, world in ((8192, 8), (1000, 3), (5, 8)):
        spans = [shard_bounds(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
        assert max(e - s for s, e in spans) - min(e - s for s, e in spans) <= 1
    assert sorted(sum((shard_views(10, r, 4) for r in range(4)), [])) == list(range(10))


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) prints ONE JSON line with the contract's
    keys; run here on a small grid so that it takes seconds."""
    import json
    import subprocess
    import sys
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--grid", "24",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "rays/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    # under torchrun every rank but 0 exits without work
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    q = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--grid", "24"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert q.returncode == 0 and q.stdout.strip() == ""
