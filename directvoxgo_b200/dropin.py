"""Run the UNMODIFIED reference (lib/dvgo.py, lib/dmpigo.py, lib/tri_dvgo.py, lib/masked_adam.py)
on the B200 kernels.

The reference builds its three extensions at import time with
`torch.utils.cpp_extension.load(name='render_utils_cuda' | 'total_variation_cuda' | 'adam_upd_cuda',
sources=[...])` (lib/dvgo.py:12-26, lib/tri_dvgo.py:18-32, lib/masked_adam.py:5-10) and imports
`torch_scatter` by name (lib/dvgo.py:10).  `install()` therefore
  1. wraps `cpp_extension.load` so that those three names resolve to our prebuilt modules (any other
     name still goes to the real JIT), and
  2. registers `directvoxgo_b200.torch_scatter_shim` as `torch_scatter` if that package is absent.
Usage, with no edit to the reference tree:

    import directvoxgo_b200.dropin; directvoxgo_b200.dropin.install()
    import run            # the reference's driver; lib.dvgo now runs on libdvgo_b200.so
"""
import importlib.util
import sys

_NAMES = ("render_utils_cuda", "total_variation_cuda", "adam_upd_cuda")
_installed = False


def install(modules=None, grid_sample=False):
    """`modules`: optional {name: module} override (the test-suite injects the CPU oracle here to run
    the reference's Python orchestration on a GPU-less box; the product never does).
    `grid_sample=True` additionally routes `torch.nn.functional.grid_sample` -- which the reference calls by
    name for DenseGrid sampling (lib/dvgo.py:321, lib/dmpigo.py:169) and tri-plane sampling
    (lib/tri_dvgo.py:462-464) -- to `dvgo_grid_sample_3d_norm / _2d_norm` for the call shapes the reference
    uses; any other call still reaches ATen."""
    global _installed
    import torch.utils.cpp_extension as cpp_ext
    if grid_sample:
        import torch.nn.functional as F
        from .ops import make_grid_sample
        if not hasattr(F.grid_sample, "_dvgo_real"):
            F.grid_sample = make_grid_sample(F.grid_sample)

    if modules is None:
        import directvoxgo_b200 as pkg
        modules = {n: getattr(pkg, n) for n in _NAMES}
    real_load = getattr(cpp_ext.load, "_dvgo_real_load", cpp_ext.load)

    def load(name, sources=None, *args, **kwargs):
        if name in modules:
            return modules[name]
        return real_load(name, sources, *args, **kwargs)

    load._dvgo_real_load = real_load
    cpp_ext.load = load
    if "torch_scatter" not in modules and importlib.util.find_spec("torch_scatter") is None:
        from . import torch_scatter_shim
        sys.modules["torch_scatter"] = torch_scatter_shim
    elif "torch_scatter" in modules:
        sys.modules["torch_scatter"] = modules["torch_scatter"]
    _installed = True


def uninstall():
    global _installed
    import torch.utils.cpp_extension as cpp_ext
    real = getattr(cpp_ext.load, "_dvgo_real_load", None)
    if real is not None:
        cpp_ext.load = real
    import torch.nn.functional as F
    if hasattr(F.grid_sample, "_dvgo_real"):
        F.grid_sample = F.grid_sample._dvgo_real
    mod = sys.modules.get("torch_scatter")
    if mod is not None and getattr(mod, "__name__", "").startswith("directvoxgo_b200"):
        del sys.modules["torch_scatter"]
    _installed = False
