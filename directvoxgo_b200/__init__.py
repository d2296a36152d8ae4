"""directvoxgo_b200 -- B200-native (sm_100a) implementation of DirectVoxGO's per-ray
volume-rendering and grid-optimisation hot path, behind the reference's operator surface.

Layout
  csrc/                CUDA kernels + the C ABI (include/dvgo_b200.h) + the thin torch binding
  render_utils_cuda,   the three extension modules the reference JIT-builds
  total_variation_cuda,  (lib/dvgo.py:12-26, lib/masked_adam.py:5-10), same function names
  adam_upd_cuda
  ops                  autograd glue (Raw2Alpha, Alphas2Weights, trilinear sampler, segment_coo)
  masked_adam          MaskedAdam optimiser (lib/masked_adam.py)
  dvgo / dmpigo        DirectVoxGO / DirectMPIGO modules with the reference's ctor + forward contract
  fused                the fused B200 trainer / renderer (one ray kernel, tensor-core rgbnet, one sweep)
  dropin               makes the UNMODIFIED reference lib/*.py import these instead of JIT-building

There is no CPU fallback: importing the package without the built extension raises.
"""
import importlib
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))


def _load_native():
    try:
        import torch  # noqa: F401  (libtorch must be loaded before the extension)
        return importlib.import_module("directvoxgo_b200._C")
    except ImportError as e:  # fail loudly: no silent eager fallback
        raise ImportError(
            "directvoxgo_b200: the native extension directvoxgo_b200/_C*.so (and libdvgo_b200.so) "
            "is missing or does not load (%s). Build it in-tree with "
            "`python directvoxgo_b200/build.py`; there is no CPU/PyTorch fallback." % (e,)) from e


_C = _load_native()
render_utils_cuda = _C.render_utils_cuda
total_variation_cuda = _C.total_variation_cuda
adam_upd_cuda = _C.adam_upd_cuda
ext = _C.ext

# make `import directvoxgo_b200.render_utils_cuda` style imports work too
for _n in ("render_utils_cuda", "total_variation_cuda", "adam_upd_cuda", "ext"):
    sys.modules[__name__ + "." + _n] = getattr(_C, _n)

LIB_PATH = os.path.join(_HERE, "libdvgo_b200.so")
EXT_PATH = _C.__file__

__all__ = ["render_utils_cuda", "total_variation_cuda", "adam_upd_cuda", "ext", "LIB_PATH", "EXT_PATH"]
