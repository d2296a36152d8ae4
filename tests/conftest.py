import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def ref_gpu():
    """The reference's own CUDA extensions built from /root/reference into oracle/_ref (GPU only)."""
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.isdir(ref_dir):
        pytest.skip("oracle/_ref not built")
    sys.path.insert(0, ref_dir)
    import torch  # noqa: F401
    try:
        import ref_adam_upd_cuda
        import ref_render_utils_cuda
        import ref_total_variation_cuda
    except ImportError as e:
        pytest.skip("oracle/_ref not importable: %s" % e)
    import types
    return types.SimpleNamespace(render_utils_cuda=ref_render_utils_cuda,
                                 total_variation_cuda=ref_total_variation_cuda,
                                 adam_upd_cuda=ref_adam_upd_cuda)
