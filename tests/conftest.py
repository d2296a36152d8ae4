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
    import torch
    # On a CUDA box a missing oracle/_ref is a FAILURE, not a skip: these are the only tests that compare with the
    # reference's own kernels live, and a silent skip would read as green (round-1 verdict).  oracle/_ref is built by
    # __graft_entry__.build() where /root/reference exists and travels with the snapshot.
    missing = pytest.fail if torch.cuda.is_available() else pytest.skip
    if not os.path.isdir(ref_dir):
        missing("oracle/_ref not built (run __graft_entry__.build() where /root/reference exists)")
    sys.path.insert(0, ref_dir)
    try:
        import ref_adam_upd_cuda
        import ref_render_utils_cuda
        import ref_total_variation_cuda
    except ImportError as e:
        missing("oracle/_ref not importable: %s" % e)
    import types
    return types.SimpleNamespace(render_utils_cuda=ref_render_utils_cuda,
                                 total_variation_cuda=ref_total_variation_cuda,
                                 adam_upd_cuda=ref_adam_upd_cuda)
