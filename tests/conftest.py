import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """GPU tests must never pass silently without a GPU: skip them (visibly) on CPU boxes."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def ref_cython():
    """The real reference utils/nms.pyx (oracle/_ref), or skip."""
    from oracle import build_ref
    build_ref.build()
    mod = build_ref.load()
    if mod is None:
        pytest.skip("oracle/_ref not built (reference tree absent)")
    return mod


@pytest.fixture(scope="session")
def ref_py(ref_cython):
    """The reference's own Python functions (needs /root/reference mounted), or skip."""
    from oracle import ref_py2
    if not ref_py2.available():
        pytest.skip("/root/reference not mounted")
    return ref_py2.RefFunctions(ref_cython)
