import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a machine without a GPU skips the gpu-marked tests instead of erroring in
    the fixture.  With `-m gpu` (the GPU box) nothing is skipped: a missing device fails loudly."""
    if "gpu" in (config.getoption("-m") or ""):
        return
    from delayrepay_b200 import _lib
    if _lib.gpu_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device (gpu-marked tests run with -m gpu on the B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def gpu():
    """The CUDA engine on device 0; fails loudly (no skip, no fallback) without a device."""
    import delayrepay_b200 as dr
    dr.set_device(0)
    return dr
