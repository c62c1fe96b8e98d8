import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "dem-engine_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_device_count():
    """Number of CUDA devices, asked of the driver itself (no torch import: it takes a minute on a fresh box)."""
    import ctypes
    try:
        cuda = ctypes.CDLL("libcuda.so.1")
    except OSError:
        return 0
    n = ctypes.c_int(0)
    if cuda.cuInit(0) != 0 or cuda.cuDeviceGetCount(ctypes.byref(n)) != 0:
        return 0
    return n.value


def pytest_collection_modifyitems(config, items):
    """Tests marked gpu are skipped (not failed) on a machine without a CUDA device; the product itself never falls
    back to the CPU (dem_ctx_create returns DEM_ERR_NO_GPU)."""
    if _cuda_device_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device: the product path has no CPU fallback")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def built():
    """Build the oracle (+ oracle/_ref when the reference tree is present) and make sure libdemcore.so exists."""
    import __graft_entry__ as g
    g.build()
    return True
