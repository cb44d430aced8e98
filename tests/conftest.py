import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def repo_root():
    return ROOT


def _cuda_devices():
    """Number of usable CUDA devices as the C-ABI library sees them (0 when the library is
    missing, the driver is absent, or no GPU is visible)."""
    try:
        import ctypes
        from deeprecsys_b200 import _lib
        n = ctypes.c_int(0)
        if _lib.load().drs_device_count(ctypes.byref(n)) != 0:
            return 0
        return int(n.value)
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a box without a GPU: skip the gpu tier instead of failing in drs_create.
    An explicit `-m gpu` run is left alone — there a missing device must fail loudly."""
    if "gpu" in (config.getoption("-m") or ""):
        return
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if not gpu_items or _cuda_devices() > 0:
        return
    skip = pytest.mark.skip(reason="no usable CUDA device (gpu tier runs on the B200 box)")
    for it in gpu_items:
        it.add_marker(skip)
