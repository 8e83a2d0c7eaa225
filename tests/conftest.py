import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _have_gpu() -> bool:
    try:
        import ctypes

        from rmf_crowdsim_b200 import _native

        lib = _native.load()
        desc = _native.SimDesc(4.0, 4.0, 1.0, 0.0, 0.0, 4, 0, 0)
        h = ctypes.c_void_p()
        rc = lib.rcs_sim_create(ctypes.byref(desc), ctypes.byref(h))
        if rc == 0:
            lib.rcs_sim_destroy(h)
            return True
        return False
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu_available():
    return _have_gpu()


def pytest_collection_modifyitems(config, items):
    # A `-m gpu` run on a box without a device must FAIL loudly, not skip: nothing to do here.
    # A plain run (no -m) on a CPU box skips the gpu tests.
    markexpr = config.getoption("-m") or ""
    if "gpu" in markexpr:
        return
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
