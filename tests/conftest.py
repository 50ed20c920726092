import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_devices():
    try:
        import ctypes as C
        from frankenz_b200 import _lib
        n = C.c_int(0)
        if _lib.load().fzb_device_count(C.byref(n)) != 0:
            return 0
        return n.value
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """`-m gpu` tests need the CUDA library and a device: skip (not fail) where there is none."""
    gpu_items = [it for it in items if it.get_closest_marker("gpu") is not None]
    if not gpu_items or _cuda_devices() > 0:
        return
    skip = pytest.mark.skip(reason="no usable CUDA device (the product has no CPU fallback)")
    for it in gpu_items:
        it.add_marker(skip)


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.fixture(scope="session")
def sdss_mock():
    d = golden("sdss_cww_mock.npz")
    return d["phot_obs"], d["phot_err"], d["redshifts"]


def same_special(a, b):
    """inf / nan positions agree."""
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return (np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(np.isposinf(a), np.isposinf(b))
            and np.array_equal(np.isneginf(a), np.isneginf(b)))


def max_rel(a, b, floor=0.0):
    """max |a-b| / max(|b|, floor) over entries where b is finite."""
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    ok = np.isfinite(b)
    if not ok.any():
        return 0.0
    den = np.maximum(np.abs(b[ok]), floor) if floor > 0 else np.abs(b[ok])
    num = np.abs(a[ok] - b[ok])
    with np.errstate(divide="ignore", invalid="ignore"):
        r = np.where(num == 0, 0.0, num / den)
    return float(np.max(r))
