import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        import ctypes as C
        from exp_b200 import _lib
        n = C.c_int(0)
        return _lib.lib().expb200_device_count(C.byref(n)) == 0 and n.value > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


_TABLES = {}


def get_table(kind, lmax, nmax, numr=2000, rmin=1e-3, rmax=50.0):
    """SLGridSph tables are expensive-ish to build; cache them per session."""
    from exp_b200 import slgrid
    key = (kind, lmax, nmax, numr, rmin, rmax)
    if key not in _TABLES:
        model = slgrid.HernquistModel() if kind == "hernquist" else slgrid.NFWModel()
        _TABLES[key] = slgrid.make_slgrid(model, lmax, nmax, numr=numr, rmin=rmin, rmax=rmax)
    return _TABLES[key]


@pytest.fixture(scope="session")
def table_A():
    """config A: Hernquist, lmax=4, nmax=10"""
    return get_table("hernquist", 4, 10)


@pytest.fixture(scope="session")
def table_B():
    """config B: NFW, lmax=6, nmax=18"""
    return get_table("nfw", 6, 18, rmax=20.0)


def rel_rms(a, b):
    return float(np.sqrt(((np.asarray(a) - np.asarray(b)) ** 2).sum() / (np.asarray(b) ** 2).sum()))


def acc_rel_err(acc, ref):
    """per-particle error relative to the acceleration norm"""
    an = np.linalg.norm(ref, axis=1)
    return np.linalg.norm(acc - ref, axis=1) / np.maximum(an, 1e-300)
