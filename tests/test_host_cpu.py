"""CPU-side checks of the C++ host driver (libexphost.so): it loads, exports every symbol of
exp_b200/host/capi.h, and its initialize_multistep tables (src/multistep.cc:630-679) agree
with the oracle's restatement."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from exp_b200 import _lib
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def H():
    _lib.build_host()
    return C.CDLL(_lib.HOST_LIB_PATH)


def test_exports(H):
    hdr = open(os.path.join(ROOT, "exp_b200", "host", "capi.h")).read()
    names = set(re.findall(r"\b(exphost_\w+)\s*\(", hdr))
    assert len(names) >= 18
    for n in names:
        assert hasattr(H, n), n


@pytest.mark.parametrize("ms", [0, 1, 3, 6])
def test_multistep_tables(H, ms):
    Mstep = 1 << ms
    ip = C.POINTER(C.c_int)
    mintvl = np.zeros(ms + 1, dtype=np.int32)
    mfirst = np.zeros(Mstep + 1, dtype=np.int32)
    dL = np.zeros((ms + 1, Mstep), dtype=np.int32)
    dN = np.zeros((ms + 1, Mstep), dtype=np.int32)
    rc = H.exphost_multistep_tables(ms, mintvl.ctypes.data_as(ip), mfirst.ctypes.data_as(ip),
                                    dL.ctypes.data_as(ip), dN.ctypes.data_as(ip))
    assert rc == 0
    o_mintvl, o_mfirst, o_dL, o_dN = O.multistep_tables(ms)
    assert np.array_equal(mintvl, o_mintvl) and np.array_equal(mfirst, o_mfirst)
    assert np.array_equal(dL, o_dL) and np.array_equal(dN, o_dN)
    # CHK_STEP identity (src/step.cc:253-264): the drift intervals of a level sum to dtime
    for M in range(ms + 1):
        active = sum(1 for s in range(Mstep) if mfirst[s] <= M)
        assert active * mintvl[M] == Mstep


def test_errors_are_status_codes(H):
    H.exphost_last_error.restype = C.c_char_p
    assert H.exphost_multistep_tables(99, None, None, None, None) != 0
    assert b"multistep" in H.exphost_last_error()
