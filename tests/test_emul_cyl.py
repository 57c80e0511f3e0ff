"""CPU check of the cylinder's factorised device arithmetic (exp_b200/csrc/cyl_math.cuh via
tests/emul/cyl_emul.cu) against the oracle."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import acc_rel_err, rel_rms
from exp_b200 import _lib, cyltable, ics
from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


@pytest.fixture(scope="module")
def emul():
    so = os.path.join(HERE, "emul", "libcyl_emul.so")
    src = os.path.join(HERE, "emul", "cyl_emul.cu")
    if not os.path.exists(so) or os.path.getmtime(src) > os.path.getmtime(so):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++",
                        "-I/usr/local/cuda/include", "-o", so, src], check=True)
    return C.CDLL(so)


def disk(n, seed=44, a=0.01, h=0.001):
    pos, vel, mass = ics.exponential_disk(n, seed=seed, a=a, h=h)
    pos = pos + np.array([2e-4, -1e-4, 5e-5])
    pos[:6] *= 40.0          # beyond the table and beyond rcylmax*acyl
    pos[6:10] *= 12.0        # into the erf blending zone
    return pos, vel, mass


def rotations():
    th, ph = 0.3, -0.5
    Rz = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]])
    Rx = np.array([[1.0, 0, 0], [0, np.cos(ph), -np.sin(ph)], [0, np.sin(ph), np.cos(ph)]])
    B = Rz @ Rx
    return B, B.T.copy()


@pytest.mark.parametrize("flags", [dict(), dict(EVEN_M=True), dict(MMIN=1, MLIM=3, NMIN=1, NLIM=4),
                                   dict(rot=True)])
def test_cylinder_math_matches_oracle(emul, flags):
    flags = dict(flags)
    rot = flags.pop("rot", False)
    tab = cyltable.synthetic_cyl_table(mmax=4, norder=6, numx=48, numy=24)
    n = 20000
    pos, vel, mass = disk(n)
    body, orig = rotations() if rot else (None, None)
    g = O.CylGridPy(tab, **flags)
    P = O.cyl_params(rcylmax=20.0, acyl=tab.ascale, body=body, orig=orig)
    S = O.ParticleSet(pos, vel, mass)
    cs, sn, used, cm = O.cyl_coefficients(g, P, S)
    O.cyl_accel(g, P, S, cs, sn, cm)

    d = dict(EVEN_M=0, MMIN=0, MLIM=2 ** 30, NMIN=0, NLIM=2 ** 30)
    d.update({k: int(v) for k, v in flags.items()})
    cfg = _lib.CylinderConfig(tab.mmax, tab.norder, tab.numx, tab.numy, tab.cmapr, tab.cmapz,
                              tab.ascale, tab.hscale, tab.rmin, tab.rmax, 20.0, tab.ascale,
                              d["EVEN_M"], d["MMIN"], d["MLIM"], d["NMIN"], d["NLIM"], 0)
    tabs = [np.ascontiguousarray(getattr(tab, k)) for k in
            ("potC", "potS", "rforceC", "rforceS", "zforceC", "zforceS")]
    cs2, sn2 = np.zeros_like(cs), np.zeros_like(sn)
    used2, cm2 = C.c_long(0), C.c_double(0)
    ctr = (C.c_double * 3)(0, 0, 0)
    b = None if body is None else np.ascontiguousarray(body)
    o = None if orig is None else np.ascontiguousarray(orig)
    rc = emul.emul_cyl_coefficients(C.byref(cfg), dp(tabs[0]), dp(tabs[1]), C.c_long(n), dp(S.pos),
                                    dp(S.mass), ctr, C.c_double(1e20), C.c_double(1.0), dp(b),
                                    dp(cs2), dp(sn2), C.byref(used2), C.byref(cm2))
    assert rc == 0 and used2.value == used and abs(cm2.value - cm) < 1e-14
    assert rel_rms(cs2, cs) <= 1e-12 and rel_rms(sn2[1:], sn[1:]) <= 1e-12
    acc2, pot2 = np.zeros_like(S.acc), np.zeros_like(S.pot)
    ptrs = (C.POINTER(C.c_double) * 6)(*[dp(t) for t in tabs])
    emul.emul_cyl_accel(C.byref(cfg), ptrs, dp(cs), dp(sn), C.c_double(cm), C.c_long(n), dp(S.pos),
                        dp(S.mass), ctr, C.c_double(1e20), dp(b), dp(o), dp(acc2), dp(pot2))
    assert acc_rel_err(acc2, S.acc).max() <= 1e-9
    assert np.abs(pot2 - S.pot).max() <= 1e-9 * np.abs(S.pot).max()
