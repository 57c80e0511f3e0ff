"""CPU check of the factorised device arithmetic (exp_b200/csrc/sph_math.cuh, compiled for
the host by tests/emul/sph_emul.cu) against the oracle: the algebra the CUDA kernels
implement is validated here without a GPU; the kernels' own choreography is covered by the
-m gpu tests."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import acc_rel_err, get_table, rel_rms
from exp_b200 import _lib, ics
from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))


@pytest.fixture(scope="module")
def emul():
    so = os.path.join(HERE, "emul", "libsph_emul.so")
    src = os.path.join(HERE, "emul", "sph_emul.cu")
    if not os.path.exists(so) or os.path.getmtime(src) > os.path.getmtime(so):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++",
                        "-I/usr/local/cuda/include", "-o", so, src], check=True)
    return C.CDLL(so)


def _cfg(tab, **kw):
    d = dict(scale=1.0, NO_L0=0, NO_L1=0, EVEN_L=0, EVEN_M=0, M0_only=0)
    d.update({k: (int(v) if k != "scale" else v) for k, v in kw.items()})
    return _lib.SphereConfig(tab.lmax, tab.nmax, tab.numr, tab.cmap, tab.rmin, tab.rmax, tab.rmap,
                             d["scale"], tab.rmin, tab.rmax, d["NO_L0"], d["NO_L1"], d["EVEN_L"],
                             d["EVEN_M"], d["M0_only"], 0)


@pytest.mark.parametrize("kind,L,N", [("hernquist", 4, 10), ("nfw", 6, 18), ("hernquist", 1, 6)])
@pytest.mark.parametrize("flags", [dict(), dict(EVEN_M=True), dict(NO_L0=True, NO_L1=True),
                                   dict(M0_only=True), dict(scale=1.3)])
def test_factorised_math_matches_oracle(emul, kind, L, N, flags):
    tab = get_table(kind, L, N, rmax=50.0 if kind == "hernquist" else 20.0)
    n = 20000
    gen = ics.hernquist_halo if kind == "hernquist" else ics.nfw_halo
    pos, vel, mass = gen(n, seed=17)
    pos = pos + np.array([0.01, -0.02, 0.005])
    pos[:5] *= 100.0
    pos[5:8] *= 1e-7
    g = O.Grid(tab)
    S = O.ParticleSet(pos, vel, mass)
    P = O.sph_params(rmin=tab.rmin, rmax=tab.rmax, **flags)
    coef, used = O.sph_coefficients(g, P, S)
    O.sph_accel(g, P, S, coef)

    cfg = _cfg(tab, **flags)
    ev, ef, p0 = (np.ascontiguousarray(a) for a in (tab.ev, tab.ef, tab.p0))
    coef2 = np.zeros_like(coef)
    used2 = C.c_long(0)
    ctr = (C.c_double * 3)(0, 0, 0)
    rc = emul.emul_sph_coefficients(C.byref(cfg), dp(ev), dp(ef), dp(p0), C.c_long(n), dp(S.pos),
                                    dp(S.mass), ctr, C.c_double(1e20), C.c_double(1.0), dp(coef2),
                                    C.byref(used2))
    assert rc == 0 and used2.value == used
    assert rel_rms(coef2, coef) <= 1e-12          # budget is 1e-10; re-association only
    acc2 = np.zeros_like(S.acc)
    pot2 = np.zeros_like(S.pot)
    emul.emul_sph_accel(C.byref(cfg), dp(ev), dp(ef), dp(p0), dp(coef), C.c_long(n), dp(S.pos),
                        dp(S.mass), ctr, C.c_double(1e20), dp(acc2), dp(pot2))
    assert acc_rel_err(acc2, S.acc).max() <= 1e-9  # budget is 1e-8
    assert np.abs(pot2 - S.pot).max() <= 1e-9 * np.abs(S.pot).max()
