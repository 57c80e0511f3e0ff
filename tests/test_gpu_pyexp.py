"""GPU parity of the pyEXP-facing twins (exp_b200.pyexp) against the CPU restatement of
expui/BiorthBasis.cc (oracle/pyexp_oracle.py): createFromArray coefficients <= 1e-10 rel RMS,
getFields / getAccel <= 1e-8 relative."""
import numpy as np
import pytest

from conftest import get_table, rel_rms
from exp_b200 import ics, pyexp
from oracle import oracle as O
from oracle import pyexp_oracle as PO

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("L,N,flags", [(4, 10, {}), (2, 6, dict(NO_L0=True)), (3, 8, dict(EVEN_M=True))])
def test_spherical_createFromArray_and_getFields(L, N, flags):
    tab = get_table("hernquist", L, N)
    n = 3000
    pos, vel, mass = ics.hernquist_halo(n, seed=3)
    ctr = np.array([0.02, -0.01, 0.03])
    pos = pos + ctr + np.array([0.01, -0.02, 0.005])
    g = O.Grid(tab)
    want, used = PO.accumulate(g, mass, pos - ctr, rmin=tab.rmin, rmax=tab.rmax)

    b = pyexp.SphericalSL(tab, **flags)
    cf = b.createFromArray(mass, pos, time=1.5, center=ctr)
    assert b.used == used and cf.time == 1.5 and cf.coefs.shape == ((L + 1) * (L + 2) // 2, N)
    assert rel_rms(cf.coefs.real, PO.load_coefs(want, L).real) <= 1e-10
    assert rel_rms(cf.coefs.imag, PO.load_coefs(want, L).imag) <= 1e-10
    # posvelrows layout and a rotation give the same answer as transforming by hand
    th = 0.4
    R = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]])
    cf2 = b.createFromArray(mass, np.ascontiguousarray(pos.T), center=ctr, rot=R, posvelrows=True)
    want2, _ = PO.accumulate(g, mass, (pos - ctr) @ R.T, rmin=tab.rmin, rmax=tab.rmax)
    assert rel_rms(cf2.coefs.real, PO.load_coefs(want2, L).real) <= 1e-10

    # fields: install the reference coefficients on both sides
    b.set_coefs(pyexp.SphStruct(L, N, 1.0, 0.0, PO.load_coefs(want, L), tuple(ctr)))
    rng = np.random.default_rng(1)
    pts = rng.normal(size=(40, 3)) * np.array([1.0, 1.0, 0.6]) * rng.uniform(0.05, 5.0, size=(40, 1))
    got = b.getFields(pts[:, 0] + ctr[0], pts[:, 1] + ctr[1], pts[:, 2] + ctr[2])
    acc = b.getAccel(pts + ctr)
    for k, p in enumerate(pts):
        ref = np.array(PO.crt_eval(g, want, p[0], p[1], p[2], **flags))
        fnorm = np.linalg.norm(ref[6:9])
        assert np.abs(got[3:6, k] - ref[3:6]).max() <= 1e-8 * max(abs(ref[5]), abs(ref[3])), k
        assert np.abs(got[6:9, k] - ref[6:9]).max() <= 1e-8 * fnorm, k
        assert np.abs(acc[k] - ref[6:9]).max() <= 1e-8 * fnorm, k
        assert np.abs(got[0:3, k] - ref[0:3]).max() <= 1e-8 * max(abs(ref[2]), abs(ref[0])), k
    one = b.getFields(float(pts[0, 0] + ctr[0]), float(pts[0, 1] + ctr[1]), float(pts[0, 2] + ctr[2]))
    assert len(one) == 9 and abs(one[5] - got[5, 0]) <= 1e-14 * abs(got[5, 0])


def test_cylindrical_createFromArray():
    from exp_b200 import cyltable
    from test_emul_cyl import disk
    tab = cyltable.synthetic_cyl_table(mmax=4, norder=6, numx=48, numy=24)
    n = 20000
    pos, vel, mass = disk(n)
    g = O.CylGridPy(tab)
    P = O.cyl_params(rcylmax=20.0, acyl=tab.ascale)
    S = O.ParticleSet(pos, vel, mass)
    cs, sn, used, cm = O.cyl_coefficients(g, P, S)
    b = pyexp.Cylindrical(tab)
    cf = b.createFromArray(mass, pos)
    assert b.used == used and cf.coefs.shape == (5, 6)
    assert rel_rms(cf.coefs.real, cs) <= 1e-10 and rel_rms(cf.coefs.imag[1:], sn[1:]) <= 1e-10
    O.cyl_accel(g, P, S, cs, sn, cm)
    acc = b.getAccel(pos[:500])
    err = np.linalg.norm(acc - S.acc[:500], axis=1) / np.linalg.norm(S.acc[:500], axis=1)
    assert err.max() <= 1e-8
