"""GPU parity of the cylinder path (Cylinder + EmpCylSL tables) through the C ABI against the
oracle: coefficients <= 1e-10 rel RMS, accelerations/potentials <= 1e-8 relative."""
import numpy as np
import pytest

from conftest import acc_rel_err, rel_rms
from exp_b200 import cyltable, ics
from exp_b200.component import Component
from exp_b200.cylinder import Cylinder
from oracle import oracle as O
from test_emul_cyl import disk, rotations

pytestmark = pytest.mark.gpu

_TABS = {}


def table(mmax, norder, numx, numy):
    key = (mmax, norder, numx, numy)
    if key not in _TABS:
        _TABS[key] = cyltable.synthetic_cyl_table(mmax=mmax, norder=norder, numx=numx, numy=numy)
    return _TABS[key]


def oracle_all(tab, pos, vel, mass, body=None, orig=None, lst=None, **flags):
    g = O.CylGridPy(tab, **flags)
    P = O.cyl_params(rcylmax=20.0, acyl=tab.ascale, body=body, orig=orig)
    S = O.ParticleSet(pos, vel, mass)
    cs, sn, used, cm = O.cyl_coefficients(g, P, S, lst=lst)
    O.cyl_accel(g, P, S, cs, sn, cm)
    return cs, sn, used, cm, S.acc, S.pot


@pytest.mark.parametrize("shape,n", [((6, 18, 128, 64), 200000), ((4, 6, 48, 24), 30000),
                                     ((0, 4, 16, 8), 3000), ((2, 5, 33, 17), 500)])
def test_coef_and_force_parity(shape, n):
    tab = table(*shape)
    pos, vel, mass = disk(n)
    cs_o, sn_o, used_o, cm_o, acc_o, pot_o = oracle_all(tab, pos, vel, mass)
    c = Component(n)
    c.upload(pos, vel, mass)
    cyl = Cylinder(tab)
    cs, sn = cyl.determine_coefficients(c)
    assert cyl.used == used_o
    assert abs(cyl.cylmass - cm_o) <= 1e-11 * cm_o      # serial (oracle) vs tree (GPU) summation
    assert rel_rms(cs, cs_o) <= 1e-10
    if shape[0]:
        assert rel_rms(sn[1:], sn_o[1:]) <= 1e-10
    cyl.set_coefficients(cs_o, sn_o, cm_o)
    c.zero_pot_acc()
    cyl.get_acceleration_and_potential(c)
    out = c.download()
    assert np.array_equal(out["pos"], pos)
    assert acc_rel_err(out["acc"], acc_o).max() <= 1e-8
    assert np.abs(out["pot"] - pot_o).max() <= 1e-8 * np.abs(pot_o).max()


@pytest.mark.parametrize("flags", [dict(EVEN_M=True), dict(MMIN=1, MLIM=3, NMIN=1, NLIM=4)])
def test_flags(flags):
    tab = table(4, 6, 48, 24)
    n = 20000
    pos, vel, mass = disk(n, seed=45)
    cs_o, sn_o, used_o, cm_o, acc_o, pot_o = oracle_all(tab, pos, vel, mass, **flags)
    c = Component(n)
    c.upload(pos, vel, mass)
    cyl = Cylinder(tab, **flags)
    cs, sn = cyl.determine_coefficients(c)
    assert rel_rms(cs, cs_o) <= 1e-10 and rel_rms(sn[1:], sn_o[1:]) <= 1e-10
    cyl.set_coefficients(None)                 # device vector straight into the force
    c.zero_pot_acc()
    cyl.get_acceleration_and_potential(c)
    assert acc_rel_err(c.download()["acc"], acc_o).max() <= 1e-8


def test_orientation_center_and_determinism():
    tab = table(4, 6, 48, 24)
    n = 40000
    pos, vel, mass = disk(n, seed=46)
    body, orig = rotations()
    ctr = np.array([0.002, -0.001, 0.0005])
    g = O.CylGridPy(tab)
    P = O.cyl_params(rcylmax=20.0, acyl=tab.ascale, body=body, orig=orig)
    S = O.ParticleSet(pos + ctr, vel, mass, center=ctr, rtrunc=0.08)
    cs_o, sn_o, used_o, cm_o = O.cyl_coefficients(g, P, S)
    O.cyl_accel(g, P, S, cs_o, sn_o, cm_o)
    res = []
    for _ in range(2):
        c = Component(n)
        c.upload(pos + ctr, vel, mass)
        c.set_center(ctr, rtrunc=0.08)
        cyl = Cylinder(tab)
        cyl.set_orientation(body, orig)
        cs, sn = cyl.determine_coefficients(c)
        assert cyl.used == used_o
        assert rel_rms(cs, cs_o) <= 1e-10 and rel_rms(sn[1:], sn_o[1:]) <= 1e-10
        cyl.set_coefficients(cs_o, sn_o, cm_o)
        c.zero_pot_acc()
        cyl.get_acceleration_and_potential(c)
        out = c.download()
        assert acc_rel_err(out["acc"], S.acc).max() <= 1e-8
        res.append((cs.tobytes(), sn.tobytes(), out["acc"].tobytes()))
        c.close(); cyl.close()
    assert res[0] == res[1]


def test_levels_and_multistep_tables():
    tab = table(4, 6, 48, 24)
    n = 30000
    pos, vel, mass = disk(n, seed=47)
    rng = np.random.default_rng(1)
    lev = rng.integers(0, 3, n).astype(np.uint32)
    g = O.CylGridPy(tab)
    P = O.cyl_params(rcylmax=20.0, acyl=tab.ascale)
    S = O.ParticleSet(pos, vel, mass, level=lev)
    c = Component(n)
    c.upload(pos, vel, mass, level=lev)
    cyl = Cylinder(tab, multistep=2)
    tot_c = np.zeros((tab.mmax + 1, tab.norder)); tot_s = np.zeros_like(tot_c); tot_m = 0.0
    for M in range(3):
        cs_o, sn_o, used_o, cm_o = O.cyl_coefficients(g, P, S, lst=S.levlist(M))
        cyl.set_multistep_level(M)
        cs, sn = cyl.determine_coefficients(c)
        assert cyl.used == used_o and abs(cyl.cylmass - cm_o) <= 1e-11 * max(cm_o, 1e-30)
        assert rel_rms(cs, cs_o) <= 1e-10 and rel_rms(sn[1:], sn_o[1:]) <= 1e-10
        cyl.multistep_commit(M)
        tot_c += cs_o; tot_s += sn_o; tot_m += cm_o
    cyl.compute_multistep_coefficients(4, cylmass=tot_m)     # mdrft = Mstep: plain sum of levels
    cyl.set_multistep_level(1)
    c.zero_pot_acc(0)
    cyl.get_acceleration_and_potential(c)
    act = np.nonzero(lev >= 1)[0].astype(np.int32)
    O.cyl_accel(g, P, S, tot_c, tot_s, tot_m, lst=act)
    out = c.download()
    assert acc_rel_err(out["acc"][act], S.acc[act]).max() <= 1e-8
    assert np.all(out["acc"][lev < 1] == 0.0)
