"""GPU parity of the multistep machinery: level assignment (bit-exact), per-level
coefficient tables, interpolation, level-switch updates and the full do_step ordering."""
import numpy as np
import pytest

from conftest import acc_rel_err, get_table, rel_rms
from exp_b200 import ics
from exp_b200.component import Component
from exp_b200.sphere import Sphere
from exp_b200.stepper import SphereStepper
from oracle import oracle as O
from ref_stepper import RefSphereStepper

pytestmark = pytest.mark.gpu


def _state(n, seed=31):
    pos, vel, mass = ics.hernquist_halo(n, seed=seed)
    pos = pos + np.array([0.01, -0.02, 0.005])
    return pos, vel, mass


@pytest.mark.parametrize("dtime", [0.01, 0.0078125, 1.0])     # generic, and exact powers of two
def test_level_assignment_bit_exact(dtime):
    n = 200000
    rng = np.random.default_rng(5)
    pos, vel, mass = _state(n)
    acc = rng.normal(size=(n, 3)) * np.exp(rng.normal(size=(n, 1)) * 2.0)
    pot = -np.abs(rng.normal(size=n))
    vel = vel * np.exp(rng.normal(size=(n, 1)))
    # adversarial: make dtv exactly dtime/2^k for some particles (exact powers of two)
    vel[:64] = 0.0; vel[:64, 0] = 1.0
    acc[:64] = 0.0; acc[:64, 1] = np.repeat(2.0 ** np.arange(8), 8) / dtime
    lev0 = rng.integers(0, 5, n).astype(np.uint32)
    scale = np.where(rng.random(n) < 0.5, -1.0, rng.random(n)).astype(np.float32)
    for kw in [dict(), dict(shiftlevl=1), dict(mfirst_mdrft=2), dict(dynfracV=1.0, dynfracA=1e30, dynfracP=1e30)]:
        S = O.ParticleSet(pos, vel, mass, level=lev0, scale=scale)
        S.acc[:] = acc; S.pot[:] = pot
        LP = O.level_params(dtime, 4, **kw)
        nsw, which, offlo, offhi = O.adjust_levels(S, LP)
        c = Component(n)
        c.upload(pos, vel, mass, level=lev0, scale=scale)
        c.set_acc_pot(acc, pot)
        st = c.adjust_multistep_level(dtime, 4, **kw)
        out = c.download(pos=False, vel=False, acc=False, pot=False, level=True, dtreq=True)
        assert np.array_equal(out["level"], S.level), kw
        assert out["dtreq"].tobytes() == S.dtreq.tobytes(), kw
        assert st["switched"] == nsw and st["offlo"] == offlo and st["offhi"] == offhi
        assert np.array_equal(c.level_counts(4), np.bincount(S.level, minlength=5))
        c.close()


def test_per_level_coefficients_and_force():
    tab = get_table("hernquist", 4, 10)
    n = 60000
    pos, vel, mass = _state(n, seed=33)
    rng = np.random.default_rng(2)
    lev = rng.integers(0, 4, n).astype(np.uint32)
    g = O.Grid(tab)
    P = O.sph_params(rmin=tab.rmin, rmax=tab.rmax)
    S = O.ParticleSet(pos, vel, mass, level=lev)
    c = Component(n)
    c.upload(pos, vel, mass, level=lev)
    s = Sphere(tab, multistep=3)
    tot = np.zeros(((tab.lmax + 1) ** 2, tab.nmax))
    for M in range(4):
        co, used = O.sph_coefficients(g, P, S, lst=S.levlist(M))
        s.set_multistep_level(M)
        cg = s.determine_coefficients(c)
        assert s.used == used
        assert rel_rms(cg, co) <= 1e-10
        s.multistep_commit(M)
        tot += co
    # force on levels >= 2 only, with the summed coefficients (mdrft = Mstep: no interpolation)
    s.compute_multistep_coefficients(8)
    s.set_multistep_level(2)
    c.zero_pot_acc(0)
    s.get_acceleration_and_potential(c)
    act = np.nonzero(lev >= 2)[0].astype(np.int32)
    O.sph_accel(g, P, S, tot, lst=act)
    out = c.download()
    assert acc_rel_err(out["acc"][act], S.acc[act]).max() <= 1e-8
    assert np.all(out["acc"][lev < 2] == 0.0)


@pytest.mark.gpu
def test_move_levels_equals_per_level_calls():
    """expb200_move_levels = the per-level incr_velocity / incr_position loop of a sub-step (reference
    src/step.cc:215-237, 253-262), bit for bit, with and without a known level partition"""
    tab = get_table("hernquist", 4, 10)
    n = 50000
    pos, vel, mass = _state(n, seed=35)
    rng = np.random.default_rng(3)
    lev = rng.integers(0, 5, n).astype(np.uint32)
    acc = rng.normal(size=(n, 3))
    dtk = np.array([0.0, 0.0, 0.004, 0.002, 0.001])
    dtd = 2.0 * dtk
    out = []
    for fused in (False, True, True):
        c = Component(n)
        c.upload(pos, vel, mass, level=lev)
        c.set_acc_pot(acc, np.zeros(n))
        if len(out) == 2:           # a sorted component: the active tail of the permutation is walked
            s = Sphere(tab, multistep=4)
            s.set_multistep_level(0)
            s.determine_coefficients(c)
            c.set_acc_pot(acc, np.zeros(n))
        if fused:
            c.move_levels(dtk, dtd, 2)
            c.move_levels(dtk, None, 3)
        else:
            for M in (2, 3, 4):
                c.incr_velocity(dtk[M], M)
                c.incr_position(dtd[M], M)
            for M in (3, 4):
                c.incr_velocity(dtk[M], M)
        out.append(c.download())
    for k in ("pos", "vel"):
        assert np.array_equal(out[0][k], out[1][k]) and np.array_equal(out[0][k], out[2][k])
    assert np.array_equal(out[0]["pos"][lev < 2], pos[lev < 2])
    assert not np.array_equal(out[0]["pos"][lev >= 2], pos[lev >= 2])


def test_multistep_interpolation_and_update():
    tab = get_table("hernquist", 2, 8, numr=400)
    n = 20000
    pos, vel, mass = _state(n, seed=35)
    rng = np.random.default_rng(3)
    ms = 3
    s = Sphere(tab, multistep=ms)
    shape = ((tab.lmax + 1) ** 2, tab.nmax)
    cL, cN = rng.normal(size=(ms + 1,) + shape), rng.normal(size=(ms + 1,) + shape)
    for M in range(ms + 1):
        s.set_level_coefficients("L", M, cL[M])
        s.set_level_coefficients("N", M, cN[M])
    c = Component(n)
    c.upload(pos, vel, mass)
    g = O.Grid(tab)
    P = O.sph_params(rmin=tab.rmin, rmax=tab.rmax)
    for mdrft in range((1 << ms) + 1):
        want = O.multistep_interp(ms, mdrft, cL, cN)
        s.compute_multistep_coefficients(mdrft)
        # read back through a force evaluation with the oracle using `want`
        S = O.ParticleSet(pos[:2000], vel[:2000], mass[:2000])
        O.sph_accel(g, P, S, want)
        c2 = Component(2000)
        c2.upload(pos[:2000], vel[:2000], mass[:2000])
        s.set_multistep_level(0)
        s.get_acceleration_and_potential(c2)
        assert acc_rel_err(c2.download()["acc"], S.acc).max() <= 1e-8, mdrft
        c2.close()

    # level switches: differ tables
    old = rng.integers(0, ms + 1, n).astype(np.uint32)
    c.set_levels(old)
    acc = rng.normal(size=(n, 3)) * np.exp(rng.normal(size=(n, 1)) * 2.0)
    c.set_acc_pot(acc, -np.ones(n))
    c.adjust_multistep_level(0.05, ms)
    new = c.download(pos=False, vel=False, acc=False, pot=False, level=True)["level"]
    assert (new != old).sum() > 100
    s.multistep_update(c, finish=False)
    S = O.ParticleSet(pos, vel, mass)
    differ = np.zeros((ms + 1, shape[0] * shape[1]))
    for i in np.nonzero(new != old)[0]:
        O.sph_multistep_update(g, P, S, int(i), int(old[i]), int(new[i]), differ)
    for M in range(ms + 1):
        assert rel_rms(s.level_coefficients("differ", M).ravel(), differ[M]) <= 1e-10, M
    s.multistep_update_finish()
    for M in range(ms + 1):
        assert rel_rms(s.level_coefficients("N", M).ravel(), cN[M].ravel() + differ[M]) <= 1e-10


@pytest.mark.parametrize("multistep", [0, 3])
def test_full_step_against_reference_ordering(multistep):
    tab = get_table("hernquist", 2, 8, numr=400)
    n = 20000
    pos, vel, mass = _state(n, seed=37)
    dtime = 0.02
    dyn = dict(dynfracV=0.02, dynfracA=0.06, dynfracP=0.1)
    ref = RefSphereStepper(tab, pos, vel, mass, dtime, multistep, **dyn)
    ref.initialize()
    c = Component(n)
    c.upload(pos, vel, mass)
    s = Sphere(tab, multistep=multistep)
    drv = SphereStepper(c, s, dtime, multistep, **dyn)
    drv.initialize()
    for _ in range(2):
        ref.step()
        drv.step()
    out = c.download(level=True, dtreq=True)
    assert np.array_equal(out["level"], ref.S.level)                   # bit-exact levels
    if multistep:
        assert len(np.unique(ref.S.level)) >= 3                          # the hierarchy is exercised
    assert np.abs(out["pos"] - ref.S.pos).max() <= 1e-10 * np.abs(ref.S.pos).max()
    assert np.abs(out["vel"] - ref.S.vel).max() <= 1e-9 * np.abs(ref.S.vel).max()
    assert acc_rel_err(out["acc"], ref.S.acc).max() <= 1e-8
    if multistep:
        for M in range(multistep + 1):
            assert rel_rms(s.level_coefficients("N", M), ref.coefN[M]) <= 1e-9, M


def test_cylinder_multistep_update():
    """CylEXP::multistep_update (src/CylEXP.cc:159-188, src/Cylinder.cc:1751-1773): level switches
    move each particle's contribution between the per-level cos/sin tables."""
    from exp_b200 import cyltable
    from exp_b200.cylinder import Cylinder
    from test_emul_cyl import disk, rotations
    tab = cyltable.synthetic_cyl_table(mmax=4, norder=6, numx=48, numy=24)
    n, ms = 20000, 3
    pos, vel, mass = disk(n)
    rng = np.random.default_rng(11)
    cyl = Cylinder(tab, multistep=ms)
    body, orig = rotations()
    cyl.set_orientation(body, orig)            # the update ignores the orientation (reference)
    c = Component(n)
    c.upload(pos, vel, mass)
    old = rng.integers(0, ms + 1, n).astype(np.uint32)
    c.set_levels(old)
    acc = rng.normal(size=(n, 3)) * np.exp(rng.normal(size=(n, 1)) * 2.0)
    c.set_acc_pot(acc, -np.ones(n))
    # per-level tables before the update
    cN = []
    for M in range(ms + 1):
        cyl.set_multistep_level(M)
        cyl.determine_coefficients(c, to_host=False)
        cyl.multistep_commit(M)
        cN.append(cyl.level_coefficients("N", M))
    c.adjust_multistep_level(0.05, ms)
    new = c.download(pos=False, vel=False, acc=False, pot=False, level=True)["level"]
    assert (new != old).sum() > 100
    cyl.multistep_update(c, finish=False)
    g = O.CylGridPy(tab)
    S = O.ParticleSet(pos, vel, mass)
    shape = (ms + 1, tab.mmax + 1, tab.norder)
    dC, dS = np.zeros(shape), np.zeros(shape)
    for i in np.nonzero(new != old)[0]:
        O.cyl_multistep_update(g, S, int(i), int(old[i]), int(new[i]), dC, dS)
    scale = np.abs(dC).max()
    for M in range(ms + 1):
        gc, gs = cyl.level_coefficients("differ", M)
        assert np.abs(gc - dC[M]).max() <= 1e-10 * scale, M
        assert np.abs(gs[1:] - dS[M][1:]).max() <= 1e-10 * scale, M
    cyl.multistep_update_finish()
    for M in range(ms + 1):
        gc, gs = cyl.level_coefficients("N", M)
        assert np.abs(gc - (cN[M][0] + dC[M])).max() <= 1e-10 * max(scale, np.abs(cN[M][0]).max())
