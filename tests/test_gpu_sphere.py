"""GPU parity tests of the sphereSL path: the CUDA library called through the C ABI
against the CPU oracle on identical particle sets.

Tolerances are BASELINE.json's: coefficients <= 1e-10 relative RMS (fp64), accelerations
and potentials <= 1e-8 relative, particle order bit-exact."""
import numpy as np
import pytest

from conftest import acc_rel_err, get_table, rel_rms
from exp_b200 import ics
from exp_b200.component import Component
from exp_b200.sphere import Sphere
from oracle import oracle as O

pytestmark = pytest.mark.gpu

COEF_TOL = 1e-10
ACC_TOL = 1e-8


def _halo(kind, n, seed=7):
    gen = ics.hernquist_halo if kind == "hernquist" else ics.nfw_halo
    pos, vel, mass = gen(n, seed=seed)
    pos = pos + np.array([0.01, -0.02, 0.005])      # break the symmetry: all (l,m) populated
    return pos, vel, mass


def _oracle_all(tab, pos, vel, mass, **flags):
    g = O.Grid(tab)
    S = O.ParticleSet(pos, vel, mass)
    P = O.sph_params(rmin=tab.rmin, rmax=tab.rmax, **flags)
    coef, used = O.sph_coefficients(g, P, S)
    O.sph_accel(g, P, S, coef)
    return coef, used, S.acc, S.pot


@pytest.mark.parametrize("kind,L,N,n", [("hernquist", 4, 10, 100000), ("nfw", 6, 18, 200000),
                                        ("hernquist", 2, 10, 10000), ("nfw", 0, 6, 5000),
                                        ("hernquist", 7, 12, 30000)])
def test_coef_and_force_parity(kind, L, N, n):
    tab = get_table(kind, L, N, rmax=50.0 if kind == "hernquist" else 20.0)
    pos, vel, mass = _halo(kind, n)
    pos[:5] *= 100.0        # beyond rmax: skipped by coefs, multipole continuation in force
    pos[5:8] *= 1e-7        # inside rmin
    coef_o, used_o, acc_o, pot_o = _oracle_all(tab, pos, vel, mass)

    c = Component(n)
    c.upload(pos, vel, mass)
    s = Sphere(tab)
    coef = s.determine_coefficients(c)
    assert s.used == used_o
    assert rel_rms(coef, coef_o) <= COEF_TOL

    s.set_coefficients(coef_o)          # same coefficients on both sides for the force check
    c.zero_pot_acc()
    s.get_acceleration_and_potential(c)
    out = c.download()
    assert np.array_equal(out["pos"], pos)       # order and values untouched
    assert acc_rel_err(out["acc"], acc_o).max() <= ACC_TOL
    assert np.abs(out["pot"] - pot_o).max() <= ACC_TOL * np.abs(pot_o).max()
    mask = np.abs(pot_o) > 1e-3 * np.abs(pot_o).max()
    assert (np.abs(out["pot"] - pot_o)[mask] / np.abs(pot_o[mask])).max() <= ACC_TOL


def test_device_coefficients_feed_force():
    """coefficients accumulated on the device, used without a host round trip"""
    tab = get_table("hernquist", 4, 10)
    n = 50000
    pos, vel, mass = _halo("hernquist", n, seed=11)
    coef_o, _, acc_o, pot_o = _oracle_all(tab, pos, vel, mass)
    c = Component(n)
    c.upload(pos, vel, mass)
    s = Sphere(tab)
    s.determine_coefficients(c, to_host=False)
    s.set_coefficients(None)
    c.zero_pot_acc()
    s.get_acceleration_and_potential(c)
    out = c.download()
    assert acc_rel_err(out["acc"], acc_o).max() <= ACC_TOL


def test_deterministic():
    """same input -> bit-identical coefficients and forces, run to run"""
    tab = get_table("nfw", 6, 18, rmax=20.0)
    n = 300000
    pos, vel, mass = _halo("nfw", n, seed=3)
    res = []
    for _ in range(3):
        c = Component(n)
        c.upload(pos, vel, mass)
        s = Sphere(tab)
        coef = s.determine_coefficients(c)
        s.set_coefficients(None)
        c.zero_pot_acc()
        s.get_acceleration_and_potential(c)
        out = c.download()
        res.append((coef.tobytes(), out["acc"].tobytes(), out["pot"].tobytes()))
        c.close(); s.close()
    assert res[0] == res[1] == res[2]


@pytest.mark.parametrize("flags", [dict(EVEN_M=True), dict(NO_L0=True, NO_L1=True),
                                   dict(EVEN_L=True), dict(M0_only=True), dict(scale=1.3)])
def test_flags(flags):
    tab = get_table("hernquist", 4, 10)
    n = 20000
    pos, vel, mass = _halo("hernquist", n, seed=5)
    coef_o, used_o, acc_o, pot_o = _oracle_all(tab, pos, vel, mass, **flags)
    c = Component(n)
    c.upload(pos, vel, mass)
    s = Sphere(tab, **flags)
    coef = s.determine_coefficients(c)
    assert rel_rms(coef, coef_o) <= COEF_TOL
    s.set_coefficients(coef_o)
    c.zero_pot_acc()
    s.get_acceleration_and_potential(c)
    out = c.download()
    assert acc_rel_err(out["acc"], acc_o).max() <= ACC_TOL


def test_center_and_rtrunc():
    tab = get_table("hernquist", 4, 10)
    n = 20000
    pos, vel, mass = _halo("hernquist", n, seed=9)
    ctr = np.array([0.3, -0.1, 0.2])
    g = O.Grid(tab)
    S = O.ParticleSet(pos + ctr, vel, mass, center=ctr, rtrunc=5.0)
    P = O.sph_params(rmin=tab.rmin, rmax=tab.rmax)
    coef_o, used_o = O.sph_coefficients(g, P, S)
    O.sph_accel(g, P, S, coef_o)
    c = Component(n)
    c.upload(pos + ctr, vel, mass)
    c.set_center(ctr, rtrunc=5.0)
    s = Sphere(tab)
    coef = s.determine_coefficients(c)
    assert s.used == used_o
    assert rel_rms(coef, coef_o) <= COEF_TOL
    s.set_coefficients(coef_o)
    c.zero_pot_acc()
    s.get_acceleration_and_potential(c)
    out = c.download()
    assert acc_rel_err(out["acc"], S.acc).max() <= ACC_TOL
    frozen = np.linalg.norm(pos, axis=1) > 5.0
    assert frozen.any() and np.all(out["acc"][frozen] == 0.0)


def test_leapfrog_and_step():
    """kick/drift against the oracle, and the fused single-level step against the
    sequence of separate calls: identical arithmetic, but the two components have different
    storage-order histories (the summation order inside a radial cell is the storage order),
    so they agree to rounding, and the fused step repeated on the same history is
    bit-identical."""
    tab = get_table("hernquist", 4, 10)
    n = 30000
    pos, vel, mass = _halo("hernquist", n, seed=13)
    dt = 0.01
    c = Component(n)
    c.upload(pos, vel, mass)
    s = Sphere(tab)
    s.determine_coefficients(c, to_host=False)
    s.set_coefficients(None)
    c.zero_pot_acc()
    s.get_acceleration_and_potential(c)
    st = c.download()

    S = O.ParticleSet(pos, vel, mass)
    S.acc[:] = st["acc"]
    O.incr_velocity(S, 0.5 * dt)
    O.incr_position(S, dt)
    c.incr_velocity(0.5 * dt)
    c.incr_position(dt)
    a = c.download()
    assert np.abs(a["vel"] - S.vel).max() <= 1e-15 * np.abs(S.vel).max()
    assert np.abs(a["pos"] - S.pos).max() <= 1e-15 * np.abs(S.pos).max()

    # separate calls for the rest of the step
    s.determine_coefficients(c, to_host=False)
    s.set_coefficients(None)
    c.zero_pot_acc()
    s.get_acceleration_and_potential(c)
    c.incr_velocity(0.5 * dt)
    sep = c.download()

    c2 = Component(n)
    c2.upload(pos, vel, mass)
    c2.set_acc_pot(st["acc"], st["pot"])
    s2 = Sphere(tab)
    s2.step(c2, dt)
    fused = c2.download()
    for k in ("pos", "vel", "acc", "pot"):
        assert np.abs(fused[k] - sep[k]).max() <= 1e-12 * np.abs(sep[k]).max(), k
    assert fused["pos"].tobytes() == sep["pos"].tobytes()      # no force involved: exact

    c3 = Component(n)
    c3.upload(pos, vel, mass)
    c3.set_acc_pot(st["acc"], st["pot"])
    s3 = Sphere(tab)
    s3.step(c3, dt)
    again = c3.download()
    for k in ("pos", "vel", "acc", "pot"):
        assert fused[k].tobytes() == again[k].tobytes(), k

    # several steps in one call (first kick + drift of step k+1 fused into the force pass of
    # step k) == the same number of single-step calls, bit for bit
    for _ in range(3):
        s2.step(c2, dt)
    s3.steps(c3, dt, 3)
    one, many = c2.download(), c3.download()
    for k in ("pos", "vel", "acc", "pot"):
        assert one[k].tobytes() == many[k].tobytes(), k


def test_empty_and_tiny():
    tab = get_table("hernquist", 2, 10)
    s = Sphere(tab)
    for n in (1, 2, 31, 33):
        pos, vel, mass = _halo("hernquist", n, seed=n)
        coef_o, used_o, acc_o, pot_o = _oracle_all(tab, pos, vel, mass)
        c = Component(64)
        c.upload(pos, vel, mass)
        coef = s.determine_coefficients(c)
        assert s.used == used_o
        assert rel_rms(coef, coef_o) <= COEF_TOL
        s.set_coefficients(coef_o)
        c.zero_pot_acc()
        s.get_acceleration_and_potential(c)
        out = c.download()
        assert acc_rel_err(out["acc"], acc_o).max() <= ACC_TOL
    c = Component(16)
    c.upload(np.zeros((0, 3)), np.zeros((0, 3)), np.zeros(0))
    coef = s.determine_coefficients(c)
    assert np.all(coef == 0.0) and s.used == 0


@pytest.mark.parametrize("n", [127, 128, 129, 255, 257, 1000, 4099])
def test_force_tile_edges(n):
    """The pipelined force kernel walks tiles of 128 sorted particles: counts around the tile
    size, a sparse set (every tile spans more table rows than a stage holds, so the lanes read
    the table through L2) and the fused multi-step driver on the same sets."""
    tab = get_table("hernquist", 4, 10)
    pos, vel, mass = _halo("hernquist", n, seed=100 + n)
    pos[: max(1, n // 50)] *= 200.0              # a dead-cell tail beyond rmax
    coef_o, used_o, acc_o, pot_o = _oracle_all(tab, pos, vel, mass)
    c = Component(n)
    c.upload(pos, vel, mass)
    s = Sphere(tab)
    coef = s.determine_coefficients(c)
    assert s.used == used_o and rel_rms(coef, coef_o) <= COEF_TOL
    s.set_coefficients(coef_o)
    c.zero_pot_acc()
    s.get_acceleration_and_potential(c)
    out = c.download()
    assert acc_rel_err(out["acc"], acc_o).max() <= ACC_TOL
    assert np.abs(out["pot"] - pot_o).max() <= ACC_TOL * np.abs(pot_o).max()
    # fused steps == separate steps, bit for bit
    dt = 1e-3
    c2, c3 = Component(n), Component(n)
    s2, s3 = Sphere(tab), Sphere(tab)
    for cc, ss in ((c2, s2), (c3, s3)):
        cc.upload(pos, vel, mass)
        ss.determine_coefficients(cc, to_host=False)
        ss.set_coefficients(None)
        cc.zero_pot_acc()
        ss.get_acceleration_and_potential(cc)
    for _ in range(2):
        s2.step(c2, dt)
    s3.steps(c3, dt, 2)
    one, many = c2.download(), c3.download()
    for k in ("pos", "vel", "acc", "pot"):
        assert one[k].tobytes() == many[k].tobytes(), k


def test_force_all_dead_and_two_pass_keys():
    """every particle outside [rmin, rmax] (no coefficients, continuation force only), and a
    4000-node table: 12-bit keys take two radix passes, the sorted keys then share a buffer
    with the next step's keys and the fused driver must fall back to the one-shot kernel."""
    tab = get_table("hernquist", 2, 10)
    n = 3000
    pos, vel, mass = _halo("hernquist", n, seed=77)
    far = pos * (60.0 / np.linalg.norm(pos, axis=1))[:, None]
    coef_o, used_o, acc_o, pot_o = _oracle_all(tab, far, vel, mass)
    assert used_o == 0
    c = Component(n)
    c.upload(far, vel, mass)
    s = Sphere(tab)
    coef = s.determine_coefficients(c)
    assert s.used == 0 and not coef.any()
    ref_coef, *_ = _oracle_all(tab, pos, vel, mass)
    g, P = O.Grid(tab), O.sph_params(rmin=tab.rmin, rmax=tab.rmax)
    S = O.ParticleSet(far, vel, mass)
    O.sph_accel(g, P, S, ref_coef)
    s.set_coefficients(ref_coef)
    c.zero_pot_acc()
    s.get_acceleration_and_potential(c)
    out = c.download()
    assert acc_rel_err(out["acc"], S.acc).max() <= ACC_TOL

    tab4 = get_table("hernquist", 2, 8, numr=4000)
    n = 20000
    pos, vel, mass = _halo("hernquist", n, seed=78)
    c2, c3 = Component(n), Component(n)
    s2, s3 = Sphere(tab4), Sphere(tab4)
    for cc, ss in ((c2, s2), (c3, s3)):
        cc.upload(pos, vel, mass)
        coef = ss.determine_coefficients(cc)
        ss.set_coefficients(None)
        cc.zero_pot_acc()
        ss.get_acceleration_and_potential(cc)
    coef_o, used_o, acc_o, pot_o = _oracle_all(tab4, pos, vel, mass)
    assert rel_rms(coef, coef_o) <= COEF_TOL
    assert acc_rel_err(c2.download()["acc"], acc_o).max() <= ACC_TOL
    for _ in range(3):
        s2.step(c2, 1e-3)
    s3.steps(c3, 1e-3, 3)
    one, many = c2.download(), c3.download()
    for k in ("pos", "vel", "acc", "pot"):
        assert one[k].tobytes() == many[k].tobytes(), k
    # and the stepped state agrees with the oracle's leapfrog
    S = O.ParticleSet(pos, vel, mass)
    g, P = O.Grid(tab4), O.sph_params(rmin=tab4.rmin, rmax=tab4.rmax)
    cf, _ = O.sph_coefficients(g, P, S)
    O.sph_accel(g, P, S, cf)
    for _ in range(3):
        O.incr_velocity(S, 0.5e-3); O.incr_position(S, 1e-3)
        S.acc[:] = 0.0; S.pot[:] = 0.0
        cf, _ = O.sph_coefficients(g, P, S)
        O.sph_accel(g, P, S, cf)
        O.incr_velocity(S, 0.5e-3)
    assert np.abs(many["pos"] - S.pos).max() <= 1e-11 * np.abs(S.pos).max()
    assert acc_rel_err(many["acc"], S.acc).max() <= ACC_TOL
