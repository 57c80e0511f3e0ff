"""GPU parity against the reference's OWN code: the CUDA library through the C ABI vs
oracle/_ref/libexpref_path.so (SphericalBasis / SLGridSph / EmpCylSL member functions compiled
from /root/reference where they lie; built in the build container, travels with the snapshot).

Tolerances are BASELINE.json's: coefficients <= 1e-10 relative RMS, accelerations and potentials
<= 1e-8 relative, level vector and dtreq bit-exact."""
import math

import numpy as np
import pytest

from conftest import acc_rel_err, get_table, rel_rms
from exp_b200 import cyltable, ics
from exp_b200.component import Component
from exp_b200.cylinder import Cylinder
from exp_b200.sphere import Sphere
from oracle import oracle as O
from oracle import ref_path as R

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not R.available(), reason="oracle/_ref/libexpref_path.so not built")]


@pytest.mark.parametrize("kind,L,N,n", [("hernquist", 4, 10, 100000), ("nfw", 6, 18, 100000)])
def test_sphere_vs_reference_code(kind, L, N, n):
    """BASELINE configs[0] (1e5, lmax=4, nmax=10) and the shape of configs[1] (lmax=6, nmax=18)"""
    tab = get_table(kind, L, N, rmax=50.0 if kind == "hernquist" else 20.0)
    gen = ics.hernquist_halo if kind == "hernquist" else ics.nfw_halo
    pos, vel, mass = gen(n, seed=17)
    pos = pos + np.array([0.01, -0.02, 0.005])
    pos[:5] *= 100.0
    ref = R.RefSphere(tab)
    rc = ref.coefficients(pos, mass, nthreads=4)
    racc, rpot = ref.accel(pos, rc["coef"], nthreads=4)

    c = Component(n)
    c.upload(pos, vel, mass)
    s = Sphere(tab)
    coef = s.determine_coefficients(c)
    assert s.used == rc["used"]
    assert rel_rms(coef, rc["coef"]) <= 1e-10
    s.set_coefficients(rc["coef"])
    c.zero_pot_acc()
    s.get_acceleration_and_potential(c)
    out = c.download()
    assert acc_rel_err(out["acc"], racc).max() <= 1e-8
    assert np.abs(out["pot"] - rpot).max() <= 1e-8 * np.abs(rpot).max()


def test_cylinder_vs_reference_code():
    tab = cyltable.synthetic_cyl_table(mmax=6, norder=18, numx=128, numy=64)
    g = O.CylGridPy(tab)
    consts = (g.c.Rtable, g.c.XMIN, g.c.dX, g.c.YMIN, g.c.dY)
    n = 20000
    pos, vel, mass = ics.exponential_disk(n, seed=44)
    ref = R.RefCylinder(tab, consts)
    rc = ref.accumulate(pos, mass)
    c = Component(n)
    c.upload(pos, vel, mass)
    cyl = Cylinder(tab, rcylmax=1e6)
    cs, sn = cyl.determine_coefficients(c)
    assert cyl.used == rc["used"]
    assert rel_rms(cs, rc["cos"]) <= 1e-10
    assert rel_rms(sn[1:], rc["sin"][1:]) <= 1e-10
    # force inside the blend radius (ratio < 0.75): pure EmpCylSL::accumulated_eval
    cyl.set_coefficients(rc["cos"], rc["sin"], 0.0)
    c.zero_pot_acc()
    cyl.get_acceleration_and_potential(c)
    out = c.download()
    rr = np.sqrt((pos ** 2).sum(axis=1))
    idx = np.nonzero(rr < 0.7 * g.c.Rtable * tab.ascale)[0][:3000]
    facc, fpot = np.zeros((len(idx), 3)), np.zeros(len(idx))
    for k, i in enumerate(idx):
        x, y, z = pos[i]
        r = math.sqrt(x * x + y * y)
        phi = math.atan2(y, x)
        p0, p, fr, fz, fp = ref.accumulated_eval(rc["cos"], rc["sin"], r, z, phi)
        # src/Cylinder.cc:1389-1391, 1400
        facc[k] = (fr * x / r - fp * y / (r * r), fr * y / r + fp * x / (r * r), fz)
        fpot[k] = p
    assert acc_rel_err(out["acc"][idx], facc).max() <= 1e-8
    assert np.abs(out["pot"][idx] - fpot).max() <= 1e-8 * np.abs(fpot).max()


def test_levels_and_leapfrog_vs_reference_code():
    rng = np.random.default_rng(5)
    n = 200000
    pos, vel, mass = ics.hernquist_halo(n, seed=9)
    acc = rng.normal(size=(n, 3)) * np.exp(rng.uniform(-4, 4, n))[:, None]
    pot = -np.exp(rng.uniform(-6, 1, n))
    c = Component(n)
    c.upload(pos, vel, mass)
    c.set_acc_pot(acc, pot)
    c.incr_velocity(0.0123)
    c.incr_position(0.0456)
    out = c.download()
    v_ref = R.incr_velocity(pos, vel, acc, 0.0123, nthreads=4)
    p_ref = R.incr_position(pos, v_ref, 0.0456, nthreads=4)
    # fused multiply-add on the device vs separate multiply and add in the reference build
    assert np.abs(out["vel"] - v_ref).max() <= 1e-15 * max(1.0, np.abs(v_ref).max())
    assert np.abs(out["pos"] - p_ref).max() <= 1e-15 * max(1.0, np.abs(p_ref).max())

    c.upload(pos, vel, mass)
    c.set_acc_pot(acc, pot)
    c.adjust_multistep_level(dtime=0.01, multistep=6, reset_dtreq=True)
    got = c.download(level=True, dtreq=True)
    lev, dq, *_ = R.adjust_levels(vel, acc, pot, np.zeros(n), None, np.zeros(n, np.uint32),
                                  np.zeros(n, np.float32), 0.01, 6, nthreads=4)
    assert np.array_equal(got["level"], lev)
    assert got["dtreq"].tobytes() == dq.tobytes()


def test_external_component_uses_owner_centre():
    """interaction forces (ComponentContainer::compute_potential, src/ComponentContainer.cc:785-823):
    another component's particles are evaluated about the OWNER's centre; freeze uses theirs"""
    tab = get_table("hernquist", 4, 10)
    n = 30000
    pos, vel, mass = ics.hernquist_halo(n, seed=23)
    own, theirs = (0.05, 0.02, -0.03), (0.4, 0.0, 0.0)
    ref = R.RefSphere(tab)
    rc = ref.coefficients(pos, mass, center=own)
    other = pos[:20000] * 0.8 + np.array([0.1, 0.0, 0.0])
    racc, rpot = ref.accel(other, rc["coef"], center=theirs, rtrunc=3.0, own_center=own, nthreads=4)

    cA, cB = Component(n), Component(len(other))
    cA.upload(pos, vel, mass)
    cA.set_center(own)
    cB.upload(other, vel[:20000], mass[:20000])
    cB.set_center(theirs, rtrunc=3.0)
    s = Sphere(tab)
    s.set_owner(cA)
    coef = s.determine_coefficients(cA)
    assert rel_rms(coef, rc["coef"]) <= 1e-10
    s.set_coefficients(rc["coef"])
    cB.zero_pot_acc()
    s.get_acceleration_and_potential(cB)
    out = cB.download()
    frozen = ((other - np.asarray(theirs)) ** 2).sum(axis=1) > 9.0
    assert frozen.any() and not frozen.all()
    assert np.all(out["acc"][frozen] == 0.0) and np.all(racc[frozen] == 0.0)
    assert acc_rel_err(out["acc"][~frozen], racc[~frozen]).max() <= 1e-8
    assert np.abs(out["pot"] - rpot).max() <= 1e-8 * np.abs(rpot).max()

    # cylinder: origin = owner's centre, no freeze test in the force pass (src/Cylinder.cc:1343-1349)
    tabC = cyltable.synthetic_cyl_table(mmax=4, norder=6, numx=48, numy=24)
    g = O.CylGridPy(tabC)
    nd = 8000
    pd, vd, md = ics.exponential_disk(nd, seed=8)
    cs = np.random.default_rng(1).normal(size=(5, 6)); sn = np.random.default_rng(2).normal(size=(5, 6)); sn[0] = 0
    S = O.ParticleSet(pd + np.array([0.003, -0.001, 0.0005]), vd, md, center=(0.003, -0.001, 0.0005))
    O.cyl_accel(g, O.cyl_params(rcylmax=20.0, acyl=tabC.ascale), S, cs, sn, 0.7)
    cD, cH = Component(nd), Component(nd)
    cD.upload(pd, vd, md)
    cD.set_center((0.003, -0.001, 0.0005))
    cH.upload(S.pos, vd, md)
    cH.set_center((0.9, 0.9, 0.9), rtrunc=1e-3)            # would freeze everything if it were used
    cyl = Cylinder(tabC)
    cyl.set_owner(cD)
    cyl.set_coefficients(cs, sn, 0.7)
    cH.zero_pot_acc()
    cyl.get_acceleration_and_potential(cH)
    out = cH.download()
    assert acc_rel_err(out["acc"], S.acc).max() <= 1e-8
    assert np.abs(out["pot"] - S.pot).max() <= 1e-8 * np.abs(S.pot).max()


def test_sphere_pcaeof_tvar_vs_reference_code():
    """a10: the pcaeof branch tvar (src/SphericalBasis.cc:536-545, 573-581)"""
    tab = get_table("hernquist", 4, 10)
    n = 20000
    pos, vel, mass = ics.hernquist_halo(n, seed=41)
    pos = pos + np.array([0.01, -0.02, 0.005])
    r = R.RefSphere(tab).coefficients(pos, mass, adb=0.9, sampT=3, pcaeof=True)
    c = Component(n)
    c.upload(pos, vel, mass)
    s = Sphere(tab)
    tvar = s.covariance_eof(c, adb=0.9)
    assert np.abs(tvar - r["tvar"]).max() <= 1e-10 * np.abs(r["tvar"]).max()
    cov = s.covariance(c, 3, adb=0.9)
    assert np.abs(cov["coefM"] - r["coefM"].real).max() <= 1e-10 * np.abs(r["coefM"]).max()
    assert np.abs(cov["coefT"] - r["coefT"]).max() <= 1e-10 * np.abs(r["coefT"]).max()
    assert np.array_equal(cov["countT"], r["countT"])


@pytest.mark.parametrize("which", ["pcavar", "pcaeof"])
def test_cylinder_pca_variants_vs_reference_code(which):
    """a15: the PCAVAR / PCAEOF branches of EmpCylSL::accumulate (exputil/EmpCylSL.cc:4089-4119)"""
    tab = cyltable.synthetic_cyl_table(mmax=4, norder=6, numx=48, numy=24)
    g = O.CylGridPy(tab)
    consts = (g.c.Rtable, g.c.XMIN, g.c.dX, g.c.YMIN, g.c.dY)
    n = 6000
    pos, vel, mass = ics.exponential_disk(n, seed=46)
    sampT = 4 if which == "pcavar" else 1
    rV, rM = R.RefCylinder(tab, consts).accumulate_pca(pos, mass, which, sampT=sampT)
    c = Component(n)
    c.upload(pos, vel, mass)
    cyl = Cylinder(tab, rcylmax=1e6)
    out = cyl.covariance_pca(c, sampT, which)
    assert np.abs(out["covM"] - rM).max() <= 1e-10 * np.abs(rM).max()
    if which == "pcavar":
        assert np.abs(out["covV"] - rV).max() <= 1e-10 * np.abs(rV).max()


@pytest.mark.parametrize("cmap", [0, 2])
def test_sphere_other_coordinate_maps_vs_reference_code(cmap):
    """cmap = 0 (xi = r) and 2 (xi = log r) of SLGridSph::r_to_xi / d_xi_to_r
    (exputil/SLGridMP2.cc:711-765); the table values are those of the cmap = 1 solve re-read on the
    other grid, which is all the interpolation cares about"""
    import copy
    tab = copy.copy(get_table("hernquist", 4, 10))
    tab.cmap = cmap
    n = 30000
    pos, vel, mass = ics.hernquist_halo(n, seed=19)
    pos = pos + np.array([0.01, -0.02, 0.005])
    ref = R.RefSphere(tab)
    rc = ref.coefficients(pos, mass, nthreads=4)
    racc, rpot = ref.accel(pos, rc["coef"], nthreads=4)
    c = Component(n)
    c.upload(pos, vel, mass)
    s = Sphere(tab)
    coef = s.determine_coefficients(c)
    assert s.used == rc["used"]
    assert rel_rms(coef, rc["coef"]) <= 1e-10
    s.set_coefficients(rc["coef"])
    c.zero_pot_acc()
    s.get_acceleration_and_potential(c)
    out = c.download()
    assert acc_rel_err(out["acc"], racc).max() <= 1e-8
    assert np.abs(out["pot"] - rpot).max() <= 1e-8 * np.abs(rpot).max()


def test_level_powers_of_two_vs_reference_code():
    """dtreq == dtime / 2^k exactly and one float ulp either side, up to 2^20: the device's
    floor(log(dtime/dtreq)/log 2) must equal libm's, also where log(x)/log(2) rounds up to k for
    x just below 2^k"""
    multistep, dtime = 20, 1.0
    vals = []
    for k in range(0, 21):
        t = np.float32(dtime / 2 ** k)
        vals += [t, np.nextafter(t, np.float32(0)), np.nextafter(t, np.float32(2))]
    # doubles whose quotient dtime/dtreq sits within a few ulp below 2^k cannot arise from a float
    # dtreq with dtime = 1; use a dtime that makes the quotient inexact
    vals = np.array(vals, dtype=np.float64)
    for dt_scale in (1.0, 3.0, 0.7):
        n = len(vals)
        vel = np.zeros((n, 3)); vel[:, 0] = 1.0
        acc = np.zeros((n, 3)); pot = np.zeros(n)
        scale = (vals * dt_scale * (1.0 + 1e-10)).astype(np.float32)
        c = Component(n)
        c.upload(np.zeros((n, 3)), vel, np.ones(n), scale=scale)
        c.set_acc_pot(acc, pot)
        c.adjust_multistep_level(dtime=dtime * dt_scale, multistep=multistep, dynfracV=1e30, dynfracD=1e30,
                                 reset_dtreq=True)
        got = c.download(level=True, dtreq=True)
        lev, dq, *_ = R.adjust_levels(vel, acc, pot, np.zeros(n), scale, np.zeros(n, np.uint32),
                                      np.zeros(n, np.float32), dtime * dt_scale, multistep, dynfracV=1e30,
                                      dynfracD=1e30)
        assert np.array_equal(got["level"], lev), (dt_scale, np.nonzero(got["level"] != lev)[0][:5])
        assert got["dtreq"].tobytes() == dq.tobytes()


def test_leapfrog_dim2_vs_reference_code():
    """Component::dim = 2: only x and y advance (src/incpos.cc:63-66, src/incvel.cc:68-69)"""
    rng = np.random.default_rng(12)
    n = 50000
    pos, vel, acc = rng.normal(size=(n, 3)), rng.normal(size=(n, 3)), rng.normal(size=(n, 3))
    c = Component(n)
    c.upload(pos, vel, np.ones(n))
    c.set_acc_pot(acc, np.zeros(n))
    P = dict(dtime=0.01, multistep=3, dim=2, reset_dtreq=True)
    c.adjust_multistep_level(**P)
    got = c.download(level=True, dtreq=True)
    lev, dq, *_ = R.adjust_levels(vel, acc, np.zeros(n), np.zeros(n), None, np.zeros(n, np.uint32),
                                  np.zeros(n, np.float32), 0.01, 3, dim=2, nthreads=2)
    assert np.array_equal(got["level"], lev) and got["dtreq"].tobytes() == dq.tobytes()


def test_cylinder_default_grid_vs_reference_code():
    """BASELINE configs[2]'s table shape (NUMX = 256, NUMY = 128: 257 x 129 nodes, mmax = 6,
    norder = 18) against EmpCylSL::accumulate / accumulated_eval of the reference"""
    tab = cyltable.synthetic_cyl_table(mmax=6, norder=18, numx=256, numy=128)
    g = O.CylGridPy(tab)
    consts = (g.c.Rtable, g.c.XMIN, g.c.dX, g.c.YMIN, g.c.dY)
    n = 20000
    pos, vel, mass = ics.exponential_disk(n, seed=47)
    ref = R.RefCylinder(tab, consts)
    rc = ref.accumulate(pos, mass)
    c = Component(n)
    c.upload(pos, vel, mass)
    cyl = Cylinder(tab, rcylmax=1e6)
    cs, sn = cyl.determine_coefficients(c)
    assert cyl.used == rc["used"]
    assert rel_rms(cs, rc["cos"]) <= 1e-10 and rel_rms(sn[1:], rc["sin"][1:]) <= 1e-10
    cyl.set_coefficients(rc["cos"], rc["sin"], 0.0)
    c.zero_pot_acc()
    cyl.get_acceleration_and_potential(c)
    out = c.download()
    rr = np.sqrt((pos ** 2).sum(axis=1))
    idx = np.nonzero(rr < 0.7 * g.c.Rtable * tab.ascale)[0][:1500]
    facc = np.zeros((len(idx), 3))
    for k, i in enumerate(idx):
        x, y, z = pos[i]
        r = math.sqrt(x * x + y * y)
        p0, p, fr, fz, fp = ref.accumulated_eval(rc["cos"], rc["sin"], r, z, math.atan2(y, x))
        facc[k] = (fr * x / r - fp * y / (r * r), fr * y / r + fp * x / (r * r), fz)
    assert acc_rel_err(out["acc"][idx], facc).max() <= 1e-8


def test_pyexp_sphere_against_reference_code():
    """exp_b200.pyexp.SphericalSL (createFromArray, getFields, getAccel) against the reference's own
    BasisClasses::Spherical code compiled in place (expui/BiorthBasis.cc:583-958)"""
    from exp_b200 import pyexp
    tab = get_table("hernquist", 4, 10)
    n = 4000
    pos, vel, mass = ics.hernquist_halo(n, seed=3)
    pos = pos + np.array([0.01, -0.02, 0.005])
    rs = R.RefSphere(tab)
    rp = R.RefPyexp(rs, rmin=tab.rmin, rmax=tab.rmax)
    want, used = rp.accumulate(mass, pos)
    b = pyexp.SphericalSL(tab)
    cf = b.createFromArray(mass, pos, time=0.0)
    assert b.used == used
    L = tab.lmax
    k = 0
    for l in range(L + 1):                      # (l, m) complex packing of load_coefs (:480-516)
        for m in range(l + 1):
            row = l * l + (0 if m == 0 else 2 * m - 1)
            assert rel_rms(cf.coefs[k].real, want[row]) <= 1e-10
            if m:
                assert rel_rms(cf.coefs[k].imag, want[row + 1]) <= 1e-10
            k += 1
    rp.set_coefs(want)
    rng = np.random.default_rng(1)
    pts = rng.normal(size=(50, 3)) * np.array([1.0, 1.0, 0.6]) * rng.uniform(0.05, 5.0, size=(50, 1))
    ref = rp.eval("crt", pts)
    got = b.getFields(pts[:, 0], pts[:, 1], pts[:, 2])
    acc = b.getAccel(pts)
    racc = rp.accel(pts)
    for j in range(len(pts)):
        fn = np.linalg.norm(ref[j, 6:9])
        assert np.abs(got[3:6, j] - ref[j, 3:6]).max() <= 1e-8 * max(abs(ref[j, 5]), abs(ref[j, 3]))
        assert np.abs(got[6:9, j] - ref[j, 6:9]).max() <= 1e-8 * fn
        assert np.abs(got[0:3, j] - ref[j, 0:3]).max() <= 1e-8 * max(abs(ref[j, 2]), abs(ref[j, 0]))
        assert np.abs(acc[j] - racc[j]).max() <= 1e-8 * fn


@pytest.mark.parametrize("rotate", [False, True])
def test_cylinder_threads_vs_reference_code(rotate):
    """the GPU cylinder path against the reference's own Cylinder thread functions
    (src/Cylinder.cc:748-896, 1266-1446): centre, rtrunc and rcylmax gates, adiabatic factor, body-frame
    rotation, the erf blend to the monopole, multistep levels"""
    tab = cyltable.synthetic_cyl_table(mmax=4, norder=8, numx=64, numy=32)
    g = O.CylGridPy(tab)
    consts = (g.c.Rtable, g.c.XMIN, g.c.dX, g.c.YMIN, g.c.dY)
    ref = R.RefCylinder(tab, consts)
    n = 30000
    pos, vel, mass = ics.exponential_disk(n, seed=45)
    rng = np.random.default_rng(13)
    pos[: n // 6] *= rng.uniform(5.0, 40.0, size=(n // 6, 1))       # blend region and beyond the table
    ctr = np.array([0.002, -0.001, 0.0005])
    pos = pos + ctr
    lev = rng.integers(0, 3, n).astype(np.uint32)
    body = orig = None
    if rotate:
        th, ph = 0.3, 0.2
        A = (np.array([[math.cos(th), -math.sin(th), 0], [math.sin(th), math.cos(th), 0], [0, 0, 1.0]])
             @ np.array([[1.0, 0, 0], [0, math.cos(ph), -math.sin(ph)], [0, math.sin(ph), math.cos(ph)]]))
        body, orig = np.ascontiguousarray(A), np.ascontiguousarray(A.T)
    kw = dict(rcylmax=6.0, acyl=tab.ascale)
    c = Component(n)
    c.upload(pos, vel, mass, level=lev)
    c.set_center(ctr, rtrunc=0.2)
    cyl = Cylinder(tab, multistep=2, **kw)
    if rotate:
        cyl.set_orientation(body, orig)
    tot_c, tot_s, tot_m = 0.0, 0.0, 0.0
    for M in range(3):
        rc, rs, rused, rcm = ref.coef_thread(pos, mass, level=lev, nlev=3, mlevel=M, center=ctr, rtrunc=0.2,
                                             adb=0.7, body=body, orig=orig, **kw)
        cyl.set_multistep_level(M)
        cs, sn = cyl.determine_coefficients(c, adb=0.7)
        assert cyl.used == rused
        assert rel_rms(cs, rc) <= 1e-10 and rel_rms(sn[1:], rs[1:]) <= 1e-10
        tot_c, tot_s, tot_m = tot_c + rc, tot_s + rs, tot_m + rcm
    cyl.set_coefficients(tot_c, tot_s, tot_m)
    cyl.set_multistep_level(1)
    c.zero_pot_acc()
    cyl.get_acceleration_and_potential(c)
    out = c.download()
    racc, rpot = ref.accel_thread(pos, tot_c, tot_s, tot_m, level=lev, nlev=3, mlevel=1, center=ctr, rtrunc=0.2,
                                  body=body, orig=orig)
    act = lev >= 1
    assert acc_rel_err(out["acc"][act], racc[act]).max() <= 1e-8
    assert np.abs(out["pot"][act] - rpot[act]).max() <= 1e-8 * np.abs(rpot).max()
    assert np.all(out["acc"][~act] == 0.0)
