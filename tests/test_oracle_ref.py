"""Pins the CPU restatement (oracle/exp_oracle.cc, cyl_oracle.cc) to the reference's OWN code:
oracle/_ref/libexpref_path.so holds the reference's hot-path member functions, cut from
/root/reference by oracle/ref_extract.py and compiled unmodified (oracle/ref_path_wrap.cc).

Bit-exact (`tobytes()`) wherever the oracle keeps the reference's operation order; where a
reference expression runs through Eigen expression templates (covariance outer products)
the comparison is to 1e-14 of the largest entry.  Runs on CPU; the .so is built in the build
container (where /root/reference exists) and travels with the repo snapshot."""
import ctypes as C
import math

import numpy as np
import pytest

from conftest import get_table
from exp_b200 import cyltable, ics
from oracle import oracle as O
from oracle import ref_path as R

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref/libexpref_path.so not built")


def _radii(tab, n=400, seed=0):
    rng = np.random.default_rng(seed)
    r = np.exp(rng.uniform(np.log(tab.rmin * 0.5), np.log(tab.rmax * 1.5), n))
    return np.concatenate([r, [tab.rmin, tab.rmax, tab.rmin * 1e-3, tab.rmax * 10.0, 1.0]])


@pytest.mark.parametrize("cfg", [("hernquist", 4, 10, 1), ("nfw", 6, 18, 1), ("hernquist", 2, 6, 2),
                                 ("hernquist", 2, 6, 0)])
def test_slgrid_tables_bit_exact(cfg):
    """a5-a7: r_to_xi, d_xi_to_r, the grid constants, get_pot, get_force, get_dens
    (exputil/SLGridMP2.cc:711-765, 872-989, 1362-1376), all three coordinate maps."""
    kind, L, N, cmap = cfg
    import copy
    tab = copy.copy(get_table(kind, L, N, rmax=20.0 if kind == "nfw" else 50.0))
    tab.cmap = cmap                       # the table values are just numbers for this comparison
    g = O.Grid(tab)
    ref = R.RefSphere(tab)
    assert ref.consts() == (g.c.xmin, g.c.xmax, g.c.dxi)
    L_ = O.lib()
    for r in _radii(tab):
        assert ref.r_to_xi(r) == L_.orc_r_to_xi(C.byref(g.c), float(r))
        xi = ref.r_to_xi(r)
        if cmap != 1 or -1.0 <= xi < 1.0:
            assert ref.d_xi_to_r(xi) == L_.orc_d_xi_to_r(C.byref(g.c), xi)
        assert ref.get_pot(r).tobytes() == g.get_pot(r).tobytes()
        assert ref.get_force(r).tobytes() == g.get_force(r).tobytes()
        d = np.zeros((L + 1, N))
        L_.orc_slgrid_get_dens(C.byref(g.c), float(r), O._dp(d))
        assert ref.get_dens(r).tobytes() == d.tobytes()


def test_factorial_bit_exact():
    """a4: the factorial(l,m) table of src/SphericalBasis.cc:328-335"""
    for L in (2, 4, 6):
        for l in range(L + 1):
            for m in range(l + 1):
                assert R.lib().ref_factorial(L, l, m) == O.lib().orc_factorial(l, m)


def _halo(kind, n, seed, shift=(0.0, 0.0, 0.0)):
    pos, vel, mass = (ics.hernquist_halo if kind == "hernquist" else ics.nfw_halo)(n, seed=seed)
    return pos + np.asarray(shift), vel, mass


@pytest.mark.parametrize("cfg", [("hernquist", 4, 10), ("nfw", 6, 18)])
@pytest.mark.parametrize("nthreads", [1, 3])
def test_sphere_coefficients_bit_exact(cfg, nthreads):
    """a8/a9: SphericalBasis::determine_coefficients_thread + the thread sum
    (src/SphericalBasis.cc:429-599, 855-858), with centre offset, rtrunc freeze and rmin/rmax cut."""
    kind, L, N = cfg
    tab = get_table(kind, L, N, rmax=20.0 if kind == "nfw" else 50.0)
    pos, vel, mass = _halo(kind, 3000, 11, shift=(0.02, -0.01, 0.03))
    center, rtrunc = (0.02, -0.01, 0.03), 15.0
    g = O.Grid(tab)
    S = O.ParticleSet(pos, vel, mass, center=center, rtrunc=rtrunc)
    P = O.sph_params(scale=1.0, rmin=0.05, rmax=12.0)
    lst = np.arange(0, 3000, 1, dtype=np.int32)[::-1].copy()      # a non-trivial levlist order
    coef, used = O.sph_coefficients(g, P, S, lst=lst, adb=0.7, nthreads=nthreads)
    ref = R.RefSphere(tab, scale=1.0, rmin=0.05, rmax=12.0)
    out = ref.coefficients(pos, mass, lst=lst, adb=0.7, nthreads=nthreads, center=center, rtrunc=rtrunc)
    assert used == out["used"] and 0 < used < 3000
    assert coef.tobytes() == out["coef"].tobytes()


def test_sphere_scale_and_m0_only():
    tab = get_table("hernquist", 4, 10)
    pos, vel, mass = _halo("hernquist", 1500, 5)
    g = O.Grid(tab)
    S = O.ParticleSet(pos, vel, mass)
    P = O.sph_params(scale=1.7, rmin=tab.rmin, rmax=tab.rmax, M0_only=True)
    coef, used = O.sph_coefficients(g, P, S)
    out = R.RefSphere(tab, scale=1.7, M0_only=True).coefficients(pos, mass)
    assert used == out["used"]
    assert coef.tobytes() == out["coef"].tobytes()


@pytest.mark.parametrize("flags", [dict(), dict(NO_L0=True), dict(NO_L1=True), dict(EVEN_L=True),
                                   dict(EVEN_M=True), dict(M0_only=True)])
def test_sphere_accel_bit_exact(flags):
    """a11: SphericalBasis::determine_acceleration_and_potential_thread (src/SphericalBasis.cc:
    1476-1660) incl. get_pot_coefs_safe, Sphere::get_dpotl, the r > rmax continuation and the
    NO_L0/NO_L1/EVEN_L/EVEN_M/M0_only switches."""
    tab = get_table("hernquist", 4, 10)
    pos, vel, mass = _halo("hernquist", 2500, 21, shift=(0.01, 0.0, -0.02))
    pos[:40] *= 30.0                                        # beyond rmax
    pos[40] = np.array([0.01, 0.0, 0.7 - 0.02])             # on the pole of the shifted frame
    center = (0.01, 0.0, -0.02)
    g = O.Grid(tab)
    P = O.sph_params(scale=1.0, rmin=tab.rmin, rmax=10.0, **flags)
    S = O.ParticleSet(pos, vel, mass, center=center, rtrunc=400.0)
    coef, _ = O.sph_coefficients(g, O.sph_params(rmin=tab.rmin, rmax=10.0), S)
    O.sph_accel(g, P, S, coef, nthreads=2)
    ref = R.RefSphere(tab, scale=1.0, rmin=tab.rmin, rmax=10.0, **flags)
    acc, pot = ref.accel(pos, coef, nthreads=2, center=center, rtrunc=400.0)
    assert S.acc.tobytes() == acc.tobytes()
    assert S.pot.tobytes() == pot.tobytes()
    assert np.abs(acc).max() > 0


def test_sphere_accel_external_component_uses_owner_centre():
    """use_external: the other component's particles are taken Inertial and converted with the
    OWNING component's centre (src/SphericalBasis.cc:1524-1540); only freeze() uses theirs."""
    tab = get_table("hernquist", 4, 10)
    pos, vel, mass = _halo("hernquist", 1200, 31)
    g = O.Grid(tab)
    P = O.sph_params(rmin=tab.rmin, rmax=tab.rmax)
    S0 = O.ParticleSet(pos, vel, mass)
    coef, _ = O.sph_coefficients(g, P, S0)
    own, theirs = (0.05, 0.02, -0.03), (0.4, 0.0, 0.0)
    ref = R.RefSphere(tab)
    acc, pot = ref.accel(pos, coef, center=theirs, rtrunc=3.0, own_center=own)
    # restatement: evaluate about the owner's centre, freeze about theirs
    r2 = ((pos - np.asarray(theirs)) ** 2).sum(axis=1)
    keep = np.nonzero(r2 <= 9.0)[0].astype(np.int32)
    S = O.ParticleSet(pos, vel, mass, center=own)
    O.sph_accel(g, P, S, coef, lst=keep)
    assert S.acc.tobytes() == acc.tobytes() and S.pot.tobytes() == pot.tobytes()


def test_sphere_covariance_pcavar():
    """a10: the pcavar/pcaeof branches (src/SphericalBasis.cc:501-509, 525-545, 563-581)"""
    tab = get_table("hernquist", 2, 6) if False else get_table("hernquist", 4, 10)
    pos, vel, mass = _halo("hernquist", 800, 41)
    g = O.Grid(tab)
    S = O.ParticleSet(pos, vel, mass)
    P = O.sph_params(rmin=tab.rmin, rmax=tab.rmax)
    sampT = 7
    o = O.sph_coefficients_cov(g, P, S, sampT, adb=0.9)
    r = R.RefSphere(tab).coefficients(pos, mass, adb=0.9, sampT=sampT, pcaeof=True)
    assert o["used"] == r["used"]
    assert o["coef"].tobytes() == r["coef"].tobytes()
    assert o["massT"].tobytes() == r["massT"].tobytes()
    assert np.array_equal(o["countT"], r["countT"])
    assert o["muse"] == r["muse"]
    assert np.abs(r["coefM"].imag).max() == 0.0
    sT = np.abs(r["coefT"]).max()
    assert np.abs(o["coefT"] - r["coefT"]).max() <= 1e-15 * sT
    sM = np.abs(r["coefM"]).max()
    assert np.abs(o["coefM"] - r["coefM"].real).max() <= 1e-15 * sM
    # pcaeof tvar = sum over bins of the m=0-style products wk wk / mass (no facL^2 for m>0)
    assert r["tvar"] is not None and np.abs(r["tvar"]).max() > 0


# ---- cylinder ---------------------------------------------------------------------------
def _cyl(numx=32, numy=16, **kw):
    tab = cyltable.synthetic_cyl_table(mmax=4, norder=6, numx=numx, numy=numy, **kw)
    g = O.CylGridPy(tab)
    consts = (g.c.Rtable, g.c.XMIN, g.c.dX, g.c.YMIN, g.c.dY)
    return tab, g, consts


@pytest.mark.parametrize("cmap", [(1, 1), (0, 0), (1, 2)])
def test_cylinder_maps_and_get_pot_bit_exact(cmap):
    """a14: EmpCylSL::r_to_xi / z_to_y / get_pot (exputil/EmpCylSL.cc:6446-6463, 7109-7117,
    5557-5632)"""
    tab, g, consts = _cyl(cmapr=cmap[0], cmapz=cmap[1])
    ref = R.RefCylinder(tab, consts)
    L = O.lib()
    L.orc_cyl_r_to_xi.restype = C.c_double
    L.orc_cyl_r_to_xi.argtypes = [C.POINTER(O.CylGrid), C.c_double]
    L.orc_cyl_z_to_y.restype = C.c_double
    L.orc_cyl_z_to_y.argtypes = [C.POINTER(O.CylGrid), C.c_double]
    rng = np.random.default_rng(3)
    M1, N = tab.mmax + 1, tab.norder
    for _ in range(300):
        r = float(np.exp(rng.uniform(np.log(1e-5), np.log(0.5))))
        z = float(rng.normal() * 0.004) if rng.uniform() < 0.8 else float(rng.normal() * 0.5)
        assert ref.r_to_xi(r) == L.orc_cyl_r_to_xi(C.byref(g.c), r)
        assert ref.z_to_y(z) == L.orc_cyl_z_to_y(C.byref(g.c), z)
        Vc, Vs = np.zeros((M1, N)), np.zeros((M1, N))
        L.orc_cyl_get_pot(C.byref(g.c), r, z, O._dp(Vc), O._dp(Vs))
        rc, rs = ref.get_pot(r, z)
        assert Vc.tobytes() == rc.tobytes() and Vs.tobytes() == rs.tobytes()


def test_cylinder_accumulate_bit_exact():
    """a15/a16: EmpCylSL::accumulate (exputil/EmpCylSL.cc:4049-4146) under the gates of
    Cylinder::determine_coefficients_thread (src/Cylinder.cc:823-890), incl. the cylmass1 quirk"""
    tab, g, consts = _cyl()
    pos, vel, mass = ics.exponential_disk(2000, seed=44)
    pos[:25] *= 40.0                                     # off the table
    P = O.cyl_params(rcylmax=1e6, acyl=0.01)
    S = O.ParticleSet(pos, vel, mass)
    cs, sn, used, cylmass = O.cyl_coefficients(g, P, S, adb=0.8)
    ref = R.RefCylinder(tab, consts)
    out = ref.accumulate(pos, mass * 0.8)
    assert used >= out["used"] > 0
    assert cs.tobytes() == out["cos"].tobytes()
    assert sn.tobytes() == out["sin"].tobytes()
    # cylmass1 is incremented inside the m loop (x (MMAX+1)), exputil/EmpCylSL.cc:4143
    inside = np.sqrt((pos ** 2).sum(axis=1)) / tab.ascale <= g.c.Rtable
    assert math.isclose(out["cylmass1"], (tab.mmax + 1) * (mass[inside] * 0.8).sum(), rel_tol=1e-12)


def test_cylinder_covariance_covar():
    """a15: the covar branch (exputil/EmpCylSL.cc:4123-4141); the reference evaluates it through
    Eigen expressions, so agreement is to rounding"""
    tab, g, consts = _cyl()
    pos, vel, mass = ics.exponential_disk(600, seed=45)
    P = O.cyl_params(rcylmax=1e6, acyl=0.01)
    S = O.ParticleSet(pos, vel, mass)
    sampT = 5
    o = O.cyl_coefficients_cov(g, P, S, sampT)
    r = R.RefCylinder(tab, consts).accumulate(pos, mass, sampT=sampT)
    assert o["cos"].tobytes() == r["cos"].tobytes() and o["sin"].tobytes() == r["sin"].tobytes()
    assert np.array_equal(o["numbT"], r["numbT"])
    assert o["massT"].tobytes() == r["massT"].tobytes()
    assert np.abs(o["VC"] - r["VC"]).max() <= 1e-14 * np.abs(r["VC"]).max()
    assert np.abs(o["MV"] - r["MV"]).max() <= 1e-14 * np.abs(r["MV"]).max()


@pytest.mark.parametrize("flags", [dict(), dict(EVEN_M=True), dict(MMIN=1, MLIM=3), dict(NMIN=1, NLIM=4)])
def test_cylinder_accumulated_eval_bit_exact(flags):
    """a18: EmpCylSL::accumulated_eval (exputil/EmpCylSL.cc:5256-5410)"""
    tab = cyltable.synthetic_cyl_table(mmax=4, norder=6, numx=32, numy=16)
    g = O.CylGridPy(tab, **flags)
    consts = (g.c.Rtable, g.c.XMIN, g.c.dX, g.c.YMIN, g.c.dY)
    ref = R.RefCylinder(tab, consts, **flags)
    rng = np.random.default_rng(9)
    cos, sin = rng.normal(size=(5, 6)), rng.normal(size=(5, 6))
    sin[0] = 0.0
    L = O.lib()
    L.orc_cyl_accumulated_eval.argtypes = [C.POINTER(O.CylGrid), O.c_dp, O.c_dp, C.c_double, C.c_double,
                                           C.c_double] + [O.c_dp] * 5
    for _ in range(300):
        r = float(np.exp(rng.uniform(np.log(1e-5), np.log(0.3))))
        z = float(rng.normal() * 0.004)
        phi = float(rng.uniform(-np.pi, np.pi))
        o = np.zeros(5)
        L.orc_cyl_accumulated_eval(C.byref(g.c), O._dp(cos), O._dp(sin), r, z, phi,
                                   *[o[k:].ctypes.data_as(O.c_dp) for k in range(5)])
        assert tuple(o) == ref.accumulated_eval(cos, sin, r, z, phi)


# ---- leapfrog and level assignment ----------------------------------------------------------
@pytest.mark.parametrize("mlevel", [-1, 0, 2])
def test_leapfrog_bit_exact(mlevel):
    """a21: incr_position_thread / incr_velocity_thread (src/incpos.cc:15-70, src/incvel.cc:15-89)"""
    rng = np.random.default_rng(2)
    n = 5000
    pos, vel, acc = rng.normal(size=(n, 3)), rng.normal(size=(n, 3)), rng.normal(size=(n, 3))
    level = rng.integers(0, 4, n).astype(np.uint32)
    for dim in (3, 2):
        S = O.ParticleSet(pos, vel, None, level=level)
        S.acc[:] = acc
        lst = None if mlevel < 0 else S.levlist(mlevel)
        O.incr_velocity(S, 0.0123, lst=lst, dim=dim)
        v_ref = R.incr_velocity(pos, vel, acc, 0.0123, level=level, nlev=4, mlevel=mlevel, dim=dim, nthreads=3)
        assert S.vel.tobytes() == v_ref.tobytes()
        O.incr_position(S, 0.0456, lst=lst, dim=dim)
        p_ref = R.incr_position(pos, v_ref, 0.0456, level=level, nlev=4, mlevel=mlevel, dim=dim, nthreads=3)
        assert S.pos.tobytes() == p_ref.tobytes()


@pytest.mark.parametrize("kw", [dict(), dict(shiftlevl=1), dict(mfirst_mdrft=2), dict(dynfracS=0.5, with_scale=True),
                                dict(noswitch=True), dict(dim=2)])
def test_level_assignment_bit_exact(kw):
    """a22: adjust_multistep_level_thread (src/multistep.cc:52-236): dtreq through float, the
    floor(log(dtime/dtreq)/log 2) level incl. exact powers of two, shiftlevl, mfirst, offgrid counts"""
    kw = dict(kw)
    with_scale = kw.pop("with_scale", False)
    rng = np.random.default_rng(7)
    n = 20000
    vel = rng.normal(size=(n, 3)) * np.exp(rng.uniform(-6, 2, n))[:, None]
    acc = rng.normal(size=(n, 3)) * np.exp(rng.uniform(-6, 6, n))[:, None]
    pot = -np.exp(rng.uniform(-8, 2, n))
    potext = rng.normal(size=n) * 0.01
    scale = np.exp(rng.uniform(-3, 1, n)).astype(np.float32) if with_scale else None
    level = rng.integers(0, 5, n).astype(np.uint32)
    dtime, multistep = 0.01, 4
    # plant exact powers of two: dtv = dynfracV*sqrt(v2/(a2+eps)) is hard to hit, so make the float
    # dtreq land on dtime/2^k through dtd (dynfracD small) for a few particles in a second pass
    S = O.ParticleSet(np.zeros((n, 3)), vel, None, level=level, scale=scale)
    S.acc[:] = acc; S.pot[:] = pot; S.potext[:] = potext
    # the wrapper evaluates at mstep == 0 with DTreset on: NoSwitch components reset dtreq first
    P = O.level_params(dtime=dtime, multistep=multistep, reset_dtreq=bool(kw.get("noswitch")),
                       **{k: v for k, v in kw.items()})
    nsw, which, offlo, offhi = O.adjust_levels(S, P)
    lev, dq, nsw_r, lo_r, hi_r = R.adjust_levels(vel, acc, pot, potext, scale, level, np.zeros(n, np.float32),
                                                 dtime, multistep, nthreads=3, **kw)
    assert np.array_equal(S.level, lev)
    assert S.dtreq.tobytes() == dq.tobytes()
    assert (nsw, offlo, offhi) == (nsw_r, lo_r, hi_r)
    assert len(np.unique(lev)) >= 3


def test_level_assignment_powers_of_two():
    """dtreq == dtime / 2^k exactly (and one float ulp either side): libm's floor(log(x)/log(2))"""
    multistep, dtime = 8, 1.0
    vals = []
    for k in range(0, 9):
        t = np.float32(dtime / 2 ** k)
        vals += [t, np.nextafter(t, np.float32(0)), np.nextafter(t, np.float32(2))]
    vals = np.array(vals, dtype=np.float64)
    n = len(vals)
    # dts = dynfracS * scale / |sqrt(v2) + eps| with v = 1 and dynfracS chosen so that dts = vals
    vel = np.zeros((n, 3)); vel[:, 0] = 1.0
    acc = np.zeros((n, 3))
    pot = np.zeros(n); potext = np.zeros(n)
    scale = (vals * (1.0 + 1e-10)).astype(np.float32)
    level = np.zeros(n, dtype=np.uint32)
    S = O.ParticleSet(np.zeros((n, 3)), vel, None, level=level, scale=scale)
    P = O.level_params(dtime=dtime, multistep=multistep, dynfracV=1e30, dynfracD=1e30)
    O.adjust_levels(S, P)
    lev, dq, *_ = R.adjust_levels(vel, acc, pot, potext, scale, level, np.zeros(n, np.float32), dtime,
                                  multistep, dynfracV=1e30, dynfracD=1e30)
    assert np.array_equal(S.level, lev) and S.dtreq.tobytes() == dq.tobytes()
    assert len(np.unique(lev)) >= 8


def test_full_step_bit_exact():
    """do_step's single-level sequence driven with the reference's own thread functions
    (RefSphereStepper, the CPU baseline bench.py times) == the same sequence with the oracle"""
    tab = get_table("nfw", 6, 18, rmax=20.0)
    pos, vel, mass = _halo("nfw", 20000, 77)
    st = R.RefSphereStepper(tab, pos, vel, mass, nthreads=3)
    g = O.Grid(tab)
    P = O.sph_params(rmin=tab.rmin, rmax=tab.rmax)
    S = O.ParticleSet(pos, vel, mass)
    st.force()
    coef, _ = O.sph_coefficients(g, P, S, nthreads=3)
    O.sph_accel(g, P, S, coef, nthreads=3)
    for _ in range(2):
        st.step(1e-3)
        O.incr_velocity(S, 0.5e-3)
        O.incr_position(S, 1e-3)
        coef, _ = O.sph_coefficients(g, P, S, nthreads=3)
        S.acc[:] = 0.0
        S.pot[:] = 0.0
        O.sph_accel(g, P, S, coef, nthreads=3)
        O.incr_velocity(S, 0.5e-3)
    o = st.read()
    for k, v in (("pos", S.pos), ("vel", S.vel), ("acc", S.acc), ("pot", S.pot), ("coef", coef)):
        assert o[k].tobytes() == v.tobytes(), k


# ---- pyEXP twin: BasisClasses::Spherical (expui/BiorthBasis.cc) ---------------------------------
@pytest.mark.parametrize("flags,scale", [({}, 1.0), (dict(NO_L0=True), 1.0), (dict(NO_L1=True, EVEN_M=True), 0.5),
                                         (dict(EVEN_L=True), 1.0), (dict(M0_only=True), 2.0)])
def test_pyexp_twin_against_reference(flags, scale):
    """oracle/pyexp_oracle.py (accumulate, sph/crt_eval) against the reference's own
    Spherical::accumulate / sph_eval / cyl_eval / crt_eval / computeAccel and BiorthBasis::legendre_R
    (expui/BiorthBasis.cc:583-651+660-665, 711-958, 1034-1122) compiled in place.  The restatement sums
    with numpy dot products and forms the Y_lm prefactor as sqrt((2l+1)/4pi) exp(..)/2 instead of
    sqrt((0.5l+0.25)/pi exp(..)): agreement is to rounding (1e-13), not bit for bit."""
    from oracle import pyexp_oracle as PO
    tab = get_table("hernquist", 4, 8)
    n = 2500
    pos, vel, mass = ics.hernquist_halo(n, seed=5)
    pos = pos * scale + np.array([0.01, -0.02, 0.005])
    rmin, rmax = tab.rmin * scale * 3.0, tab.rmax * scale * 0.5      # both cuts bite
    rs = R.RefSphere(tab)
    rp = R.RefPyexp(rs, scale=scale, rmin=rmin, rmax=rmax, **flags)
    g = O.Grid(tab)
    c_ref, used = rp.accumulate(mass, pos)
    c_or, used_or = PO.accumulate(g, mass, pos, scale=scale, rmin=rmin, rmax=rmax)
    assert used == used_or and 0 < used < n
    assert np.abs(c_ref - c_or).max() <= 1e-13 * np.abs(c_ref).max()
    rp.set_coefs(c_ref)
    rng = np.random.default_rng(2)
    pts = rng.normal(size=(60, 3)) * np.array([1.0, 1.0, 0.6]) * rng.uniform(0.05, 5.0, size=(60, 1)) * scale
    f_ref = rp.eval("crt", pts)
    f_or = np.array([PO.crt_eval(g, c_ref, *p, scale=scale, **flags) for p in pts])
    for k in range(9):
        assert np.abs(f_ref[:, k] - f_or[:, k]).max() <= 1e-12 * max(np.abs(f_ref[:, k]).max(), 1e-300), k
    # sph_eval directly, and computeAccel against the force columns of crt_eval
    r = np.linalg.norm(pts, axis=1)
    sph = np.stack([r, pts[:, 2] / r, np.arctan2(pts[:, 1], pts[:, 0])], axis=1)
    s_ref = rp.eval("sph", sph)
    s_or = np.array([PO.sph_eval(g, c_ref, *q, scale=scale, **flags) for q in sph])
    assert np.abs(s_ref - s_or).max() <= 1e-12 * np.abs(s_ref).max()
    acc = rp.accel(pts)
    assert np.abs(acc - f_ref[:, 6:9]).max() <= 1e-11 * np.abs(acc).max()


# ---- Cylinder thread functions (src/Cylinder.cc) ------------------------------------------------
def _rot(th=0.3, ph=0.2):
    cz, sz, cx, sx = math.cos(th), math.sin(th), math.cos(ph), math.sin(ph)
    A = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1.0]]) @ np.array([[1.0, 0, 0], [0, cx, -sx], [0, sx, cx]])
    return np.ascontiguousarray(A), np.ascontiguousarray(A.T)


@pytest.mark.parametrize("rotate", [False, True])
def test_cylinder_threads_bit_exact(rotate):
    """a16/a19: Cylinder::determine_coefficients_thread and determine_acceleration_and_potential_thread
    (src/Cylinder.cc:748-896, 1266-1446) -- centre, rtrunc and rcylmax gates, adiabatic factor, body-frame
    rotation, erf blend to the monopole, the force rotated back, levels >= mlevel -- against the oracle's
    restatement, bit for bit"""
    from test_emul_cyl import disk
    tab, g, consts = _cyl()
    ref = R.RefCylinder(tab, consts)
    n = 6000
    pos, vel, mass = disk(n)
    rng = np.random.default_rng(12)
    pos[: n // 6] *= rng.uniform(5.0, 40.0, size=(n // 6, 1))        # beyond rcylmax, into the blend and past the table
    ctr = np.array([0.002, -0.001, 0.0005])
    pos = pos + ctr
    lev = rng.integers(0, 3, n).astype(np.uint32)
    body, orig = _rot() if rotate else (None, None)
    kw = dict(rcylmax=6.0, acyl=tab.ascale)
    P = O.cyl_params(body=body, orig=orig, **kw)
    S = O.ParticleSet(pos, vel, mass, center=ctr, rtrunc=0.2, level=lev)
    for M in (0, 2):
        cs, sn, used, cm = O.cyl_coefficients(g, P, S, lst=S.levlist(M), adb=0.7)
        rc, rs, rused, rcm = ref.coef_thread(pos, mass, level=lev, nlev=3, mlevel=M, center=ctr, rtrunc=0.2,
                                             adb=0.7, body=body, orig=orig, **kw)
        assert used == rused and 0 < used < len(S.levlist(M))
        assert cs.tobytes() == rc.tobytes() and sn[1:].tobytes() == rs[1:].tobytes()
        assert cm == rcm
    cs, sn, used, cm = O.cyl_coefficients(g, P, S)
    act = np.nonzero(lev >= 1)[0].astype(np.int32)
    S.acc[:] = 0.01
    S.pot[:] = -0.5
    O.cyl_accel(g, P, S, cs, sn, cm, lst=act)
    racc, rpot = ref.accel_thread(pos, cs, sn, cm, level=lev, nlev=3, mlevel=1, center=ctr, rtrunc=0.2,
                                  body=body, orig=orig, acc=np.full((n, 3), 0.01), pot=np.full(n, -0.5))
    assert S.acc.tobytes() == racc.tobytes() and S.pot.tobytes() == rpot.tobytes()
    assert np.all(racc[lev < 1] == 0.01)
    # use_external: the origin and the orientation are the owning component's, the particles another's
    own = ctr + np.array([0.01, 0.02, -0.005])
    S2 = O.ParticleSet(pos, vel, mass, center=own, rtrunc=0.2, level=lev)
    O.cyl_accel(g, P, S2, cs, sn, cm, lst=act)
    racc2, rpot2 = ref.accel_thread(pos, cs, sn, cm, level=lev, nlev=3, mlevel=1, center=ctr, rtrunc=0.2,
                                    owner_center=own, body=body, orig=orig)
    assert S2.acc.tobytes() == racc2.tobytes() and S2.pot.tobytes() == rpot2.tobytes()


@pytest.mark.parametrize("mlevel,dim", [(0, 3), (1, 3), (0, 2)])
def test_fix_positions_bit_exact(mlevel, dim):
    """f1: fix_positions_thread (src/Component.cc:3280-3352) -- per-level mass, m*pos, m*vel, m*acc of the
    non-frozen particles of the levels >= mlevel -- against the oracle, bit for bit"""
    n = 8000
    pos, vel, mass = ics.hernquist_halo(n, seed=21)
    rng = np.random.default_rng(4)
    acc = rng.normal(size=(n, 3))
    lev = rng.integers(0, 4, n).astype(np.uint32)
    ctr = np.array([0.05, -0.02, 0.01])
    S = O.ParticleSet(pos, vel, mass, center=ctr, rtrunc=2.0, level=lev)
    S.acc[:] = acc
    sums, _ = O.fix_positions(S, mlevel, 3, dim=dim)
    ref = R.fix_positions(pos, vel, acc, mass, lev, 4, mlevel=mlevel, dim=dim, center=ctr, rtrunc=2.0)
    assert sums.tobytes() == ref.tobytes()
    assert np.all(ref[:mlevel] == 0.0) and ref[mlevel:, 0].sum() < mass[lev >= mlevel].sum()   # rtrunc bites


def test_sphere_multistep_update_bit_exact():
    """a13: SphericalBasis::multistep_update (src/SphericalBasis.cc:1156-1228): the switching particle's
    contribution leaves differ[old level] and enters differ[new level] (freeze and rmax gates, adiabatic
    factor, scale) -- oracle against the reference's own code, bit for bit"""
    tab = get_table("hernquist", 4, 8)
    n = 3000
    pos, vel, mass = ics.hernquist_halo(n, seed=23)
    ctr = np.array([0.01, -0.02, 0.005])
    pos = pos * 0.5 + ctr
    rng = np.random.default_rng(6)
    idx = rng.choice(n, 400, replace=False).astype(np.int32)
    frm = rng.integers(0, 4, 400).astype(np.int32)
    to = (frm + rng.integers(1, 4, 400)).astype(np.int32) % 4
    rmax = tab.rmax * 0.5 * 0.3                     # the r < rmax gate bites
    ref = R.RefSphere(tab, scale=0.5, rmin=tab.rmin * 0.5, rmax=rmax)
    want = ref.multistep_update(pos, mass, idx, frm, to, 4, center=ctr, rtrunc=1.5, adb=0.8)
    g = O.Grid(tab)
    P = O.sph_params(scale=0.5, rmin=tab.rmin * 0.5, rmax=rmax)
    S = O.ParticleSet(pos, vel, mass, center=ctr, rtrunc=1.5)
    d = np.zeros((4, (tab.lmax + 1) ** 2, tab.nmax))
    for i, a, b in zip(idx, frm, to):
        O.sph_multistep_update(g, P, S, int(i), int(a), int(b), d, adb=0.8)
    assert d.tobytes() == want.tobytes()
    assert np.abs(want).max() > 0 and np.abs(want.sum(axis=0)).max() <= 1e-12 * np.abs(want).max()


def test_cylinder_multistep_update_bit_exact():
    """a20: Cylinder::multistep_update -> CylEXP::multistep_update (src/Cylinder.cc:1752-1773,
    src/CylEXP.cc:159-188): the switching particle's contribution moves from differ[old level] to
    differ[new level] (freeze and table-radius gates, adiabatic factor) -- oracle against the reference's
    own code, bit for bit"""
    from test_emul_cyl import disk
    tab, g, consts = _cyl()
    ref = R.RefCylinder(tab, consts)
    n = 4000
    pos, vel, mass = disk(n)
    rng = np.random.default_rng(14)
    pos[: n // 8] *= rng.uniform(5.0, 60.0, size=(n // 8, 1))          # some beyond the table radius
    ctr = np.array([0.002, -0.001, 0.0005])
    pos = pos + ctr
    idx = rng.choice(n, 500, replace=False).astype(np.int32)
    frm = rng.integers(0, 3, 500).astype(np.int32)
    to = (frm + rng.integers(1, 3, 500)).astype(np.int32) % 3
    wC, wS = ref.multistep_update(pos, mass, idx, frm, to, 3, center=ctr, rtrunc=0.3, adb=0.9)
    S = O.ParticleSet(pos, vel, mass, center=ctr, rtrunc=0.3)
    dC, dS = np.zeros_like(wC), np.zeros_like(wS)
    for i, a, b in zip(idx, frm, to):
        O.cyl_multistep_update(g, S, int(i), int(a), int(b), dC, dS, adb=0.9)
    assert dC.tobytes() == wC.tobytes() and dS.tobytes() == wS.tobytes()
    assert np.abs(wC).max() > 0 and np.abs(wC.sum(axis=0)).max() <= 1e-12 * np.abs(wC).max()


@pytest.mark.parametrize("multistep", [1, 3])
def test_compute_multistep_coefficients_bit_exact(multistep):
    """a12: SphericalBasis::compute_multistep_coefficients (src/SphericalBasis.cc:1231-1474) at every drift
    sub-step of the hierarchy, with the reference's own code fed the oracle's initialize_multistep tables"""
    tab = get_table("hernquist", 3, 6)
    ref = R.RefSphere(tab)
    _, mfirst, dL, dN = O.multistep_tables(multistep)
    rng = np.random.default_rng(7)
    shape = (multistep + 1, (tab.lmax + 1) ** 2, tab.nmax)
    coefL, coefN = rng.normal(size=shape), rng.normal(size=shape)
    for mdrft in range(0, (1 << multistep) + 1):
        want = ref.compute_multistep_coefficients(mdrft, mfirst, dL, dN, coefL, coefN)
        got = O.multistep_interp(multistep, mdrft, coefL, coefN)
        assert got.reshape(want.shape).tobytes() == want.tobytes(), mdrft


def test_cylinder_compute_multistep_coefficients_bit_exact():
    """a20: CylEXP::compute_multistep_coefficients (src/CylEXP.cc:192-282) at every drift sub-step"""
    tab, g, consts = _cyl()
    ref = R.RefCylinder(tab, consts)
    multistep = 2
    _, mfirst, dL, dN = O.multistep_tables(multistep)
    rng = np.random.default_rng(8)
    shape = (multistep + 1, tab.mmax + 1, tab.norder)
    cL, cN, sL, sN = (rng.normal(size=shape) for _ in range(4))
    for a in (sL, sN):
        a[:, 0] = 0.0                       # no sine terms at m = 0
    for mdrft in range(0, (1 << multistep) + 1):
        wc, ws = ref.compute_multistep_coefficients(mdrft, mfirst, dL, dN, cL, cN, sL, sN)
        gc = O.multistep_interp(multistep, mdrft, cL, cN).reshape(wc.shape)
        gs = O.multistep_interp(multistep, mdrft, sL, sN).reshape(ws.shape)
        assert gc.tobytes() == wc.tobytes() and gs[1:].tobytes() == ws[1:].tobytes(), mdrft
