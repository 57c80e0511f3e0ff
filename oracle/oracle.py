"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Never imported by exp_b200/.
Pinned bit for bit to the reference's own member functions compiled in place
(oracle/_ref/libexpref_path.so, tests/test_oracle_ref.py; see exp_oracle.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_up = C.POINTER(C.c_uint)
c_fp = C.POINTER(C.c_float)
c_lp = C.POINTER(C.c_long)


def build(force: bool = False) -> None:
    """Compile liboracle.so (and _ref/ when /root/reference is present)."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith(("_oracle.cc", ".h"))]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if os.path.isdir("/root/reference"):
        # the reference's own code compiled in place (only in the build container)
        refs = [os.path.join(_HERE, "_ref", f) for f in ("libexpref_basis.so", "libexpref_path.so")]
        rsrc = [os.path.join(_HERE, f) for f in ("ref_path_wrap.cc", "ref_extract.py", "ref_basis_wrap.cc")]
        rsrc += [os.path.join(_HERE, "ref_shim", f) for f in os.listdir(os.path.join(_HERE, "ref_shim"))]
        stale = stale or any(not os.path.exists(r) for r in refs) or \
            max(os.path.getmtime(s) for s in rsrc) > min(os.path.getmtime(r) for r in refs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s"], check=True, capture_output=True)


class SLGrid(C.Structure):
    _fields_ = [("lmax", C.c_int), ("nmax", C.c_int), ("numr", C.c_int), ("cmap", C.c_int),
                ("rmin", C.c_double), ("rmax", C.c_double), ("rmap", C.c_double),
                ("ev", c_dp), ("ef", c_dp), ("p0", c_dp), ("d0", c_dp),
                ("xmin", C.c_double), ("xmax", C.c_double), ("dxi", C.c_double)]


class SphParams(C.Structure):
    _fields_ = [("scale", C.c_double), ("rmin", C.c_double), ("rmax", C.c_double),
                ("NO_L0", C.c_int), ("NO_L1", C.c_int), ("EVEN_L", C.c_int),
                ("EVEN_M", C.c_int), ("M0_only", C.c_int)]


class Particles(C.Structure):
    _fields_ = [("n", C.c_long), ("pos", c_dp), ("vel", c_dp), ("acc", c_dp), ("pot", c_dp),
                ("potext", c_dp), ("mass", c_dp), ("scale", c_dp), ("level", c_up),
                ("dtreq", c_fp), ("center", C.c_double * 3), ("rtrunc", C.c_double)]


class LevelParams(C.Structure):
    _fields_ = [("dtime", C.c_double), ("dynfracS", C.c_double), ("dynfracD", C.c_double),
                ("dynfracV", C.c_double), ("dynfracA", C.c_double), ("dynfracP", C.c_double),
                ("multistep", C.c_int), ("shiftlevl", C.c_int), ("mfirst_mdrft", C.c_int),
                ("noswitch", C.c_int), ("reset_dtreq", C.c_int), ("apply", C.c_int),
                ("dim", C.c_int)]


class CylGrid(C.Structure):
    _fields_ = [("MMAX", C.c_int), ("NORDER", C.c_int), ("NUMX", C.c_int), ("NUMY", C.c_int),
                ("CMAPR", C.c_int), ("CMAPZ", C.c_int), ("ASCALE", C.c_double), ("HSCALE", C.c_double),
                ("RMIN", C.c_double), ("RMAX", C.c_double), ("EVEN_M", C.c_int), ("MMIN", C.c_int),
                ("MLIM", C.c_int), ("NMIN", C.c_int), ("NLIM", C.c_int),
                ("potC", c_dp), ("potS", c_dp), ("rforceC", c_dp), ("rforceS", c_dp),
                ("zforceC", c_dp), ("zforceS", c_dp),
                ("Rtable", C.c_double), ("XMIN", C.c_double), ("XMAX", C.c_double), ("dX", C.c_double),
                ("YMIN", C.c_double), ("YMAX", C.c_double), ("dY", C.c_double)]


class CylParams(C.Structure):
    _fields_ = [("rcylmax", C.c_double), ("acyl", C.c_double), ("body", c_dp), ("orig", c_dp)]


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        build()
        L = C.CDLL(os.path.join(_HERE, "liboracle.so"))
        L.orc_factorial.restype = C.c_double
        L.orc_factorial.argtypes = [C.c_int, C.c_int]
        L.orc_r_to_xi.restype = C.c_double
        L.orc_r_to_xi.argtypes = [C.POINTER(SLGrid), C.c_double]
        L.orc_d_xi_to_r.restype = C.c_double
        L.orc_d_xi_to_r.argtypes = [C.POINTER(SLGrid), C.c_double]
        L.orc_sph_coefficients.restype = C.c_long
        L.orc_sph_coefficients_cov.restype = C.c_long
        L.orc_adjust_levels.restype = C.c_long
        for name in ("orc_slgrid_get_pot", "orc_slgrid_get_force", "orc_slgrid_get_dens"):
            getattr(L, name).argtypes = [C.POINTER(SLGrid), C.c_double, c_dp]
        L.orc_legendre_R.argtypes = [C.c_int, C.c_double, c_dp]
        L.orc_dlegendre_R.argtypes = [C.c_int, C.c_double, c_dp, c_dp]
        L.orc_sinecosine_R.argtypes = [C.c_int, C.c_double, c_dp, c_dp]
        L.orc_sph_coefficients.argtypes = [C.POINTER(SLGrid), C.POINTER(SphParams),
                                           C.POINTER(Particles), c_ip, C.c_long, C.c_double,
                                           C.c_int, c_dp]
        L.orc_sph_coefficients_cov.argtypes = [C.POINTER(SLGrid), C.POINTER(SphParams),
                                               C.POINTER(Particles), c_ip, C.c_long, C.c_double,
                                               C.c_int, c_dp, c_dp, c_dp, c_up, c_dp, c_dp]
        L.orc_sph_accel.argtypes = [C.POINTER(SLGrid), C.POINTER(SphParams),
                                    C.POINTER(Particles), c_ip, C.c_long, c_dp, C.c_int]
        L.orc_sph_multistep_update.argtypes = [C.POINTER(SLGrid), C.POINTER(SphParams),
                                               C.POINTER(Particles), C.c_int, C.c_double,
                                               C.c_int, C.c_int, c_dp]
        L.orc_multistep_interp.argtypes = [C.c_int, C.c_int, c_ip, c_ip, c_ip, C.c_long,
                                           c_dp, c_dp, c_dp]
        L.orc_incr_position.argtypes = [C.POINTER(Particles), c_ip, C.c_long, C.c_double, C.c_int]
        L.orc_incr_velocity.argtypes = [C.POINTER(Particles), c_ip, C.c_long, C.c_double, C.c_int]
        L.orc_adjust_levels.argtypes = [C.POINTER(Particles), c_ip, C.c_long,
                                        C.POINTER(LevelParams), c_ip, c_lp, c_lp]
        L.orc_multistep_tables.argtypes = [C.c_int, c_ip, c_ip, c_ip, c_ip]
        L.orc_fix_positions.argtypes = [C.POINTER(Particles), C.c_int, C.c_int, C.c_int, c_dp, c_dp]
        L.orc_cyl_init.argtypes = [C.POINTER(CylGrid)]
        L.orc_cyl_get_pot.argtypes = [C.POINTER(CylGrid), C.c_double, C.c_double, c_dp, c_dp]
        L.orc_cyl_coefficients.restype = C.c_long
        L.orc_cyl_coefficients.argtypes = [C.POINTER(CylGrid), C.POINTER(CylParams), C.POINTER(Particles),
                                           c_ip, C.c_long, C.c_double, C.c_int, c_dp, c_dp, c_dp]
        L.orc_cyl_coefficients_cov.restype = C.c_long
        L.orc_cyl_coefficients_cov.argtypes = [C.POINTER(CylGrid), C.POINTER(CylParams), C.POINTER(Particles),
                                               c_ip, C.c_long, C.c_double, C.c_int, c_dp, c_dp, c_dp,
                                               c_dp, c_up, c_dp, c_dp]
        L.orc_cyl_multistep_update.argtypes = [C.POINTER(CylGrid), C.POINTER(Particles), C.c_int, C.c_double,
                                               C.c_int, C.c_int, c_dp, c_dp]
        L.orc_cyl_accel.argtypes = [C.POINTER(CylGrid), C.POINTER(CylParams), C.POINTER(Particles),
                                    c_ip, C.c_long, c_dp, c_dp, C.c_double, C.c_int]
        _LIB = L
    return _LIB


def ref_lib():
    """The reference's own src/Basis.cc compiled in place (None if not built)."""
    global _REF
    if _REF is None:
        build()
        p = os.path.join(_HERE, "_ref", "libexpref_basis.so")
        if not os.path.exists(p):
            return None
        R = C.CDLL(p)
        R.ref_legendre_R.argtypes = [C.c_int, C.c_double, c_dp]
        R.ref_dlegendre_R.argtypes = [C.c_int, C.c_double, c_dp, c_dp]
        R.ref_sinecosine_R.argtypes = [C.c_int, C.c_double, c_dp, c_dp]
        _REF = R
    return _REF


def _dp(a):
    return a.ctypes.data_as(c_dp) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Grid:
    """Holds the numpy arrays alive next to the C struct."""

    def __init__(self, tab):
        self.ev, self.ef, self.p0, self.d0 = _f64(tab.ev), _f64(tab.ef), _f64(tab.p0), _f64(tab.d0)
        self.c = SLGrid(tab.lmax, tab.nmax, tab.numr, tab.cmap, tab.rmin, tab.rmax, tab.rmap,
                        _dp(self.ev), _dp(self.ef), _dp(self.p0), _dp(self.d0), 0.0, 0.0, 0.0)
        lib().orc_slgrid_init(C.byref(self.c))
        self.lmax, self.nmax, self.numr = tab.lmax, tab.nmax, tab.numr

    def get_pot(self, r):
        m = np.zeros((self.lmax + 1, self.nmax))
        lib().orc_slgrid_get_pot(C.byref(self.c), float(r), _dp(m))
        return m

    def get_force(self, r):
        m = np.zeros((self.lmax + 1, self.nmax))
        lib().orc_slgrid_get_force(C.byref(self.c), float(r), _dp(m))
        return m


class ParticleSet:
    """numpy-backed particle arrays in the oracle's layout."""

    def __init__(self, pos, vel=None, mass=None, center=(0.0, 0.0, 0.0), rtrunc=1e20,
                 level=None, scale=None):
        n = len(pos)
        self.n = n
        self.pos = _f64(pos).copy()
        self.vel = _f64(vel).copy() if vel is not None else np.zeros((n, 3))
        self.mass = _f64(mass).copy() if mass is not None else np.full(n, 1.0 / max(n, 1))
        self.acc = np.zeros((n, 3))
        self.pot = np.zeros(n)
        self.potext = np.zeros(n)
        self.scale = _f64(scale).copy() if scale is not None else None
        self.level = (np.ascontiguousarray(level, dtype=np.uint32).copy()
                      if level is not None else np.zeros(n, dtype=np.uint32))
        self.dtreq = np.zeros(n, dtype=np.float32)
        self.center = tuple(float(c) for c in center)
        self.rtrunc = float(rtrunc)

    def c(self) -> Particles:
        return Particles(self.n, _dp(self.pos), _dp(self.vel), _dp(self.acc), _dp(self.pot),
                         _dp(self.potext), _dp(self.mass), _dp(self.scale),
                         self.level.ctypes.data_as(c_up), self.dtreq.ctypes.data_as(c_fp),
                         (C.c_double * 3)(*self.center), self.rtrunc)

    def levlist(self, lev):
        return np.nonzero(self.level == lev)[0].astype(np.int32)


def sph_params(scale=1.0, rmin=0.0, rmax=1e300, NO_L0=False, NO_L1=False, EVEN_L=False,
               EVEN_M=False, M0_only=False) -> SphParams:
    return SphParams(scale, rmin, rmax, int(NO_L0), int(NO_L1), int(EVEN_L), int(EVEN_M),
                     int(M0_only))


def _list(S, lst):
    if lst is None:
        lst = np.arange(S.n, dtype=np.int32)
    return np.ascontiguousarray(lst, dtype=np.int32)


def sph_coefficients(grid: Grid, P: SphParams, S: ParticleSet, lst=None, adb=1.0, nthreads=1):
    lst = _list(S, lst)
    coef = np.zeros(((grid.lmax + 1) ** 2, grid.nmax))
    sc = S.c()
    used = lib().orc_sph_coefficients(C.byref(grid.c), C.byref(P), C.byref(sc),
                                      lst.ctypes.data_as(c_ip), len(lst), adb, nthreads, _dp(coef))
    return coef, used


def sph_coefficients_cov(grid: Grid, P: SphParams, S: ParticleSet, sampT, lst=None, adb=1.0):
    lst = _list(S, lst)
    L, N = grid.lmax, grid.nmax
    ldim = (L + 1) * (L + 2) // 2
    coef = np.zeros(((L + 1) ** 2, N))
    muse = np.zeros(1)
    massT = np.zeros(sampT)
    countT = np.zeros(sampT, dtype=np.uint32)
    coefT = np.zeros((sampT, ldim, N, 2))
    coefM = np.zeros((sampT, ldim, N, N))
    sc = S.c()
    used = lib().orc_sph_coefficients_cov(C.byref(grid.c), C.byref(P), C.byref(sc),
                                          lst.ctypes.data_as(c_ip), len(lst), adb, sampT,
                                          _dp(coef), _dp(muse), _dp(massT),
                                          countT.ctypes.data_as(c_up), _dp(coefT), _dp(coefM))
    return dict(coef=coef, used=used, muse=muse[0], massT=massT, countT=countT,
                coefT=coefT[..., 0] + 1j * coefT[..., 1], coefM=coefM)


def sph_accel(grid: Grid, P: SphParams, S: ParticleSet, coef, lst=None, nthreads=1):
    """Adds into S.acc / S.pot."""
    lst = _list(S, lst)
    coef = _f64(coef)
    sc = S.c()
    lib().orc_sph_accel(C.byref(grid.c), C.byref(P), C.byref(sc), lst.ctypes.data_as(c_ip),
                        len(lst), _dp(coef), nthreads)


def sph_multistep_update(grid, P, S, indx, frm, to, differ, adb=1.0):
    sc = S.c()
    lib().orc_sph_multistep_update(C.byref(grid.c), C.byref(P), C.byref(sc), int(indx), adb,
                                   int(frm), int(to), _dp(differ))


def multistep_tables(multistep):
    Mstep = 1 << multistep
    mintvl = np.zeros(multistep + 1, dtype=np.int32)
    mfirst = np.zeros(Mstep + 1, dtype=np.int32)
    dL = np.zeros((multistep + 1, Mstep), dtype=np.int32)
    dN = np.zeros((multistep + 1, Mstep), dtype=np.int32)
    lib().orc_multistep_tables(multistep, mintvl.ctypes.data_as(c_ip), mfirst.ctypes.data_as(c_ip),
                               dL.ctypes.data_as(c_ip), dN.ctypes.data_as(c_ip))
    return mintvl, mfirst, dL, dN


def multistep_interp(multistep, mdrft, coefL, coefN):
    _, mfirst, dL, dN = multistep_tables(multistep)
    coefL, coefN = _f64(coefL), _f64(coefN)
    ncoef = coefL[0].size
    out = np.zeros(coefL[0].shape)
    lib().orc_multistep_interp(multistep, mdrft, mfirst.ctypes.data_as(c_ip),
                               dL.ctypes.data_as(c_ip), dN.ctypes.data_as(c_ip), ncoef,
                               _dp(coefL), _dp(coefN), _dp(out))
    return out


def incr_position(S: ParticleSet, dt, lst=None, dim=3):
    sc = S.c()
    l = None if lst is None else _list(S, lst)
    lib().orc_incr_position(C.byref(sc), l.ctypes.data_as(c_ip) if l is not None else None,
                            S.n if l is None else len(l), dt, dim)


def incr_velocity(S: ParticleSet, dt, lst=None, dim=3):
    sc = S.c()
    l = None if lst is None else _list(S, lst)
    lib().orc_incr_velocity(C.byref(sc), l.ctypes.data_as(c_ip) if l is not None else None,
                            S.n if l is None else len(l), dt, dim)


def fix_positions(S: ParticleSet, mlevel, multistep, lev_sums=None, dim=3):
    """Component::fix_positions_cpu (src/Component.cc:3354-3590): returns (lev_sums, out) with
    lev_sums[M] = [com_mas, com_lev[3], cov_lev[3], coa_lev[3]] (updated for M >= mlevel) and
    out = [mtot, com[3], cov[3], coa[3]]."""
    if lev_sums is None:
        lev_sums = np.zeros((multistep + 1, 10))
    lev_sums = np.ascontiguousarray(lev_sums, dtype=np.float64)
    out = np.zeros(10)
    sc = S.c()
    lib().orc_fix_positions(C.byref(sc), mlevel, multistep, dim, _dp(lev_sums), _dp(out))
    return lev_sums, out


def level_params(dtime, multistep, mfirst_mdrft=0, dynfracS=1.0, dynfracD=1.0e32, dynfracV=0.01,
                 dynfracA=0.03, dynfracP=0.05, shiftlevl=0, noswitch=False,
                 reset_dtreq=False, apply=True, dim=3) -> LevelParams:
    """Defaults follow the reference's src/global.cc:70-74."""
    return LevelParams(dtime, dynfracS, dynfracD, dynfracV, dynfracA, dynfracP, multistep,
                       shiftlevl, mfirst_mdrft, int(noswitch), int(reset_dtreq), int(apply), dim)


def adjust_levels(S: ParticleSet, P: LevelParams, lst=None):
    """Updates S.level / S.dtreq in place; returns (nswitch, which, offlo, offhi)."""
    sc = S.c()
    l = None if lst is None else _list(S, lst)
    which = np.zeros(S.n, dtype=np.int32)
    offlo, offhi = C.c_long(0), C.c_long(0)
    nsw = lib().orc_adjust_levels(C.byref(sc), l.ctypes.data_as(c_ip) if l is not None else None,
                                  S.n if l is None else len(l), C.byref(P),
                                  which.ctypes.data_as(c_ip), C.byref(offlo), C.byref(offhi))
    return nsw, which, offlo.value, offhi.value


# ---- cylinder --------------------------------------------------------------------
class CylGridPy:
    def __init__(self, tab, EVEN_M=False, MMIN=0, MLIM=2 ** 30, NMIN=0, NLIM=2 ** 30):
        self.arr = [_f64(getattr(tab, k)) for k in ("potC", "potS", "rforceC", "rforceS", "zforceC", "zforceS")]
        self.c = CylGrid(tab.mmax, tab.norder, tab.numx, tab.numy, tab.cmapr, tab.cmapz, tab.ascale,
                         tab.hscale, tab.rmin, tab.rmax, int(EVEN_M), MMIN, MLIM, NMIN, NLIM,
                         *[_dp(a) for a in self.arr], 0, 0, 0, 0, 0, 0, 0)
        lib().orc_cyl_init(C.byref(self.c))
        self.mmax, self.norder = tab.mmax, tab.norder


def cyl_params(rcylmax=20.0, acyl=0.01, body=None, orig=None):
    keep = [None if body is None else _f64(body), None if orig is None else _f64(orig)]
    p = CylParams(rcylmax, acyl, _dp(keep[0]), _dp(keep[1]))
    p._keep = keep
    return p


def cyl_coefficients(g: CylGridPy, P, S: ParticleSet, lst=None, adb=1.0, nthreads=1):
    lst = _list(S, lst)
    shape = (g.mmax + 1, g.norder)
    cs, sn = np.zeros(shape), np.zeros(shape)
    mass = np.zeros(1)
    sc = S.c()
    used = lib().orc_cyl_coefficients(C.byref(g.c), C.byref(P), C.byref(sc), lst.ctypes.data_as(c_ip),
                                      len(lst), adb, nthreads, _dp(cs), _dp(sn), _dp(mass))
    return cs, sn, used, mass[0]


def cyl_coefficients_cov(g: CylGridPy, P, S: ParticleSet, sampT, lst=None, adb=1.0):
    lst = _list(S, lst)
    M1, N = g.mmax + 1, g.norder
    cs, sn = np.zeros((M1, N)), np.zeros((M1, N))
    mass = np.zeros(1)
    massT, numbT = np.zeros(sampT), np.zeros(sampT, dtype=np.uint32)
    VC, MV = np.zeros((sampT, M1, N, 2)), np.zeros((sampT, M1, N, N, 2))
    sc = S.c()
    used = lib().orc_cyl_coefficients_cov(C.byref(g.c), C.byref(P), C.byref(sc), lst.ctypes.data_as(c_ip),
                                          len(lst), adb, sampT, _dp(cs), _dp(sn), _dp(mass), _dp(massT),
                                          numbT.ctypes.data_as(c_up), _dp(VC), _dp(MV))
    return dict(cos=cs, sin=sn, used=used, mass=mass[0], massT=massT, numbT=numbT,
                VC=VC[..., 0] + 1j * VC[..., 1], MV=MV[..., 0] + 1j * MV[..., 1])


def cyl_multistep_update(g: CylGridPy, S: ParticleSet, indx, frm, to, differC, differS, adb=1.0):
    sc = S.c()
    lib().orc_cyl_multistep_update(C.byref(g.c), C.byref(sc), int(indx), adb, int(frm), int(to),
                                   _dp(differC), _dp(differS))


def cyl_accel(g: CylGridPy, P, S: ParticleSet, cos, sin, cylmass, lst=None, nthreads=1):
    lst = _list(S, lst)
    cos, sin = _f64(cos), _f64(sin)
    sc = S.c()
    lib().orc_cyl_accel(C.byref(g.c), C.byref(P), C.byref(sc), lst.ctypes.data_as(c_ip), len(lst),
                        _dp(cos), _dp(sin), float(cylmass), nthreads)
