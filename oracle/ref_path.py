"""ctypes binding of oracle/_ref/libexpref_path.so: the reference's OWN hot-path member
functions (SLGridSph::get_pot/get_force, SphericalBasis::determine_coefficients_thread /
determine_acceleration_and_potential_thread, EmpCylSL::get_pot/accumulate/accumulated_eval,
incr_position_thread / incr_velocity_thread, adjust_multistep_level_thread, Basis recursions)
compiled from /root/reference where they lie (oracle/ref_extract.py + ref_path_wrap.cc).

TEST INFRASTRUCTURE ONLY: loaded by tests/ to pin the restatement in oracle/*_oracle.cc and
to check the CUDA path against the reference's arithmetic directly.  The .so is built in this
container (the GPU box has no /root/reference) and travels with the repo snapshot.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(_HERE, "_ref", "libexpref_path.so")
_LIB = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_up = C.POINTER(C.c_uint)
c_fp = C.POINTER(C.c_float)
c_ulp = C.POINTER(C.c_ulong)


class LevelParams(C.Structure):
    _fields_ = [("dtime", C.c_double), ("dynfracS", C.c_double), ("dynfracD", C.c_double),
                ("dynfracV", C.c_double), ("dynfracA", C.c_double), ("dynfracP", C.c_double),
                ("multistep", C.c_int), ("shiftlevl", C.c_int), ("mfirst_mdrft", C.c_int),
                ("minlev", C.c_int), ("noswitch", C.c_int), ("first_call", C.c_int),
                ("dim", C.c_int), ("nthreads", C.c_int)]


def available() -> bool:
    if not os.path.exists(PATH) and os.path.isdir("/root/reference"):
        from . import oracle
        oracle.build(force=True)
    return os.path.exists(PATH)


def lib():
    global _LIB
    if _LIB is None:
        if not available():
            return None
        L = C.CDLL(PATH)
        L.ref_sph_create.restype = C.c_void_p
        L.ref_sph_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                     C.c_double, c_dp, c_dp, c_dp, c_dp]
        L.ref_sph_destroy.argtypes = [C.c_void_p]
        L.ref_slgrid_consts.argtypes = [C.c_void_p, c_dp, c_dp, c_dp]
        for nm in ("ref_r_to_xi", "ref_xi_to_r", "ref_d_xi_to_r"):
            getattr(L, nm).restype = C.c_double
            getattr(L, nm).argtypes = [C.c_void_p, C.c_double]
        for nm in ("ref_slgrid_get_pot", "ref_slgrid_get_force", "ref_slgrid_get_dens"):
            getattr(L, nm).argtypes = [C.c_void_p, C.c_double, c_dp]
        L.ref_factorial.restype = C.c_double
        L.ref_factorial.argtypes = [C.c_int, C.c_int, C.c_int]
        L.ref_sph_coefficients.restype = C.c_long
        L.ref_sph_coefficients.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, c_ip,
                                           C.c_long, c_dp, c_dp, c_dp, C.c_double, c_ip, C.c_long,
                                           C.c_double, C.c_int, c_dp, C.c_int, c_dp, c_dp, c_up,
                                           c_dp, c_dp, c_dp]
        L.ref_sph_accel.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, c_ip, C.c_long,
                                    c_dp, c_dp, C.c_double, c_ip, C.c_long, c_dp, C.c_int, c_dp,
                                    c_dp, c_dp]
        L.ref_sph_multistep_update.argtypes = ([C.c_void_p, C.c_double, C.c_double, C.c_double, c_ip, C.c_long, c_dp,
                                                c_dp, c_dp, C.c_double, C.c_double, C.c_int, C.c_long, c_ip, c_ip,
                                                c_ip, c_dp])
        L.ref_sph_compute_multistep_coefficients.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, c_ip, c_ip, c_ip,
                                                             c_dp, c_dp, c_dp]
        L.ref_sph_load.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, c_ip, C.c_long, c_dp, c_dp,
                                   c_dp, C.c_int]
        L.ref_sph_step.restype = C.c_long
        L.ref_sph_step.argtypes = [C.c_void_p, C.c_double, C.c_int]
        L.ref_sph_read.argtypes = [C.c_void_p, c_dp, c_dp, c_dp, c_dp, c_dp]
        L.ref_cyl_create.restype = C.c_void_p
        L.ref_pyexp_create.restype = C.c_void_p
        L.ref_pyexp_create.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int]
        L.ref_pyexp_destroy.argtypes = [C.c_void_p]
        L.ref_pyexp_accumulate.restype = C.c_int
        L.ref_pyexp_accumulate.argtypes = [C.c_void_p, C.c_size_t, c_dp, c_dp, c_dp]
        L.ref_pyexp_set_coefs.argtypes = [C.c_void_p, c_dp]
        L.ref_pyexp_eval.argtypes = [C.c_void_p, C.c_int, C.c_size_t, c_dp, c_dp]
        L.ref_pyexp_accel.argtypes = [C.c_void_p, C.c_size_t, c_dp, c_dp]
        L.ref_cyl_create.argtypes = [C.c_int] * 6 + [C.c_double] * 7 + [c_ip] + [c_dp] * 6
        L.ref_cyl_destroy.argtypes = [C.c_void_p]
        L.ref_cylinder_multistep_update.argtypes = [C.c_void_p, C.c_long, c_dp, c_dp, c_dp, C.c_double, C.c_double,
                                                    C.c_int, C.c_long, c_ip, c_ip, c_ip, c_dp, c_dp]
        L.ref_cylexp_compute_multistep_coefficients.argtypes = ([C.c_void_p, C.c_int, C.c_int, C.c_int, c_ip, c_ip, c_ip]
                                                                + [c_dp] * 6)
        L.ref_fix_positions.argtypes = [C.c_long, c_dp, c_dp, c_dp, c_dp, c_up, C.c_int, C.c_int, C.c_int, c_dp,
                                        C.c_double, c_dp]
        L.ref_cylinder_coef_thread.argtypes = ([C.c_void_p, C.c_long, c_dp, c_dp, c_up, C.c_int, C.c_int, c_dp]
                                               + [C.c_double] * 4 + [c_dp] * 4 + [c_ip, c_dp])
        L.ref_cylinder_accel_thread.argtypes = ([C.c_void_p, C.c_long, c_dp, c_up, C.c_int, C.c_int, c_dp, C.c_double,
                                                 c_dp, c_dp, c_dp, C.c_double, c_dp, c_dp, c_dp, c_dp])
        L.ref_cyl_r_to_xi.restype = C.c_double
        L.ref_cyl_r_to_xi.argtypes = [C.c_void_p, C.c_double]
        L.ref_cyl_z_to_y.restype = C.c_double
        L.ref_cyl_z_to_y.argtypes = [C.c_void_p, C.c_double]
        L.ref_cyl_get_pot.argtypes = [C.c_void_p, C.c_double, C.c_double, c_dp, c_dp]
        L.ref_cyl_accumulate.restype = C.c_long
        L.ref_cyl_accumulate.argtypes = [C.c_void_p, C.c_long, c_dp, c_dp, c_dp, c_dp, c_ulp, c_dp,
                                         c_dp, c_dp, C.c_int, c_up, c_dp, c_dp, c_dp]
        L.ref_cyl_accumulate_pca.argtypes = [C.c_void_p, C.c_long, c_dp, c_dp, c_dp, c_dp, c_ulp, C.c_int,
                                             C.c_int, c_dp, c_dp]
        L.ref_cyl_accumulated_eval.argtypes = [C.c_void_p, c_dp, c_dp, C.c_double, C.c_double,
                                               C.c_double, c_dp, c_dp, c_dp, c_dp, c_dp]
        L.ref_incr_position.argtypes = [C.c_long, c_dp, c_dp, c_up, C.c_int, C.c_double, C.c_int,
                                        C.c_int, C.c_int]
        L.ref_incr_velocity.argtypes = [C.c_long, c_dp, c_dp, c_dp, c_up, C.c_int, C.c_double,
                                        C.c_int, C.c_int, C.c_int]
        L.ref_adjust_levels.restype = C.c_long
        L.ref_adjust_levels.argtypes = [C.c_long, c_dp, c_dp, c_dp, c_dp, c_fp, c_up, c_fp,
                                        C.POINTER(LevelParams), C.POINTER(C.c_long),
                                        C.POINTER(C.c_long)]
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(c_dp) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _flags(*f):
    return (C.c_int * len(f))(*[int(x) for x in f])


class RefSphere:
    """SLGridSph + Sphere of the reference over a table object (lmax, nmax, numr, cmap, rmin,
    rmax, rmap, ev, ef, p0, d0)."""

    def __init__(self, tab, scale=1.0, rmin=None, rmax=None, NO_L0=False, NO_L1=False,
                 EVEN_L=False, EVEN_M=False, M0_only=False):
        self.tab = tab
        self.lmax, self.nmax, self.numr = tab.lmax, tab.nmax, tab.numr
        self._keep = [_f64(tab.ev), _f64(tab.ef), _f64(tab.p0), _f64(tab.d0)]
        self.h = lib().ref_sph_create(tab.lmax, tab.nmax, tab.numr, tab.cmap, tab.rmin, tab.rmax,
                                      tab.rmap, *[_dp(a) for a in self._keep])
        self.scale = float(scale)
        self.rmin = float(tab.rmin if rmin is None else rmin)
        self.rmax = float(tab.rmax if rmax is None else rmax)
        self.flags = _flags(NO_L0, NO_L1, EVEN_L, EVEN_M, M0_only)

    def __del__(self):
        try:
            lib().ref_sph_destroy(self.h)
        except Exception:
            pass

    def multistep_update(self, pos, mass, idx, frm, to, nlev, center=(0.0, 0.0, 0.0), rtrunc=1.0e20, adb=1.0):
        """SphericalBasis::multistep_update (src/SphericalBasis.cc:1156-1228) for the switching particles
        idx (old level frm, new level to) -> differ [nlev][(lmax+1)^2][nmax]"""
        pos, mass, ctr = _f64(pos), _f64(mass), _f64(center)
        idx, frm, to = (np.ascontiguousarray(a, dtype=np.int32) for a in (idx, frm, to))
        out = np.zeros((nlev, (self.lmax + 1) ** 2, self.nmax))
        lib().ref_sph_multistep_update(self.h, self.scale, self.rmin, self.rmax, self.flags, len(pos), _dp(pos),
                                       _dp(mass), _dp(ctr), float(rtrunc), float(adb), nlev, len(idx),
                                       idx.ctypes.data_as(c_ip), frm.ctypes.data_as(c_ip),
                                       to.ctypes.data_as(c_ip), _dp(out))
        return out

    def compute_multistep_coefficients(self, mdrft, mfirst, dL, dN, coefL, coefN):
        """SphericalBasis::compute_multistep_coefficients (src/SphericalBasis.cc:1231-1474) with the tables
        of initialize_multistep; coefL, coefN [nlev][(lmax+1)^2][nmax] -> the table in use"""
        coefL, coefN = _f64(coefL), _f64(coefN)
        nlev = coefL.shape[0]
        mfirst, dL, dN = (np.ascontiguousarray(a, dtype=np.int32) for a in (mfirst, dL, dN))
        out = np.zeros(coefL.shape[1:])
        lib().ref_sph_compute_multistep_coefficients(self.h, nlev, dL.shape[1], int(mdrft),
                                                     mfirst.ctypes.data_as(c_ip), dL.ctypes.data_as(c_ip),
                                                     dN.ctypes.data_as(c_ip), _dp(coefL), _dp(coefN), _dp(out))
        return out

    def consts(self):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        lib().ref_slgrid_consts(self.h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def r_to_xi(self, r): return lib().ref_r_to_xi(self.h, float(r))
    def xi_to_r(self, x): return lib().ref_xi_to_r(self.h, float(x))
    def d_xi_to_r(self, x): return lib().ref_d_xi_to_r(self.h, float(x))

    def _mat(self, fn, r):
        m = np.zeros((self.lmax + 1, self.nmax))
        fn(self.h, float(r), _dp(m))
        return m

    def get_pot(self, r): return self._mat(lib().ref_slgrid_get_pot, r)
    def get_force(self, r): return self._mat(lib().ref_slgrid_get_force, r)
    def get_dens(self, r): return self._mat(lib().ref_slgrid_get_dens, r)

    def coefficients(self, pos, mass, lst=None, adb=1.0, nthreads=1, center=(0, 0, 0), rtrunc=1e20,
                     sampT=0, pcaeof=False):
        pos, mass = _f64(pos), _f64(mass)
        n = len(pos)
        lst = np.arange(n, dtype=np.int32) if lst is None else np.ascontiguousarray(lst, dtype=np.int32)
        L, N = self.lmax, self.nmax
        ldim = (L + 1) * (L + 2) // 2
        coef = np.zeros(((L + 1) ** 2, N))
        ctr = _f64(center)
        out = dict()
        if sampT > 0:
            muse, massT = np.zeros(1), np.zeros(sampT)
            countT = np.zeros(sampT, dtype=np.uint32)
            coefT, coefM = np.zeros((sampT, ldim, N, 2)), np.zeros((sampT, ldim, N, N, 2))
            tvar = np.zeros((ldim, N, N)) if pcaeof else None
            used = lib().ref_sph_coefficients(self.h, self.scale, self.rmin, self.rmax, self.flags, n,
                                              _dp(pos), _dp(mass), _dp(ctr), float(rtrunc),
                                              lst.ctypes.data_as(c_ip), len(lst), float(adb), 1,
                                              _dp(coef), sampT, _dp(muse), _dp(massT),
                                              countT.ctypes.data_as(c_up), _dp(coefT), _dp(coefM), _dp(tvar))
            out = dict(muse=muse[0], massT=massT, countT=countT, coefT=coefT[..., 0] + 1j * coefT[..., 1],
                       coefM=coefM[..., 0] + 1j * coefM[..., 1], tvar=tvar)
        else:
            used = lib().ref_sph_coefficients(self.h, self.scale, self.rmin, self.rmax, self.flags, n,
                                              _dp(pos), _dp(mass), _dp(ctr), float(rtrunc),
                                              lst.ctypes.data_as(c_ip), len(lst), float(adb), nthreads,
                                              _dp(coef), 0, None, None, None, None, None, None)
        out.update(coef=coef, used=used)
        return out

    def accel(self, pos, coef, lst=None, nthreads=1, center=(0, 0, 0), rtrunc=1e20, own_center=None,
              acc=None, pot=None):
        """Returns (acc, pot) = inputs (zeros when None) + this basis' field."""
        pos, coef = _f64(pos), _f64(coef)
        n = len(pos)
        lst = np.arange(n, dtype=np.int32) if lst is None else np.ascontiguousarray(lst, dtype=np.int32)
        acc = np.zeros((n, 3)) if acc is None else _f64(acc).copy()
        pot = np.zeros(n) if pot is None else _f64(pot).copy()
        ctr = _f64(center)
        own = None if own_center is None else _f64(own_center)
        lib().ref_sph_accel(self.h, self.scale, self.rmin, self.rmax, self.flags, n, _dp(pos), _dp(ctr),
                            float(rtrunc), lst.ctypes.data_as(c_ip), len(lst), _dp(coef), nthreads,
                            _dp(own), _dp(acc), _dp(pot))
        return acc, pot


class RefSphereStepper:
    """The reference's own thread functions (incr_velocity/incr_position threads, coefficient and force
    threads of SphericalBasis) driven in do_step's single-level order on `nthreads` host threads: the
    CPU baseline bench.py times (kind "reference")."""

    def __init__(self, tab, pos, vel, mass, nthreads=1, scale=1.0):
        self.ref = RefSphere(tab, scale=scale)
        self.n = len(pos)
        pos, vel, mass = _f64(pos), _f64(vel), _f64(mass)
        lib().ref_sph_load(self.ref.h, self.ref.scale, self.ref.rmin, self.ref.rmax, self.ref.flags, self.n,
                           _dp(pos), _dp(vel), _dp(mass), int(nthreads))

    def force(self):
        """coefficients + force on the current positions (no kicks), returns the number used"""
        return lib().ref_sph_step(self.ref.h, 0.0, 0)

    def step(self, dt):
        return lib().ref_sph_step(self.ref.h, float(dt), 1)

    def read(self):
        out = dict(pos=np.zeros((self.n, 3)), vel=np.zeros((self.n, 3)), acc=np.zeros((self.n, 3)), pot=np.zeros(self.n),
                   coef=np.zeros(((self.ref.lmax + 1) ** 2, self.ref.nmax)))
        lib().ref_sph_read(self.ref.h, _dp(out["pos"]), _dp(out["vel"]), _dp(out["acc"]), _dp(out["pot"]), _dp(out["coef"]))
        return out


class RefCylinder:
    """EmpCylSL of the reference over a cylinder table object; grid constants (Rtable, XMIN, dX,
    YMIN, dY) are passed in as EmpCylSL::setup_table derives them (oracle CylGrid fields)."""

    def __init__(self, tab, grid_consts, EVEN_M=False, MMIN=0, MLIM=2 ** 30, NMIN=0, NLIM=2 ** 30):
        self.mmax, self.norder = tab.mmax, tab.norder
        self._keep = [_f64(getattr(tab, k)) for k in ("potC", "potS", "rforceC", "rforceS", "zforceC", "zforceS")]
        Rtable, XMIN, dX, YMIN, dY = grid_consts
        self.h = lib().ref_cyl_create(tab.mmax, tab.norder, tab.numx, tab.numy, tab.cmapr, tab.cmapz,
                                      tab.ascale, tab.hscale, Rtable, XMIN, dX, YMIN, dY,
                                      _flags(EVEN_M, MMIN, MLIM, NMIN, NLIM), *[_dp(a) for a in self._keep])

    def __del__(self):
        try:
            lib().ref_cyl_destroy(self.h)
        except Exception:
            pass

    def r_to_xi(self, r): return lib().ref_cyl_r_to_xi(self.h, float(r))
    def z_to_y(self, z): return lib().ref_cyl_z_to_y(self.h, float(z))

    def coef_thread(self, pos, mass, level=None, nlev=1, mlevel=0, center=None, rtrunc=1.0e20, rcylmax=20.0,
                    acyl=0.01, adb=1.0, body=None, orig=None):
        """Cylinder::determine_coefficients_thread (src/Cylinder.cc:748-896) for levlist[mlevel], one thread
        -> cos, sin [(mmax+1)][norder], used, cylmass0"""
        pos, mass = _f64(pos), _f64(mass)
        lev = None if level is None else np.ascontiguousarray(level, dtype=np.uint32)
        ctr = None if center is None else _f64(center)
        b, o = (None, None) if body is None else (_f64(body), _f64(orig))
        cs, sn = np.zeros((self.mmax + 1, self.norder)), np.zeros((self.mmax + 1, self.norder))
        used, cm = C.c_int(0), C.c_double(0.0)
        lib().ref_cylinder_coef_thread(self.h, len(pos), _dp(pos), _dp(mass),
                                       None if lev is None else lev.ctypes.data_as(c_up), nlev, mlevel,
                                       None if ctr is None else _dp(ctr), rtrunc, rcylmax, acyl, adb,
                                       None if b is None else _dp(b), None if o is None else _dp(o),
                                       _dp(cs), _dp(sn), C.byref(used), C.byref(cm))
        return cs, sn, used.value, cm.value

    def multistep_update(self, pos, mass, idx, frm, to, nlev, center=(0.0, 0.0, 0.0), rtrunc=1.0e20, adb=1.0):
        """Cylinder::multistep_update -> CylEXP::multistep_update (src/Cylinder.cc:1752-1773,
        src/CylEXP.cc:159-188) for the switching particles -> differC, differS [nlev][mmax+1][norder]"""
        pos, mass, ctr = _f64(pos), _f64(mass), _f64(center)
        idx, frm, to = (np.ascontiguousarray(a, dtype=np.int32) for a in (idx, frm, to))
        dC = np.zeros((nlev, self.mmax + 1, self.norder))
        dS = np.zeros((nlev, self.mmax + 1, self.norder))
        lib().ref_cylinder_multistep_update(self.h, len(pos), _dp(pos), _dp(mass), _dp(ctr), float(rtrunc),
                                            float(adb), nlev, len(idx), idx.ctypes.data_as(c_ip),
                                            frm.ctypes.data_as(c_ip), to.ctypes.data_as(c_ip), _dp(dC), _dp(dS))
        return dC, dS

    def compute_multistep_coefficients(self, mdrft, mfirst, dL, dN, cosL, cosN, sinL, sinN):
        """CylEXP::compute_multistep_coefficients (src/CylEXP.cc:192-282); tables [nlev][mmax+1][norder]"""
        arrs = [_f64(a) for a in (cosL, cosN, sinL, sinN)]
        nlev = arrs[0].shape[0]
        mfirst, dL, dN = (np.ascontiguousarray(a, dtype=np.int32) for a in (mfirst, dL, dN))
        ac, as_ = np.zeros(arrs[0].shape[1:]), np.zeros(arrs[0].shape[1:])
        lib().ref_cylexp_compute_multistep_coefficients(self.h, nlev, dL.shape[1], int(mdrft),
                                                        mfirst.ctypes.data_as(c_ip), dL.ctypes.data_as(c_ip),
                                                        dN.ctypes.data_as(c_ip), *[_dp(a) for a in arrs],
                                                        _dp(ac), _dp(as_))
        return ac, as_

    def accel_thread(self, pos, cos, sin, cylmass, level=None, nlev=1, mlevel=0, center=None, rtrunc=1.0e20,
                     owner_center=None, body=None, orig=None, acc=None, pot=None):
        """Cylinder::determine_acceleration_and_potential_thread (src/Cylinder.cc:1266-1446) for the levels
        >= mlevel, one thread; owner_center: the use_external case.  Returns acc [n][3], pot [n] (added to
        the given ones)."""
        pos = _f64(pos)
        n = len(pos)
        lev = None if level is None else np.ascontiguousarray(level, dtype=np.uint32)
        ctr = None if center is None else _f64(center)
        own = None if owner_center is None else _f64(owner_center)
        b, o = (None, None) if body is None else (_f64(body), _f64(orig))
        acc = np.zeros((n, 3)) if acc is None else _f64(acc).copy()
        pot = np.zeros(n) if pot is None else _f64(pot).copy()
        cs, sn = _f64(cos), _f64(sin)
        lib().ref_cylinder_accel_thread(self.h, n, _dp(pos), None if lev is None else lev.ctypes.data_as(c_up),
                                        nlev, mlevel, None if ctr is None else _dp(ctr), rtrunc,
                                        None if own is None else _dp(own), _dp(cs), _dp(sn), float(cylmass),
                                        None if b is None else _dp(b), None if o is None else _dp(o),
                                        _dp(acc), _dp(pot))
        return acc, pot

    def get_pot(self, r, z):
        Vc, Vs = np.zeros((self.mmax + 1, self.norder)), np.zeros((self.mmax + 1, self.norder))
        lib().ref_cyl_get_pot(self.h, float(r), float(z), _dp(Vc), _dp(Vs))
        return Vc, Vs

    def accumulate(self, pos, mass, seq=None, sampT=0):
        """Cylinder coordinates as src/Cylinder.cc:855-869 forms them (r = sqrt(x^2+y^2),
        phi = atan2(y, x) with libm), then EmpCylSL::accumulate per particle."""
        pos, mass = _f64(pos), _f64(mass)
        n = len(pos)
        r = np.sqrt(pos[:, 0] * pos[:, 0] + pos[:, 1] * pos[:, 1])
        z = pos[:, 2].copy()
        phi = np.array([math.atan2(y, x) for x, y in zip(pos[:, 0], pos[:, 1])])
        seq = np.arange(n, dtype=np.uint64) if seq is None else np.ascontiguousarray(seq, dtype=np.uint64)
        M1, N = self.mmax + 1, self.norder
        cs, sn = np.zeros((M1, N)), np.zeros((M1, N))
        cm = np.zeros(1)
        if sampT > 0:
            numbT, massT = np.zeros(sampT, dtype=np.uint32), np.zeros(sampT)
            VC, MV = np.zeros((sampT, M1, N, 2)), np.zeros((sampT, M1, N, N, 2))
            used = lib().ref_cyl_accumulate(self.h, n, _dp(r), _dp(z), _dp(phi), _dp(mass),
                                            seq.ctypes.data_as(c_ulp), _dp(cs), _dp(sn), _dp(cm), sampT,
                                            numbT.ctypes.data_as(c_up), _dp(massT), _dp(VC), _dp(MV))
            return dict(cos=cs, sin=sn, used=used, cylmass1=cm[0], numbT=numbT, massT=massT,
                        VC=VC[..., 0] + 1j * VC[..., 1], MV=MV[..., 0] + 1j * MV[..., 1])
        used = lib().ref_cyl_accumulate(self.h, n, _dp(r), _dp(z), _dp(phi), _dp(mass),
                                        seq.ctypes.data_as(c_ulp), _dp(cs), _dp(sn), _dp(cm), 0,
                                        None, None, None, None)
        return dict(cos=cs, sin=sn, used=used, cylmass1=cm[0])

    def accumulate_pca(self, pos, mass, which, sampT=1, seq=None):
        """the PCAVAR ('pcavar') / PCAEOF ('pcaeof') branches of EmpCylSL::accumulate: covV, covM
        (for 'pcaeof': covM[0] = tvar)"""
        pos, mass = _f64(pos), _f64(mass)
        n = len(pos)
        r = np.sqrt(pos[:, 0] * pos[:, 0] + pos[:, 1] * pos[:, 1])
        z = pos[:, 2].copy()
        phi = np.array([math.atan2(y, x) for x, y in zip(pos[:, 0], pos[:, 1])])
        seq = np.arange(n, dtype=np.uint64) if seq is None else np.ascontiguousarray(seq, dtype=np.uint64)
        M1, N = self.mmax + 1, self.norder
        covV, covM = np.zeros((sampT, M1, N)), np.zeros((sampT, M1, N, N))
        lib().ref_cyl_accumulate_pca(self.h, n, _dp(r), _dp(z), _dp(phi), _dp(mass), seq.ctypes.data_as(c_ulp),
                                     sampT, {"pcavar": 1, "pcaeof": 2}[which], _dp(covV), _dp(covM))
        return covV, covM

    def accumulated_eval(self, cos, sin, r, z, phi):
        cos, sin = _f64(cos), _f64(sin)
        o = [C.c_double() for _ in range(5)]
        lib().ref_cyl_accumulated_eval(self.h, _dp(cos), _dp(sin), float(r), float(z), float(phi),
                                       *[C.byref(x) for x in o])
        return tuple(x.value for x in o)       # p0, p, fr, fz, fp


def incr_position(pos, vel, dt, level=None, nlev=1, mlevel=-1, dim=3, nthreads=1):
    pos, vel = _f64(pos).copy(), _f64(vel).copy()
    lev = None if level is None else np.ascontiguousarray(level, dtype=np.uint32)
    lib().ref_incr_position(len(pos), _dp(pos), _dp(vel), lev.ctypes.data_as(c_up) if lev is not None else None,
                            nlev, float(dt), mlevel, dim, nthreads)
    return pos


def incr_velocity(pos, vel, acc, dt, level=None, nlev=1, mlevel=-1, dim=3, nthreads=1):
    pos, vel, acc = _f64(pos).copy(), _f64(vel).copy(), _f64(acc)
    lev = None if level is None else np.ascontiguousarray(level, dtype=np.uint32)
    lib().ref_incr_velocity(len(pos), _dp(pos), _dp(vel), _dp(acc),
                            lev.ctypes.data_as(c_up) if lev is not None else None, nlev, float(dt),
                            mlevel, dim, nthreads)
    return vel


def adjust_levels(vel, acc, pot, potext, scale, level, dtreq, dtime, multistep, mfirst_mdrft=0,
                  minlev=0, dynfracS=1.0, dynfracD=1.0e32, dynfracV=0.01, dynfracA=0.03, dynfracP=0.05,
                  shiftlevl=0, noswitch=False, first_call=False, dim=3, nthreads=1):
    """adjust_multistep_level_thread over every level >= minlev; returns (level, dtreq, nswitch,
    offlo, offhi)."""
    vel, acc, pot, potext = _f64(vel), _f64(acc), _f64(pot), _f64(potext)
    n = len(vel)
    sc = np.full(n, -1.0, dtype=np.float32) if scale is None else np.ascontiguousarray(scale, dtype=np.float32)
    lev = np.ascontiguousarray(level, dtype=np.uint32).copy()
    dq = np.ascontiguousarray(dtreq, dtype=np.float32).copy()
    P = LevelParams(dtime, dynfracS, dynfracD, dynfracV, dynfracA, dynfracP, multistep, shiftlevl,
                    mfirst_mdrft, minlev, int(noswitch), int(first_call), dim, nthreads)
    lo, hi = C.c_long(0), C.c_long(0)
    nsw = lib().ref_adjust_levels(n, _dp(vel), _dp(acc), _dp(pot), _dp(potext), sc.ctypes.data_as(c_fp),
                                  lev.ctypes.data_as(c_up), dq.ctypes.data_as(c_fp), C.byref(P),
                                  C.byref(lo), C.byref(hi))
    return lev, dq, nsw, lo.value, hi.value


class RefPyexp:
    """The reference's BasisClasses::Spherical (pyEXP's sphereSL basis: expui/BiorthBasis.cc:583-958,
    1034-1122) over the SLGridSph of a RefSphere: accumulate -> expcoef, sph/cyl/crt_eval -> the nine
    field values, computeAccel."""

    def __init__(self, sph: "RefSphere", scale=1.0, rmin=0.0, rmax=1.0e300, NO_L0=False, NO_L1=False,
                 EVEN_L=False, EVEN_M=False, M0_only=False):
        self.sph = sph                      # keeps the grid alive
        self.lmax, self.nmax = sph.lmax, sph.nmax
        bits = sum(1 << k for k, f in enumerate((NO_L0, NO_L1, EVEN_L, EVEN_M, M0_only)) if f)
        self.h = lib().ref_pyexp_create(sph.h, float(scale), float(rmin), float(rmax), bits)

    def __del__(self):
        try:
            lib().ref_pyexp_destroy(self.h)
        except Exception:
            pass

    def accumulate(self, mass, pos):
        mass, pos = _f64(mass), _f64(pos)
        out = np.zeros(((self.lmax + 1) ** 2, self.nmax))
        used = lib().ref_pyexp_accumulate(self.h, len(mass), _dp(mass), _dp(pos), _dp(out))
        return out, used

    def set_coefs(self, expcoef):
        e = _f64(expcoef)
        assert e.shape == ((self.lmax + 1) ** 2, self.nmax)
        lib().ref_pyexp_set_coefs(self.h, _dp(e))

    def eval(self, which, pts):
        """which: 'sph' (r, costh, phi), 'cyl' (R, z, phi) or 'crt' (x, y, z); -> [n][9]"""
        pts = _f64(pts).reshape(-1, 3)
        out = np.zeros((len(pts), 9))
        lib().ref_pyexp_eval(self.h, {"sph": 0, "cyl": 1, "crt": 2}[which], len(pts), _dp(pts), _dp(out))
        return out

    def accel(self, pos):
        pos = _f64(pos).reshape(-1, 3)
        out = np.zeros((len(pos), 3))
        lib().ref_pyexp_accel(self.h, len(pos), _dp(pos), _dp(out))
        return out


def fix_positions(pos, vel, acc, mass, level, nlev, mlevel=0, dim=3, center=(0.0, 0.0, 0.0), rtrunc=1.0e20):
    """fix_positions_thread (src/Component.cc:3280-3352), one thread -> [nlev][10] = mtot, com, cov, coa
    per level (rows below mlevel zero)"""
    pos, vel, acc, mass = _f64(pos), _f64(vel), _f64(acc), _f64(mass)
    lev = np.ascontiguousarray(level, dtype=np.uint32)
    ctr = _f64(center)
    out = np.zeros((nlev, 10))
    lib().ref_fix_positions(len(pos), _dp(pos), _dp(vel), _dp(acc), _dp(mass), lev.ctypes.data_as(c_up), nlev,
                            mlevel, dim, _dp(ctr), float(rtrunc), _dp(out))
    return out
