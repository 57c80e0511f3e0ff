"""ctypes view of the C++ host driver (exp_b200/host/, libexphost.so): the reference's
engine-side objects for the force loop -- Component, PotAccel (Sphere, Cylinder),
ComponentContainer, begin_run, do_step, adjust_multistep_level -- written in C++ over the C
ABI.  This module only assembles a run and reads results back; all stepping logic is C++."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import c_dp

_HOST = None


class SphereDesc(C.Structure):
    _fields_ = [("lmax", C.c_int), ("nmax", C.c_int), ("numr", C.c_int), ("cmap", C.c_int),
                ("rmin", C.c_double), ("rmax", C.c_double), ("rmap", C.c_double),
                ("ev", c_dp), ("ef", c_dp), ("p0", c_dp),
                ("scale", C.c_double), ("rmin_exp", C.c_double), ("rmax_exp", C.c_double),
                ("NO_L0", C.c_int), ("NO_L1", C.c_int), ("EVEN_L", C.c_int), ("EVEN_M", C.c_int),
                ("M0_only", C.c_int), ("self_consistent", C.c_int)]


class CylinderDesc(C.Structure):
    _fields_ = [("mmax", C.c_int), ("norder", C.c_int), ("numx", C.c_int), ("numy", C.c_int),
                ("cmapr", C.c_int), ("cmapz", C.c_int),
                ("ascale", C.c_double), ("hscale", C.c_double), ("rmin", C.c_double), ("rmax", C.c_double),
                ("potC", c_dp), ("potS", c_dp), ("rforceC", c_dp), ("rforceS", c_dp),
                ("zforceC", c_dp), ("zforceS", c_dp),
                ("rcylmax", C.c_double), ("acyl", C.c_double),
                ("EVEN_M", C.c_int), ("MMIN", C.c_int), ("MLIM", C.c_int), ("NMIN", C.c_int),
                ("NLIM", C.c_int), ("self_consistent", C.c_int), ("body", c_dp), ("orig", c_dp)]


def host():
    global _HOST
    if _HOST is None:
        _lib.lib()                                   # libexpb200.so first (fails loudly if missing)
        import os
        if not os.path.exists(_lib.HOST_LIB_PATH):
            raise _lib.Expb200Error(f"{_lib.HOST_LIB_PATH} is missing: run __graft_entry__.build()")
        H = C.CDLL(_lib.HOST_LIB_PATH)
        H.exphost_last_error.restype = C.c_char_p
        H.exphost_set_params.argtypes = [C.c_double, C.c_int] + [C.c_double] * 5 + [C.c_int]
        H.exphost_set_comm.argtypes = [C.c_void_p, C.c_int, C.c_int]
        H.exphost_add_component.argtypes = [C.c_char_p, C.c_size_t, c_dp, c_dp, c_dp, C.c_int,
                                            C.POINTER(C.c_int)]
        H.exphost_component_options.argtypes = [C.c_int, c_dp, C.c_double, C.c_int, C.c_int, C.c_int]
        H.exphost_set_sphere.argtypes = [C.c_int, C.POINTER(SphereDesc)]
        H.exphost_set_cylinder.argtypes = [C.c_int, C.POINTER(CylinderDesc)]
        H.exphost_add_interaction.argtypes = [C.c_int, C.c_int]
        H.exphost_component_global.argtypes = [C.c_int, C.c_uint64, C.c_uint64]
        H.exphost_sphere_pca_options.argtypes = [C.c_int] * 9 + [C.c_double] * 4
        H.exphost_sphere_get_covariance.argtypes = [C.c_int, C.POINTER(C.c_int), c_dp, _lib.c_up, c_dp, c_dp,
                                                    c_dp, c_dp, c_dp, c_dp, c_dp]
        H.exphost_cylinder_cov_options.argtypes = [C.c_int] * 7
        H.exphost_cylinder_get_covariance.argtypes = [C.c_int, C.POINTER(C.c_int), _lib.c_up] + [c_dp] * 6
        H.exphost_pca_hall.argtypes = [C.c_int] * 6 + [C.c_double] * 4 + [c_dp] * 9
        H.exphost_cylinder_pca_options.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_int]
        H.exphost_cylinder_get_pca.argtypes = [C.c_int, C.POINTER(C.c_int)] + [c_dp] * 6
        H.exphost_cyl_pca_hall.argtypes = ([C.c_int] * 4 + [C.c_double, C.c_int, C.c_double, C.c_double, _lib.c_up]
                                           + [c_dp] * 9)
        H.exphost_do_steps.argtypes = [C.c_int]
        H.exphost_do_steps_timed.argtypes = [C.c_int, c_dp]
        H.exphost_get_particles.argtypes = [C.c_int, c_dp, c_dp, c_dp, c_dp, _lib.c_up, _lib.c_fp]
        H.exphost_get_sphere_coefs.argtypes = [C.c_int, C.c_int, C.c_int, c_dp]
        H.exphost_get_cylinder_coefs.argtypes = [C.c_int, c_dp, c_dp, c_dp]
        H.exphost_level_report.argtypes = [C.c_int] + [_lib.c_u64p] * 5
        H.exphost_sort_counters.argtypes = [C.c_int, _lib.c_u64p]
        H.exphost_sort_counters.restype = C.c_int
        H.exphost_state.argtypes = [c_dp, C.POINTER(C.c_int), c_dp, c_dp]
        H.exphost_get_com.argtypes = [C.c_int, c_dp, c_dp]
        H.exphost_dump_coefs.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_long]
        H.exphost_dump_coefs.restype = C.c_long
        _HOST = H
    return _HOST


def _check(rc):
    if rc != 0:
        raise _lib.Expb200Error(f"exphost error {rc}: {host().exphost_last_error().decode()}")


def _p(a):
    return a.ctypes.data_as(c_dp)


class Simulation:
    """One run of the C++ host driver (the process-wide `comp` container of the reference)."""

    def __init__(self, dtime, multistep=0, dynfracS=1.0, dynfracD=1.0e32, dynfracV=0.01,
                 dynfracA=0.03, dynfracP=0.05, shiftlevl=0, comm=None, myid=0, numprocs=1, device=0):
        self.H = host()
        self.device, self.multistep = device, multistep
        self._keep = []
        self.kinds = []
        self.meta = []
        _check(self.H.exphost_reset())
        _check(self.H.exphost_set_params(dtime, multistep, dynfracS, dynfracD, dynfracV, dynfracA,
                                         dynfracP, shiftlevl))
        _check(self.H.exphost_set_comm(comm, myid, numprocs))

    def add_component(self, name, pos, vel, mass, center=None, rtrunc=1.0e20, noswitch=False,
                      freezeLev=False, dtreset=True):
        pos, vel, mass = (np.ascontiguousarray(a, dtype=np.float64) for a in (pos, vel, mass))
        idx = C.c_int(-1)
        _check(self.H.exphost_add_component(name.encode(), len(pos), _p(pos), _p(vel), _p(mass),
                                            self.device, C.byref(idx)))
        ctr = None if center is None else np.ascontiguousarray(center, dtype=np.float64)
        _check(self.H.exphost_component_options(idx.value, None if ctr is None else _p(ctr), rtrunc,
                                                int(noswitch), int(freezeLev), int(dtreset)))
        self.kinds.append(None)
        self.meta.append(dict(name=name, center=np.zeros(3) if ctr is None else ctr.copy()))
        return idx.value

    def set_sphere(self, idx, tab, scale=1.0, rmin=-1.0, rmax=-1.0, NO_L0=False, NO_L1=False,
                   EVEN_L=False, EVEN_M=False, M0_only=False, self_consistent=True):
        arr = [np.ascontiguousarray(getattr(tab, k), dtype=np.float64) for k in ("ev", "ef", "p0")]
        self._keep.append(arr)
        d = SphereDesc(tab.lmax, tab.nmax, tab.numr, tab.cmap, tab.rmin, tab.rmax, tab.rmap,
                       _p(arr[0]), _p(arr[1]), _p(arr[2]), scale, rmin, rmax, int(NO_L0), int(NO_L1),
                       int(EVEN_L), int(EVEN_M), int(M0_only), int(self_consistent))
        _check(self.H.exphost_set_sphere(idx, C.byref(d)))
        self.kinds[idx] = ("sphere", (tab.lmax + 1) ** 2, tab.nmax)
        self.meta[idx].update(geometry="sphere", forceID="SphereSL", lmax=tab.lmax, nmax=tab.nmax, scale=scale)

    def set_cylinder(self, idx, tab, rcylmax=20.0, acyl=-1.0, EVEN_M=False, MMIN=0, MLIM=2 ** 30,
                     NMIN=0, NLIM=2 ** 30, self_consistent=True, body=None, orig=None):
        arr = [np.ascontiguousarray(getattr(tab, k), dtype=np.float64)
               for k in ("potC", "potS", "rforceC", "rforceS", "zforceC", "zforceS")]
        rot = [None if body is None else np.ascontiguousarray(body, dtype=np.float64),
               None if orig is None else np.ascontiguousarray(orig, dtype=np.float64)]
        self._keep.append(arr + rot)
        d = CylinderDesc(tab.mmax, tab.norder, tab.numx, tab.numy, tab.cmapr, tab.cmapz, tab.ascale,
                         tab.hscale, tab.rmin, tab.rmax, *[_p(a) for a in arr], rcylmax, acyl,
                         int(EVEN_M), MMIN, MLIM, NMIN, NLIM, int(self_consistent),
                         None if rot[0] is None else _p(rot[0]), None if rot[1] is None else _p(rot[1]))
        _check(self.H.exphost_set_cylinder(idx, C.byref(d)))
        self.kinds[idx] = ("cylinder", tab.mmax + 1, tab.norder)
        self.meta[idx].update(geometry="cylinder", forceID="Cylinder", mmax=tab.mmax, nmax=tab.norder,
                              rotation=body)

    def component_global(self, idx, indx_base, nbodies_tot):
        """sharded runs: this rank holds global particles [indx_base, indx_base + n)"""
        _check(self.H.exphost_component_global(idx, int(indx_base), int(nbodies_tot)))

    TK = {"Hall": 0, "VarianceCut": 1, "CumulativeCut": 2, "VarianceWeighted": 3, "None": 4}

    def sphere_pca_options(self, idx, npca=2 ** 31 - 1, npca0=2 ** 31 - 1, nint=0, pcavar=False,
                           pcaeof=False, samplesz=1, subsamp=False, tk_type="None", hexp=1.0, snr=1.0,
                           tksmooth=3.0, tkcum=0.95):
        """the covariance / PCA keys of the sphereSL force (src/AxisymmetricBasis.cc:64-92)"""
        _check(self.H.exphost_sphere_pca_options(idx, npca, npca0, nint, int(pcavar), int(pcaeof),
                                                 samplesz, int(subsamp), self.TK[tk_type], hexp, snr,
                                                 tksmooth, tkcum))

    def sphere_covariance(self, idx):
        """massT, countT, expcoefT, expcoefM, tvar and the Hall factors after the last compute step"""
        _, nrow, nmax = self.kinds[idx]
        L = int(round(np.sqrt(nrow))) - 1
        ldim = (L + 1) * (L + 2) // 2
        sT, mu = C.c_int(0), C.c_double(0.0)
        _check(self.H.exphost_sphere_get_covariance(idx, C.byref(sT), None, None, None, None, None, None,
                                                    None, None, C.byref(mu)))
        T = sT.value
        if T == 0:
            return None
        out = dict(sampT=T, muse=mu.value, massT=np.zeros(T), countT=np.zeros(T, dtype=np.uint32),
                   coefT=np.zeros((T, ldim, nmax, 2)), coefM=np.zeros((T, ldim, nmax, nmax)),
                   tvar=np.zeros((ldim, nmax, nmax)), b_Hall=np.zeros((ldim, nmax)),
                   s_Hall=np.ones((ldim, nmax)), weight=np.ones((ldim, nmax)))
        _check(self.H.exphost_sphere_get_covariance(
            idx, C.byref(sT), _p(out["massT"]), out["countT"].ctypes.data_as(_lib.c_up), _p(out["coefT"]),
            _p(out["coefM"]), _p(out["tvar"]), _p(out["b_Hall"]), _p(out["s_Hall"]), _p(out["weight"]),
            C.byref(mu)))
        out["coefT"] = out["coefT"][..., 0] + 1j * out["coefT"][..., 1]
        return out

    def cylinder_cov_options(self, idx, nint=0, npca=2 ** 30, npca0=2 ** 30, pcavar=False, pcaeof=False,
                             samplesz=1):
        _check(self.H.exphost_cylinder_cov_options(idx, nint, npca, npca0, int(pcavar), int(pcaeof), samplesz))

    CYL_TK = {"Hall": 0, "Truncate": 1, "None": 2}

    def cylinder_pca_options(self, idx, tk_type="None", snr=1.0, hexp=1.0, subsamp=False):
        """Hall smoothing switches of the cylinder force (src/Cylinder.cc:231-241, 469)"""
        _check(self.H.exphost_cylinder_pca_options(idx, self.CYL_TK[tk_type], snr, hexp, int(subsamp)))

    def cylinder_pca(self, idx):
        """the last EmpCylSL::pca_hall analysis: meanJK, evalJK, ratio [m][n], evecJK [m][n][n], and the
        coefficients the last smoothing started from (raw_cos, raw_sin; zero-size if none was applied)"""
        _, M1, N = self.kinds[idx]
        out = dict(meanJK=np.zeros((M1, N)), evalJK=np.zeros((M1, N)), ratio=np.zeros((M1, N)),
                   evecJK=np.zeros((M1, N, N)), raw_cos=np.zeros((M1, N)), raw_sin=np.zeros((M1, N)))
        ok = C.c_int(0)
        _check(self.H.exphost_cylinder_get_pca(idx, C.byref(ok), _p(out["meanJK"]), _p(out["evalJK"]),
                                               _p(out["ratio"]), _p(out["evecJK"]), _p(out["raw_cos"]),
                                               _p(out["raw_sin"])))
        return out if ok.value else None

    def cylinder_covariance(self, idx, covar=True, pcavar=False, pcaeof=False):
        _, M1, N = self.kinds[idx]
        sT = C.c_int(0)
        _check(self.H.exphost_cylinder_get_covariance(idx, C.byref(sT), None, None, None, None, None, None, None))
        T = sT.value
        if T == 0:
            return None
        out = dict(sampT=T, numbT=np.zeros(T, dtype=np.uint32), massT=np.zeros(T))
        if covar:
            out["VC"], out["MV"] = np.zeros((T, M1, N, 2)), np.zeros((T, M1, N, N, 2))
        if pcavar:
            out["covV"], out["covM"] = np.zeros((T, M1, N)), np.zeros((T, M1, N, N))
        if pcaeof:
            out["tvar"] = np.zeros((M1, N, N))
        g = lambda k: _p(out[k]) if k in out else None
        _check(self.H.exphost_cylinder_get_covariance(idx, C.byref(sT), out["numbT"].ctypes.data_as(_lib.c_up),
                                                      _p(out["massT"]), g("VC"), g("MV"), g("covV"), g("covM"),
                                                      g("tvar")))
        if covar:
            out["VC"] = out["VC"][..., 0] + 1j * out["VC"][..., 1]
            out["MV"] = out["MV"][..., 0] + 1j * out["MV"][..., 1]
        return out

    def add_interaction(self, frm, on):
        _check(self.H.exphost_add_interaction(frm, on))

    def begin_run(self):
        _check(self.H.exphost_begin_run())

    def do_steps(self, n=1):
        _check(self.H.exphost_do_steps(n))

    def do_steps_timed(self, n):
        ms = C.c_double(0.0)
        _check(self.H.exphost_do_steps_timed(n, C.byref(ms)))
        return ms.value

    def particles(self, idx, n):
        out = dict(pos=np.empty((n, 3)), vel=np.empty((n, 3)), acc=np.empty((n, 3)), pot=np.empty(n),
                   level=np.empty(n, dtype=np.uint32), dtreq=np.empty(n, dtype=np.float32))
        _check(self.H.exphost_get_particles(idx, _p(out["pos"]), _p(out["vel"]), _p(out["acc"]),
                                            _p(out["pot"]), out["level"].ctypes.data_as(_lib.c_up),
                                            out["dtreq"].ctypes.data_as(_lib.c_fp)))
        return out

    def sphere_coefs(self, idx, which=3, M=0):
        _, nrow, nmax = self.kinds[idx]
        out = np.empty((nrow, nmax))
        _check(self.H.exphost_get_sphere_coefs(idx, which, M, _p(out)))
        return out

    def cylinder_coefs(self, idx):
        _, m1, n = self.kinds[idx]
        cs, sn = np.empty((m1, n)), np.empty((m1, n))
        mass = C.c_double(0.0)
        _check(self.H.exphost_get_cylinder_coefs(idx, _p(cs), _p(sn), C.byref(mass)))
        return cs, sn, mass.value

    def level_report(self, idx):
        v = [C.c_uint64(0) for _ in range(4)]
        counts = np.zeros(self.multistep + 1, dtype=np.uint64)
        _check(self.H.exphost_level_report(idx, *[C.byref(x) for x in v],
                                           counts.ctypes.data_as(_lib.c_u64p)))
        return dict(examined=v[0].value, switched=v[1].value, offlo=v[2].value, offhi=v[3].value,
                    counts=counts)

    def sort_counters(self, idx):
        """device binning diagnostics of component idx: [whole-component sorts, tail sorts of the active
        levels, particles moved by the tail sorts, particles moved by the whole-component sorts]"""
        out = np.zeros(4, dtype=np.uint64)
        _check(self.H.exphost_sort_counters(idx, out.ctypes.data_as(_lib.c_u64p)))
        return out

    def dump_coefs(self, idx, kind=0):
        """bytes of the force's dump_coefs stream (kind 0: NewCoefs, 1: old style) or of its
        SphStruct/CylStruct complex array (kind 2)"""
        n = self.H.exphost_dump_coefs(idx, kind, None, 0)
        if n < 0:
            _check(int(n))
        buf = C.create_string_buffer(int(n))
        n2 = self.H.exphost_dump_coefs(idx, kind, buf, n)
        if n2 < 0:
            _check(int(n2))
        return buf.raw[:n2]

    def coef_struct(self, idx):
        """the force's SphStruct / CylStruct array: complex ((lmax+1)(lmax+2)/2 | mmax+1, nmax), cosine
        coefficient in the real part, sine in the imaginary part"""
        m = self.meta[idx]
        rows = (m["lmax"] + 1) * (m["lmax"] + 2) // 2 if m["geometry"] == "sphere" else m["mmax"] + 1
        return np.frombuffer(self.dump_coefs(idx, kind=2), dtype=np.complex128).reshape(rows, m["nmax"]).copy()

    def dump_coefs_h5(self, idx, path, config=""):
        """SphericalBasis::dump_coefs_h5 / Cylinder::dump_coefs_h5 (reference src/SphericalBasis.cc:
        1909-1975, src/Cylinder.cc:1625-1689): create the HDF5 coefficient file or append the current
        coefficients to it as one more snapshot, in the layout CoefClasses::{Sph,Cyl}Coefs write
        through HighFive (exp_b200/h5lite.py: dump_coefs_h5)."""
        from . import h5lite
        m = self.meta[idx]
        kw = dict(lmax=m["lmax"], scale=m["scale"]) if m["geometry"] == "sphere" else dict(mmax=m["mmax"])
        return h5lite.dump_coefs_h5(path, m["geometry"], m["name"], m["forceID"], self.state()["tnow"],
                                    self.coef_struct(idx), config=config, center=m["center"],
                                    rotation=m.get("rotation"), **kw)

    def com(self, idx):
        """[mtot, com[3], cov[3], coa[3]] of component idx after its last fix_positions, and the
        per-level sums it keeps."""
        out, lev = np.zeros(10), np.zeros((self.multistep + 1, 10))
        _check(self.H.exphost_get_com(idx, _p(out), _p(lev)))
        return out, lev

    def state(self):
        t, k = C.c_double(0.0), C.c_int(0)
        pc, vc = np.zeros(self.multistep + 1), np.zeros(self.multistep + 1)
        _check(self.H.exphost_state(C.byref(t), C.byref(k), _p(pc), _p(vc)))
        return dict(tnow=t.value, this_step=k.value, pos_check=pc, vel_check=vc)


def cyl_pca_hall(mmax, norder, numbT, massT, covV, covM, nbodstot, subsamp=False, tk_type="None", snr=1.0,
                 hexp=1.0, cosv=None, sinv=None):
    """EmpCylSL::pca_hall (PCAVAR part) + set_trimmed (exputil/EmpCylSL.cc:4577-4970, 5141-5207) on given
    partition arrays; host-only.  covV [T][mmax+1][norder], covM [T][mmax+1][norder][norder]."""
    H = host()
    T, M1, N = len(massT), mmax + 1, norder
    nb = np.ascontiguousarray(numbT, dtype=np.uint32)
    mT, cV, cM = (np.ascontiguousarray(a, dtype=np.float64) for a in (massT, covV, covM))
    out = dict(meanJK=np.zeros((M1, N)), evalJK=np.zeros((M1, N)), ratio=np.zeros((M1, N)),
               evecJK=np.zeros((M1, N, N)))
    cs = None if cosv is None else np.ascontiguousarray(cosv, dtype=np.float64).copy()
    sn = None if sinv is None else np.ascontiguousarray(sinv, dtype=np.float64).copy()
    _check(H.exphost_cyl_pca_hall(mmax, norder, T, int(subsamp), float(nbodstot), Simulation.CYL_TK[tk_type],
                                  snr, hexp, nb.ctypes.data_as(_lib.c_up), _p(mT), _p(cV), _p(cM),
                                  _p(out["meanJK"]), _p(out["evalJK"]), _p(out["ratio"]), _p(out["evecJK"]),
                                  None if cs is None else _p(cs), None if sn is None else _p(sn)))
    out["cos"], out["sin"] = cs, sn
    return out


def pca_hall(Lmax, nmax, massT, coefT, coefM, used, subsamp=False, tk_type="None", hexp=1.0, snr=1.0,
             tksmooth=3.0, tkcum=0.95, expcoef=None):
    """AxisymmetricBasis::pca_hall_3d + smoothing (src/AxisymmetricBasis.cc:198-660) on given subsample
    arrays; host-only.  coefT complex [T][ldim][nmax], coefM real [T][ldim][nmax][nmax]."""
    H = host()
    T = len(massT)
    ldim = (Lmax + 1) * (Lmax + 2) // 2
    cT = np.ascontiguousarray(np.stack([np.real(coefT), np.imag(coefT)], axis=-1), dtype=np.float64)
    cM = np.ascontiguousarray(coefM, dtype=np.float64)
    mT = np.ascontiguousarray(massT, dtype=np.float64)
    out = dict(b_Hall=np.zeros((ldim, nmax)), s_Hall=np.zeros((ldim, nmax)), weight=np.zeros((ldim, nmax)),
               evalJK=np.zeros((ldim, nmax)), evec=np.zeros((ldim, nmax, nmax)))
    ex = None if expcoef is None else np.ascontiguousarray(expcoef, dtype=np.float64).copy()
    _check(H.exphost_pca_hall(Lmax, nmax, T, int(used), int(subsamp), Simulation.TK[tk_type], hexp, snr,
                              tksmooth, tkcum, _p(mT), _p(cT), _p(cM), _p(out["b_Hall"]), _p(out["s_Hall"]),
                              _p(out["weight"]), _p(out["evalJK"]), _p(out["evec"]),
                              None if ex is None else _p(ex)))
    out["expcoef"] = ex
    return out
