"""Host-side mirror of the reference's ``Sphere`` / ``SphericalBasis`` force
(force id ``sphereSL``) over the C ABI.  Method names and argument meaning follow
``PotAccel`` (reference src/PotAccel.H:173-288): ``determine_coefficients``,
``get_acceleration_and_potential``, ``set_multistep_level``."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import c_dp, check
from .component import Component


class Sphere:
    def __init__(self, table, *, scale=1.0, rmin=None, rmax=None, NO_L0=False, NO_L1=False,
                 EVEN_L=False, EVEN_M=False, M0_only=False, multistep=0, device=0):
        """``table``: an object with the SLGridSph fields lmax, nmax, numr, cmap, rmin, rmax,
        rmap, ev, ef, p0 (exp_b200.slgrid.SLGridSphTable or EXP's own arrays).  rmin/rmax
        default to the table's, as the Sphere constructor does (src/Sphere.cc:65-66)."""
        self.lmax, self.nmax = table.lmax, table.nmax
        self.device = device
        self.multistep = multistep
        self.mlevel = 0
        cfg = _lib.SphereConfig(table.lmax, table.nmax, table.numr, table.cmap, table.rmin,
                                table.rmax, table.rmap, scale,
                                table.rmin if rmin is None else rmin,
                                table.rmax if rmax is None else rmax,
                                int(NO_L0), int(NO_L1), int(EVEN_L), int(EVEN_M), int(M0_only),
                                multistep)
        ev = np.ascontiguousarray(table.ev, dtype=np.float64)
        ef = np.ascontiguousarray(table.ef, dtype=np.float64)
        p0 = np.ascontiguousarray(table.p0, dtype=np.float64)
        self._h = C.c_void_p()
        check(_lib.lib().expb200_sphere_create(device, C.byref(cfg), ev.ctypes.data_as(c_dp),
                                               ef.ctypes.data_as(c_dp), p0.ctypes.data_as(c_dp),
                                               C.byref(self._h)))
        self.nrow = (self.lmax + 1) ** 2
        self.expcoef = np.zeros((self.nrow, self.nmax))
        self.used = 0

    def close(self):
        if self._h:
            _lib.lib().expb200_sphere_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_multistep_level(self, n: int):
        self.mlevel = int(n)

    def set_owner(self, c: Component | None):
        """PotAccel::component: the component whose centre is the expansion origin when the
        force is evaluated on another component's particles (use_external)."""
        check(_lib.lib().expb200_sphere_set_owner(self._h, c._h if c is not None else None))

    def determine_coefficients(self, c: Component, adb: float = 1.0, to_host: bool = True):
        """Accumulate levlist[mlevel] of ``c``; returns expcoef0 (rows x nmax)."""
        used = C.c_uint64(0)
        coef = np.empty((self.nrow, self.nmax)) if to_host else None
        check(_lib.lib().expb200_sphere_determine_coefficients(
            self._h, c._h, self.mlevel, float(adb),
            coef.ctypes.data_as(c_dp) if to_host else None, C.byref(used)))
        self.used = used.value
        return coef

    def covariance(self, c: Component, sampT: int, adb: float = 1.0, indx_base: int = 0):
        """The pcavar accumulation (src/SphericalBasis.cc:501-581) of levlist[mlevel]: per
        jackknife bin T = indx % sampT returns countT, massT, coefT (complex, [T][ldim][nmax])
        and coefM ([T][ldim][nmax][nmax]), ldim index l(l+1)/2 + m."""
        ldim = (self.lmax + 1) * (self.lmax + 2) // 2
        countT = np.zeros(sampT, dtype=np.uint32)
        massT = np.zeros(sampT)
        coefT = np.zeros((sampT, ldim, self.nmax, 2))
        coefM = np.zeros((sampT, ldim, self.nmax, self.nmax))
        check(_lib.lib().expb200_sphere_covariance(
            self._h, c._h, self.mlevel, float(adb), int(sampT), int(indx_base),
            countT.ctypes.data_as(_lib.c_up), massT.ctypes.data_as(c_dp), coefT.ctypes.data_as(c_dp),
            coefM.ctypes.data_as(c_dp)))
        return dict(countT=countT, massT=massT, coefT=coefT[..., 0] + 1j * coefT[..., 1], coefM=coefM)

    def covariance_eof(self, c: Component, adb: float = 1.0):
        """The pcaeof accumulation tvar[ldim][nmax][nmax] of levlist[mlevel]
        (src/SphericalBasis.cc:536-545, 573-581)."""
        ldim = (self.lmax + 1) * (self.lmax + 2) // 2
        tvar = np.zeros((ldim, self.nmax, self.nmax))
        check(_lib.lib().expb200_sphere_covariance_eof(self._h, c._h, self.mlevel, float(adb),
                                                       tvar.ctypes.data_as(c_dp)))
        return tvar

    def set_coefficients(self, coef=None):
        """HtoD_coefs: install the coefficients the force evaluation uses (None: the
        vector accumulated on the device)."""
        if coef is None:
            check(_lib.lib().expb200_sphere_set_coefficients(self._h, None))
        else:
            a = np.ascontiguousarray(coef, dtype=np.float64)
            assert a.size == self.nrow * self.nmax
            self.expcoef = a.reshape(self.nrow, self.nmax).copy()
            check(_lib.lib().expb200_sphere_set_coefficients(self._h, a.ctypes.data_as(c_dp)))

    def get_acceleration_and_potential(self, c: Component):
        """Adds acc/pot to particles of ``c`` with level >= mlevel."""
        check(_lib.lib().expb200_sphere_determine_acceleration(self._h, c._h, self.mlevel))

    def step(self, c: Component, dt: float, adb: float = 1.0, comm=None):
        check(_lib.lib().expb200_sphere_step(self._h, c._h, float(dt), float(adb), comm))

    def steps(self, c: Component, dt: float, nsteps: int, adb: float = 1.0, comm=None):
        """nsteps single-level KDK steps back to back (same arithmetic as step() in a loop)."""
        check(_lib.lib().expb200_sphere_steps(self._h, c._h, float(dt), int(nsteps), float(adb), comm))

    def coef_device(self):
        p = c_dp()
        st = C.c_void_p()
        check(_lib.lib().expb200_sphere_coef_device(self._h, C.byref(p), C.byref(st)))
        return p, st

    # ---- multistep tables (device resident) ---------------------------------
    def multistep_commit(self, mlevel=None):
        check(_lib.lib().expb200_sphere_multistep_commit(self._h, self.mlevel if mlevel is None else mlevel))

    def compute_multistep_coefficients(self, mdrft: int):
        check(_lib.lib().expb200_sphere_compute_multistep_coefficients(self._h, int(mdrft)))

    def multistep_update(self, c: Component, adb: float = 1.0, finish: bool = True):
        check(_lib.lib().expb200_sphere_multistep_update(self._h, c._h, float(adb), int(finish)))

    def multistep_update_finish(self):
        check(_lib.lib().expb200_sphere_multistep_update_finish(self._h))

    def differ_device(self):
        p = c_dp()
        n = C.c_size_t(0)
        st = C.c_void_p()
        check(_lib.lib().expb200_sphere_differ_device(self._h, C.byref(p), C.byref(n), C.byref(st)))
        return p, n.value, st

    def level_coefficients(self, which: str, mlevel: int):
        out = np.empty((self.nrow, self.nmax))
        w = {"N": 0, "L": 1, "differ": 2, "use": 3, "accum": 4}[which]    # use: installed for the force
        check(_lib.lib().expb200_sphere_get_level_coefficients(self._h, w, mlevel, out.ctypes.data_as(c_dp)))
        return out

    def set_level_coefficients(self, which: str, mlevel: int, coef):
        a = np.ascontiguousarray(coef, dtype=np.float64)
        w = {"N": 0, "L": 1}[which]
        check(_lib.lib().expb200_sphere_set_level_coefficients(self._h, w, mlevel, a.ctypes.data_as(c_dp)))
