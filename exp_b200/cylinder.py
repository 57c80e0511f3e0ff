"""Host-side mirror of the reference's ``Cylinder`` force (force id ``cylinder``) over the C
ABI; method names follow ``PotAccel`` (reference src/PotAccel.H:173-288)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import c_dp, check
from .component import Component

BIG = 2 ** 30


class Cylinder:
    def __init__(self, table, *, rcylmax=20.0, acyl=None, EVEN_M=False, MMIN=0, MLIM=BIG, NMIN=0,
                 NLIM=BIG, multistep=0, device=0):
        """``table``: exp_b200.cyltable.CylTable (or EXP's own EmpCylSL grids in that layout).
        Defaults follow src/Cylinder.cc:105-112 (rcylmax=20, acyl = table ascale)."""
        t = table
        self.mmax, self.norder = t.mmax, t.norder
        self.multistep = multistep
        self.mlevel = 0
        self.device = device
        cfg = _lib.CylinderConfig(t.mmax, t.norder, t.numx, t.numy, t.cmapr, t.cmapz, t.ascale,
                                  t.hscale, t.rmin, t.rmax, rcylmax,
                                  t.ascale if acyl is None else acyl, int(EVEN_M), MMIN, MLIM,
                                  NMIN, NLIM, multistep)
        arrs = [np.ascontiguousarray(getattr(t, k), dtype=np.float64)
                for k in ("potC", "potS", "rforceC", "rforceS", "zforceC", "zforceS")]
        self._h = C.c_void_p()
        check(_lib.lib().expb200_cylinder_create(device, C.byref(cfg),
                                                 *[a.ctypes.data_as(c_dp) for a in arrs],
                                                 C.byref(self._h)))
        self.used = 0
        self.cylmass = 0.0

    def close(self):
        if self._h:
            _lib.lib().expb200_cylinder_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_multistep_level(self, n):
        self.mlevel = int(n)

    def set_owner(self, c):
        """PotAccel::component: the component whose centre is the expansion origin when the
        force is evaluated on another component's particles (use_external)."""
        check(_lib.lib().expb200_cylinder_set_owner(self._h, c._h if c is not None else None))

    def set_orientation(self, body=None, orig=None):
        if body is None:
            check(_lib.lib().expb200_cylinder_set_orientation(self._h, None, None))
        else:
            b = np.ascontiguousarray(body, dtype=np.float64)
            o = np.ascontiguousarray(orig, dtype=np.float64)
            check(_lib.lib().expb200_cylinder_set_orientation(self._h, b.ctypes.data_as(c_dp),
                                                              o.ctypes.data_as(c_dp)))

    def determine_coefficients(self, c: Component, adb=1.0, to_host=True):
        """returns (cosN, sinN) [(mmax+1), norder]; sets self.used, self.cylmass (this level)"""
        shape = (self.mmax + 1, self.norder)
        cs, sn = (np.zeros(shape), np.zeros(shape)) if to_host else (None, None)
        used = C.c_uint64(0)
        mass = C.c_double(0.0)
        check(_lib.lib().expb200_cylinder_determine_coefficients(
            self._h, c._h, self.mlevel, float(adb),
            cs.ctypes.data_as(c_dp) if to_host else None,
            sn.ctypes.data_as(c_dp) if to_host else None, C.byref(used), C.byref(mass)))
        self.used, self.cylmass = used.value, mass.value
        return cs, sn

    def covariance(self, c: Component, sampT: int, adb: float = 1.0, indx_base: int = 0):
        """The `covar` accumulation of EmpCylSL::accumulate (exputil/EmpCylSL.cc:4123-4141) for
        levlist[mlevel]: per jackknife bin T = indx % sampT returns numbT, massT, VC (complex,
        [T][m][n]) and MV (complex Hermitian, [T][m][n][o])."""
        M1, N = self.mmax + 1, self.norder
        numbT = np.zeros(sampT, dtype=np.uint32)
        massT = np.zeros(sampT)
        VC = np.zeros((sampT, M1, N, 2))
        MV = np.zeros((sampT, M1, N, N, 2))
        check(_lib.lib().expb200_cylinder_covariance(
            self._h, c._h, self.mlevel, float(adb), int(sampT), int(indx_base),
            numbT.ctypes.data_as(_lib.c_up), massT.ctypes.data_as(c_dp), VC.ctypes.data_as(c_dp),
            MV.ctypes.data_as(c_dp)))
        return dict(numbT=numbT, massT=massT, VC=VC[..., 0] + 1j * VC[..., 1],
                    MV=MV[..., 0] + 1j * MV[..., 1])

    def covariance_pca(self, c: Component, sampT: int, which: str, adb: float = 1.0, indx_base: int = 0):
        """The PCAVAR ('pcavar') / PCAEOF ('pcaeof') accumulations of EmpCylSL::accumulate
        (exputil/EmpCylSL.cc:4089-4119): covV [T][m][n], covM [T][m][n][n] (real)."""
        M1, N = self.mmax + 1, self.norder
        numbT = np.zeros(sampT, dtype=np.uint32)
        massT = np.zeros(sampT)
        covV, covM = np.zeros((sampT, M1, N)), np.zeros((sampT, M1, N, N))
        check(_lib.lib().expb200_cylinder_covariance_pca(
            self._h, c._h, self.mlevel, float(adb), int(sampT), int(indx_base),
            {"pcavar": 1, "pcaeof": 2}[which], numbT.ctypes.data_as(_lib.c_up), massT.ctypes.data_as(c_dp),
            covV.ctypes.data_as(c_dp), covM.ctypes.data_as(c_dp)))
        return dict(numbT=numbT, massT=massT, covV=covV, covM=covM)

    def set_coefficients(self, cos=None, sin=None, cylmass=-1.0):
        if cos is None:
            check(_lib.lib().expb200_cylinder_set_coefficients(self._h, None, None, float(cylmass)))
        else:
            a = np.ascontiguousarray(cos, dtype=np.float64)
            b = np.ascontiguousarray(sin, dtype=np.float64)
            check(_lib.lib().expb200_cylinder_set_coefficients(self._h, a.ctypes.data_as(c_dp),
                                                               b.ctypes.data_as(c_dp), float(cylmass)))

    def get_acceleration_and_potential(self, c: Component):
        check(_lib.lib().expb200_cylinder_determine_acceleration(self._h, c._h, self.mlevel))

    def multistep_commit(self, mlevel=None):
        check(_lib.lib().expb200_cylinder_multistep_commit(self._h, self.mlevel if mlevel is None else mlevel))

    def compute_multistep_coefficients(self, mdrft, cylmass=-1.0):
        check(_lib.lib().expb200_cylinder_compute_multistep_coefficients(self._h, int(mdrft), float(cylmass)))

    def multistep_update(self, c: Component, adb=1.0, finish=True):
        """Move the contribution of every particle whose level changed in the last
        adjust_multistep_level from its old to its new level table."""
        check(_lib.lib().expb200_cylinder_multistep_update(self._h, c._h, float(adb), int(finish)))

    def multistep_update_finish(self):
        check(_lib.lib().expb200_cylinder_multistep_update_finish(self._h))

    def level_coefficients(self, which, mlevel):
        """(cos, sin) of level mlevel: which = "N" (cosN/sinN), "L" or "differ"."""
        shape = (self.mmax + 1, self.norder)
        cs, sn = np.zeros(shape), np.zeros(shape)
        check(_lib.lib().expb200_cylinder_get_level_coefficients(
            self._h, {"N": 0, "L": 1, "differ": 2}[which], int(mlevel), cs.ctypes.data_as(c_dp),
            sn.ctypes.data_as(c_dp)))
        return cs, sn

    def coef_device(self):
        p = c_dp()
        n = C.c_size_t(0)
        st = C.c_void_p()
        check(_lib.lib().expb200_cylinder_coef_device(self._h, C.byref(p), C.byref(n), C.byref(st)))
        return p, n.value, st
