"""pyEXP-facing twins of the path: ``basis.createFromArray`` / ``getFields`` / ``getAccel`` of
``BasisClasses::SphericalSL`` and ``BasisClasses::Cylindrical`` (reference
expui/BiorthBasis.cc:583-709, 711-958, 1851-1927, 4585-4757; bindings
pyEXP/BasisWrappers.cc:1308-1350, 1518, 1548-1596) with the per-particle ``accumulate`` loop and
the per-point evaluation replaced by the batched GPU kernels behind the C ABI.

Same names and argument meaning as the reference's Python API; coefficient containers mirror
``CoefClasses::SphStruct`` / ``CylStruct`` (expui/CoefStruct.H): ``coefs`` is complex,
``(lmax+1)(lmax+2)/2 x nmax`` (cos + i sin per (l,m), expui/BiorthBasis.cc:504-516) or
``(mmax+1) x nmax``.

Known differences from the reference twin, all below the 1e-10/1e-8 budgets: the radius
regulariser (1e-16 here, 1e-20/1e-18 there: expui/BiorthBasis.cc:589, 824-825), ``factorial``
through ``lgamma`` there.  Outside [rmin, rmax] of the table the N-body continuation of the
force path is used (src/SphericalBasis.cc:1555-1560); pyEXP extrapolates the table linearly.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .component import Component
from .cylinder import Cylinder
from .sphere import Sphere


@dataclass
class SphStruct:
    """CoefClasses::SphStruct (expui/CoefStruct.H)"""
    lmax: int
    nmax: int
    scale: float
    time: float
    coefs: np.ndarray                 # complex [(lmax+1)(lmax+2)/2, nmax]
    ctr: tuple = (0.0, 0.0, 0.0)
    normed: bool = True
    geometry: str = "sphere"


@dataclass
class CylStruct:
    """CoefClasses::CylStruct (expui/CoefStruct.H)"""
    mmax: int
    nmax: int
    time: float
    coefs: np.ndarray                 # complex [mmax+1, nmax]: cos + i sin
    ctr: tuple = (0.0, 0.0, 0.0)
    geometry: str = "cylinder"


FIELD_LABELS = ["dens m=0", "dens m>0", "dens", "potl m=0", "potl m>0", "potl",
                "rad force", "mer force", "azi force"]     # expui/BiorthBasis.cc:71-97 (spherical)


def _points(x, y=None, z=None):
    if y is None:
        p = np.ascontiguousarray(x, dtype=np.float64)
        return p.reshape(-1, 3)
    return np.ascontiguousarray(np.stack([np.ravel(x), np.ravel(y), np.ravel(z)], axis=1), dtype=np.float64)


def _prepare(mass, pos, center, rot, posvelrows):
    """initFromArray + the coordinate part of addFromArray (expui/BiorthBasis.cc:4585-4735):
    pp = rot (p - ctr).  Returns (mass, pos[n,3], centre to give the device)."""
    m = np.ascontiguousarray(mass, dtype=np.float64).ravel()
    p = np.asarray(pos, dtype=np.float64)
    if posvelrows:
        p = p.T
    if p.ndim != 2 or p.shape[1] < 3:
        raise RuntimeError("Basis::addFromArray: you must pass a position array with at least three "
                           f"columns for x, y, z.  Yours has {p.shape[-1] if p.ndim == 2 else 0}.")
    p = np.ascontiguousarray(p[:, :3])
    if len(m) != len(p):
        raise RuntimeError("Basis::addFromArray: mass and position arrays differ in length")
    ctr = np.asarray(center, dtype=np.float64)
    if rot is not None and not np.allclose(rot, np.eye(3), atol=0.0):
        p = (p - ctr) @ np.asarray(rot, dtype=np.float64).T
        ctr = np.zeros(3)
    return m, p, ctr


class _ComponentCache:
    """One device-resident particle store per basis object, reused across createFromArray /
    getFields / getAccel calls: a sweep over snapshots (BASELINE configs[4]) would otherwise pay a
    dozen cudaMalloc/cudaFree pairs of O(n) per call, which costs more than the kernels."""

    def _component(self, n):
        n = max(int(n), 1)
        c = getattr(self, "_comp", None)
        if c is None or c.capacity < n:
            if c is not None:
                c.close()
            c = Component(max(n, 1024), device=self.device)
            self._comp = c
        return c

    def close(self):
        c = getattr(self, "_comp", None)
        if c is not None:
            c.close()
            self._comp = None


class SphericalSL(_ComponentCache):
    """BasisClasses::SphericalSL with the SLGridSph tables given explicitly."""

    def __init__(self, table, *, scale=1.0, rmin=None, rmax=None, NO_L0=False, NO_L1=False,
                 EVEN_L=False, EVEN_M=False, M0_only=False, device=0):
        self.table = table
        self.lmax, self.nmax, self.scale = table.lmax, table.nmax, float(scale)
        self.flags = dict(NO_L0=NO_L0, NO_L1=NO_L1, EVEN_L=EVEN_L, EVEN_M=EVEN_M, M0_only=M0_only)
        self.device = device
        self._kw = dict(scale=scale, rmin=rmin, rmax=rmax, device=device, **self.flags)
        self._sph = Sphere(table, **self._kw)
        self._den = None
        self.expcoef = np.zeros(((self.lmax + 1) ** 2, self.nmax))
        self.coefctr = np.zeros(3)
        self.coefret = None
        self.used = 0
        self.totalMass = 0.0

    # ---- coefficients ---------------------------------------------------------------
    def createFromArray(self, mass, pos, time=0.0, center=(0.0, 0.0, 0.0), rot=None,
                        roundrobin=True, posvelrows=False):
        """BiorthBasis::createFromArray: coefficients of the particle set (mass[n], pos[n,3])."""
        m, p, ctr = _prepare(mass, pos, center, rot, posvelrows)
        c = self._component(len(p))
        c.upload(p, None, m)
        c.set_center(ctr)
        self.expcoef = self._sph.determine_coefficients(c)
        self.used = int(self._sph.used)
        self.coefctr = np.asarray(center, dtype=np.float64)
        self.coefret = self._load_coefs(time)
        return self.coefret

    def _load_coefs(self, time):          # Spherical::load_coefs, expui/BiorthBasis.cc:480-516
        L, N = self.lmax, self.nmax
        coefs = np.zeros(((L + 1) * (L + 2) // 2, N), dtype=np.complex128)
        L0 = 0
        for l in range(L + 1):
            for mm in range(l + 1):
                row = l * l + (2 * mm - 1 if mm else 0)
                coefs[L0] = self.expcoef[row] + (1j * self.expcoef[row + 1] if mm else 0.0)
                L0 += 1
        return SphStruct(L, N, self.scale, float(time), coefs, tuple(self.coefctr))

    def set_coefs(self, coef: SphStruct):  # Spherical::set_coefs, expui/BiorthBasis.cc:518-581
        L, N = self.lmax, self.nmax
        if coef.coefs.shape != ((L + 1) * (L + 2) // 2, N):
            raise RuntimeError(f"Spherical::set_coefs: the basis has (lmax, nmax)=({L}, {N}) and the "
                               f"coef structure has (rows, cols)={coef.coefs.shape}")
        L0 = 0
        for l in range(L + 1):
            for mm in range(l + 1):
                row = l * l + (2 * mm - 1 if mm else 0)
                self.expcoef[row] = coef.coefs[L0].real
                if mm:
                    self.expcoef[row + 1] = coef.coefs[L0].imag
                L0 += 1
        self.coefret = coef
        self.coefctr = np.asarray(coef.ctr if coef.ctr is not None else (0, 0, 0), dtype=np.float64)

    # ---- evaluation -------------------------------------------------------------------
    def _eval(self, sphere, coef, pts):
        c = self._component(len(pts))
        c.upload(pts, None, np.ones(len(pts)))
        c.set_center(self.coefctr)
        sphere.set_coefficients(coef)
        c.zero_pot_acc()
        sphere.get_acceleration_and_potential(c)
        out = c.download(pos=False, vel=False)
        return out["acc"], out["pot"]

    def getAccel(self, x, y=None, z=None):
        """Basis::getAccel: N x 3 accelerations at the points (expui/BiorthBasis.cc:818-926)."""
        return self._eval(self._sph, self.expcoef, _points(x, y, z))[0]

    def _density_sphere(self):
        if self._den is None:
            t = self.table
            # dens(l,n) = interp(ef) sqrt(ev) interp(d0) (exputil/SLGridMP2.cc:912-952): the same
            # kernels with ef*ev in place of ef and d0 in place of p0
            class _T:
                pass
            d = _T()
            for k in ("lmax", "nmax", "numr", "cmap", "rmin", "rmax", "rmap", "ev"):
                setattr(d, k, getattr(t, k))
            d.ef = np.ascontiguousarray(t.ef * t.ev[:, :, None])
            d.p0 = np.ascontiguousarray(t.d0)
            self._den = Sphere(d, **self._kw)
        return self._den

    def getFields(self, x, y, z):
        """Basis::getFields: [dens0, dens1, dens, pot0, pot1, pot, Fx, Fy, Fz] at the point(s)
        (Spherical::crt_eval, expui/BiorthBasis.cc:944-958).  Scalars give a list of 9 floats,
        arrays a (9, n) array."""
        scalar = np.ndim(x) == 0
        pts = _points(x, y, z)
        acc, pot = self._eval(self._sph, self.expcoef, pts)
        l0 = np.zeros_like(self.expcoef)
        l0[0] = self.expcoef[0]
        _, pot0 = self._eval(self._sph, l0, pts)
        den = self._density_sphere()
        densfac = self.scale * 0.25 / np.pi / self.scale ** 3     # our kernels divide by scale once
        _, d = self._eval(den, self.expcoef, pts)
        _, d0 = self._eval(den, l0, pts)
        out = np.stack([d0 * densfac, (d - d0) * densfac, d * densfac, pot0, pot - pot0, pot,
                        acc[:, 0], acc[:, 1], acc[:, 2]])
        return [float(v) for v in out[:, 0]] if scalar else out


class Cylindrical(_ComponentCache):
    """BasisClasses::Cylindrical with the EmpCylSL grids given explicitly (exp_b200.cyltable)."""

    def __init__(self, table, *, rcylmax=20.0, acyl=None, device=0, **flags):
        self.table = table
        self.mmax, self.nmax = table.mmax, table.norder
        self.device = device
        self._cyl = Cylinder(table, rcylmax=rcylmax, acyl=acyl, device=device, **flags)
        shape = (self.mmax + 1, self.nmax)
        self.cos, self.sin = np.zeros(shape), np.zeros(shape)
        self.cylmass = 0.0
        self.coefctr = np.zeros(3)
        self.coefret = None
        self.used = 0

    def createFromArray(self, mass, pos, time=0.0, center=(0.0, 0.0, 0.0), rot=None,
                        roundrobin=True, posvelrows=False):
        m, p, ctr = _prepare(mass, pos, center, rot, posvelrows)
        c = self._component(len(p))
        c.upload(p, None, m)
        c.set_center(ctr)
        self.cos, self.sin = self._cyl.determine_coefficients(c)
        self.used, self.cylmass = int(self._cyl.used), float(self._cyl.cylmass)
        self.coefctr = np.asarray(center, dtype=np.float64)
        self.coefret = CylStruct(self.mmax, self.nmax, float(time), self.cos + 1j * self.sin,
                                 tuple(self.coefctr))          # Cylindrical::load_coefs :1864-1895
        return self.coefret

    def set_coefs(self, coef: CylStruct, cylmass=None):
        if coef.coefs.shape != (self.mmax + 1, self.nmax):
            raise RuntimeError("Cylindrical::set_coefs: dimension mismatch")
        self.cos, self.sin = np.ascontiguousarray(coef.coefs.real), np.ascontiguousarray(coef.coefs.imag)
        if cylmass is not None:
            self.cylmass = float(cylmass)
        self.coefret = coef
        self.coefctr = np.asarray(coef.ctr if coef.ctr is not None else (0, 0, 0), dtype=np.float64)

    def getAccel(self, x, y=None, z=None):
        pts = _points(x, y, z)
        c = self._component(len(pts))
        c.upload(pts, None, np.ones(len(pts)))
        c.set_center(self.coefctr)
        self._cyl.set_coefficients(self.cos, self.sin, self.cylmass)
        c.zero_pot_acc()
        self._cyl.get_acceleration_and_potential(c)
        out = c.download(pos=False, vel=False)
        return out["acc"]
