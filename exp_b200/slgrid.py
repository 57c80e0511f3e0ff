"""Sturm-Liouville radial basis tables for the sphereSL force (host side).

This is *table provenance*, not the hot path: it produces the arrays that the
reference's ``SLGridSph`` hands to the force loop (``table[l].ev[n]``,
``table[l].ef(n,i)``, ``p0[i]``, ``d0[i]`` on a uniform grid in the mapped
coordinate xi; reference exputil/SLGridMP2.cc:1355-1382 ``init_table`` and
:1103-1352 ``compute_table``).  The reference solves the SL problem with the
Fortran package SLEDGE, which cannot be built here (no gfortran), so this
module solves the *same* problem

    -(p u')' + q u = lambda w u,   p = r^2 f^2,  q = (l(l+1) f - rho r^2) f,
                                   w = -rho r^2 f         (rho = 4 pi rho_model)

(reference exputil/SLGridMP2.cc:3647-3653) with the same boundary conditions
(:1153-1164: l=0 inner  u'/u = f'/f, l>0 inner u=0, outer
u'/u = -(l+1)/r - f'/f) by a symmetric finite-volume discretisation on a
refinement of the xi grid plus one Richardson extrapolation, normalised so that
int w u_n u_m dr = delta_nm, which is the reference's orthoCheck identity
(exputil/SLGridMP2.cc:1775-1836).

In a real deployment the tables come from EXP itself through the C ABI
(``expb200_sph_create``); this generator only feeds tests and the benchmark.
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np
from scipy.linalg import eigh_tridiagonal


# ----------------------------------------------------------------------------
# Spherical background models: potential f(r), its derivative and density
# ----------------------------------------------------------------------------
class HernquistModel:
    """rho = M a / (2 pi r (r+a)^3), Phi = -M/(r+a)  (G=1)."""

    def __init__(self, mass: float = 1.0, a: float = 1.0):
        self.mass, self.a = mass, a

    def pot(self, r):
        return -self.mass / (r + self.a)

    def dpot(self, r):
        return self.mass / (r + self.a) ** 2

    def dens(self, r):
        return self.mass * self.a / (2.0 * math.pi * r * (r + self.a) ** 3)

    def mass_enclosed(self, r):
        return self.mass * r * r / (r + self.a) ** 2


class NFWModel:
    """rho = rho_s / (x (1+x)^2), x = r/rs; Phi = -4 pi rho_s rs^3 ln(1+x)/r.

    ``mass`` is the mass inside ``c*rs``."""

    def __init__(self, mass: float = 1.0, rs: float = 1.0, c: float = 10.0):
        self.rs, self.c = rs, c
        mu = math.log(1.0 + c) - c / (1.0 + c)
        self.rho_s = mass / (4.0 * math.pi * rs ** 3 * mu)
        self.k = 4.0 * math.pi * self.rho_s * rs ** 3

    def pot(self, r):
        return -self.k * np.log1p(r / self.rs) / r

    def dpot(self, r):
        x = r / self.rs
        return self.k * (np.log1p(x) / (r * r) - 1.0 / (r * self.rs * (1.0 + x)))

    def dens(self, r):
        x = r / self.rs
        return self.rho_s / (x * (1.0 + x) ** 2)

    def mass_enclosed(self, r):
        x = r / self.rs
        return self.k * (np.log1p(x) - x / (1.0 + x))


class TableModel:
    """Model read from a reference-format ``SLGridSph.model`` file
    (header lines starting with ``!``/``#``, a count, then rows r, rho, M, U;
    reference exputil ``SphericalModelTable``).  Log-log cubic-free linear
    interpolation is enough for table generation."""

    def __init__(self, path: str):
        rows = []
        with open(path) as f:
            lines = [ln for ln in f if ln.strip() and ln.lstrip()[0] not in "!#"]
        n = int(lines[0].split()[0])
        for ln in lines[1:1 + n]:
            rows.append([float(v) for v in ln.split()[:4]])
        a = np.asarray(rows)
        self.r, self.rho, self.m, self.u = a.T

    @classmethod
    def from_arrays(cls, r, rho, m, u):
        self = cls.__new__(cls)
        self.r, self.rho, self.m, self.u = (np.asarray(a, dtype=np.float64) for a in (r, rho, m, u))
        return self

    def pot(self, r):
        return np.interp(r, self.r, self.u)

    def dpot(self, r):
        return np.interp(r, self.r, self.m) / (np.asarray(r) ** 2)

    def dens(self, r):
        return np.exp(np.interp(np.log(r), np.log(self.r), np.log(self.rho)))

    def mass_enclosed(self, r):
        return np.interp(r, self.r, self.m)


# ----------------------------------------------------------------------------
# The table
# ----------------------------------------------------------------------------
@dataclasses.dataclass
class SLGridSphTable:
    """Plain-array image of the reference's SLGridSph state."""
    lmax: int
    nmax: int
    numr: int
    cmap: int
    rmin: float
    rmax: float
    rmap: float
    ev: np.ndarray      # [(lmax+1), nmax]
    ef: np.ndarray      # [(lmax+1), nmax, numr]
    p0: np.ndarray      # [numr]
    d0: np.ndarray      # [numr]
    xi: np.ndarray      # [numr]
    r: np.ndarray       # [numr]

    @property
    def xmin(self):
        return r_to_xi(self.rmin, self.cmap, self.rmap)

    @property
    def xmax(self):
        return r_to_xi(self.rmax, self.cmap, self.rmap)

    @property
    def dxi(self):
        return (self.xmax - self.xmin) / (self.numr - 1)


def r_to_xi(r, cmap, rmap):
    """reference exputil/SLGridMP2.cc:711-726"""
    if cmap == 1:
        return (r / rmap - 1.0) / (r / rmap + 1.0)
    if cmap == 2:
        return np.log(r)
    return r


def xi_to_r(xi, cmap, rmap):
    """reference exputil/SLGridMP2.cc:728-746"""
    if cmap == 1:
        return (1.0 + xi) / (1.0 - xi) * rmap
    if cmap == 2:
        return np.exp(xi)
    return xi


def d_xi_to_r(xi, cmap, rmap):
    """d xi / d r as a function of xi (reference exputil/SLGridMP2.cc:748-765)"""
    if cmap == 1:
        return 0.5 * (1.0 - xi) ** 2 / rmap
    if cmap == 2:
        return np.exp(-xi)
    return np.ones_like(xi)


def _solve_level(model, l, nmax, xmin, xmax, npts, cmap, rmap):
    """Finite-volume SL solve on npts uniform xi nodes; returns (ev, u[nmax,npts])
    with int w u^2 dr = 1 (trapezoid in xi)."""
    h = (xmax - xmin) / (npts - 1)
    x = xmin + h * np.arange(npts)
    xm = x[:-1] + 0.5 * h
    r, rm = xi_to_r(x, cmap, rmap), xi_to_r(xm, cmap, rmap)
    f, fm = model.pot(r), model.pot(rm)
    rho = 4.0 * math.pi * model.dens(r)
    dxdr, dxdr_m = d_xi_to_r(x, cmap, rmap), d_xi_to_r(xm, cmap, rmap)

    # in xi:  -(P u_xi)_xi + Q u = lambda W u,  P = p dxi/dr,  Q = q dr/dxi, W = w dr/dxi
    P = (rm * rm * fm * fm) * dxdr_m
    Q = ((l * (l + 1) * f - rho * r * r) * f) / dxdr
    W = (-rho * r * r * f) / dxdr

    vol = np.full(npts, h)
    vol[0] = vol[-1] = 0.5 * h
    diag = np.zeros(npts)
    diag[:-1] += P / h
    diag[1:] += P / h
    diag += vol * Q
    off = -P / h
    wdiag = vol * W

    # Robin boundary terms  (p u_r at the ends)
    pa, pb = r[0] ** 2 * f[0] ** 2, r[-1] ** 2 * f[-1] ** 2
    dirichlet_inner = l > 0
    if not dirichlet_inner:
        # u_r = (f'/f) u at r_min  ->  + p u_r(a) u(a) on the left side
        diag[0] += pa * (model.dpot(r[0]) / f[0])
    beta = (l + 1.0) / r[-1] + model.dpot(r[-1]) / f[-1]
    diag[-1] += pb * beta

    if dirichlet_inner:
        diag, off, wdiag = diag[1:], off[1:], wdiag[1:]

    s = 1.0 / np.sqrt(wdiag)
    d2 = diag * s * s
    o2 = off * s[:-1] * s[1:]
    ev, vec = eigh_tridiagonal(d2, o2, select="i", select_range=(0, nmax - 1))
    u = (vec * s[:, None]).T                # u = W^{-1/2} y,  sum(vol W u^2) = 1
    if dirichlet_inner:
        u = np.concatenate([np.zeros((nmax, 1)), u], axis=1)
    return ev, u


def make_slgrid(model, lmax, nmax, numr=2000, rmin=1e-3, rmax=50.0, cmap=1,
                rmap=None, refine=8, richardson=True, nevsign=None) -> SLGridSphTable:
    """Build the SLGridSph table image for ``model`` (see module docstring).

    ``refine``: fine-grid nodes per table cell for the eigen-solve."""
    if rmap is None:
        rmap = 0.067 * rmax             # reference src/Sphere.cc:33
    xmin, xmax = r_to_xi(rmin, cmap, rmap), r_to_xi(rmax, cmap, rmap)
    dxi = (xmax - xmin) / (numr - 1)
    xi = xmin + dxi * np.arange(numr)
    r = xi_to_r(xi, cmap, rmap)
    p0 = model.pot(r)
    d0 = 4.0 * math.pi * model.dens(r)

    ev = np.zeros((lmax + 1, nmax))
    ef = np.zeros((lmax + 1, nmax, numr))
    for l in range(lmax + 1):
        n1 = refine * (numr - 1) + 1
        e1, u1 = _solve_level(model, l, nmax, xmin, xmax, n1, cmap, rmap)
        u1 = u1[:, ::refine]
        if richardson:
            n2 = 2 * refine * (numr - 1) + 1
            e2, u2 = _solve_level(model, l, nmax, xmin, xmax, n2, cmap, rmap)
            u2 = u2[:, ::2 * refine]
            sgn = np.sign(np.sum(u1 * u2, axis=1))[:, None]
            u = (4.0 * u2 - sgn * u1) / 3.0
            e = (4.0 * e2 - e1) / 3.0
        else:
            u, e = u1, e1
        # sign convention: positive at the first few grid points
        # (reference exputil/SLGridMP2.cc:1327-1339 uses ef[nevsign-1])
        k = (nevsign if nevsign else 4)
        sg = np.where(u[:, min(k, numr - 1)] < 0.0, -1.0, 1.0)[:, None]
        ev[l], ef[l] = e, u * sg
    return SLGridSphTable(lmax, nmax, numr, cmap, rmin, rmax, rmap, ev, ef, p0, d0, xi, r)


def ortho_check(tab: SLGridSphTable, l: int) -> np.ndarray:
    """-int r^2 pot_n dens_m dr over the table grid (trapezoid in xi); should be
    the identity (reference SLGridSph::orthoCheck, exputil/SLGridMP2.cc:1775-1836)."""
    pot = tab.ef[l] / np.sqrt(tab.ev[l])[:, None] * tab.p0[None, :]
    den = tab.ef[l] * np.sqrt(tab.ev[l])[:, None] * tab.d0[None, :]
    wts = np.full(tab.numr, tab.dxi)
    wts[0] = wts[-1] = 0.5 * tab.dxi
    jac = tab.r ** 2 / d_xi_to_r(tab.xi, tab.cmap, tab.rmap) * wts
    return -(pot * jac[None, :]) @ den.T


def table_from_cache(path: str, model) -> SLGridSphTable:
    """The table image from a reference SLGridSph HDF5 cache (SLGridSph::ReadH5Cache,
    exputil/SLGridMP2.cc:490-617) plus the model for p0 / d0 (SLGridSph::init_table, :1355-1382:
    p0 = sphpot(r), d0 = sphdens(r) on the xi grid).  Read with exp_b200.h5lite (no HDF5 library)."""
    from . import h5lite
    c = h5lite.read_slgrid_cache(path)
    lmax, nmax, numr, cmap = int(c["lmax"]), int(c["nmax"]), int(c["numr"]), int(c["cmap"])
    rmin, rmax, rmap = float(c["rmin"]), float(c["rmax"]), float(c["rmapping"])
    xmin, xmax = r_to_xi(rmin, cmap, rmap), r_to_xi(rmax, cmap, rmap)
    xi = xmin + (xmax - xmin) / (numr - 1) * np.arange(numr)
    r = xi_to_r(xi, cmap, rmap)
    return SLGridSphTable(lmax, nmax, numr, cmap, rmin, rmax, rmap, c["ev"], c["ef"], model.pot(r),
                          4.0 * math.pi * model.dens(r), xi, r)
