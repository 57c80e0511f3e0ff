"""EmpCylSL table images for the cylinder force (host side; test/bench provenance).

In EXP the six grids potC/potS/rforceC/rforceS/zforceC/zforceS come out of the EOF basis
construction (exputil/EmpCylSL.cc:2189-3870), which is one-time setup and out of scope
(SURVEY.md 2d, 8f rank 3).  The hot path only interpolates them, so tests and the benchmark
use smooth synthetic tables on the same mapped (x, y) grid (exputil/EmpCylSL.cc:2123-2137).
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np


@dataclasses.dataclass
class CylTable:
    mmax: int
    norder: int
    numx: int
    numy: int
    cmapr: int
    cmapz: int
    ascale: float
    hscale: float
    rmin: float          # EmpCylSL::RMIN (units of ascale)
    rmax: float          # EmpCylSL::RMAX (units of ascale); Rtable = rmax / sqrt(2)
    potC: np.ndarray     # [(mmax+1), norder, numx+1, numy+1]
    potS: np.ndarray
    rforceC: np.ndarray
    rforceS: np.ndarray
    zforceC: np.ndarray
    zforceS: np.ndarray

    @property
    def rtable(self):
        return math.sqrt(0.5) * self.rmax


def r_to_xi(r, cmapr, ascale):
    return (r / ascale - 1.0) / (r / ascale + 1.0) if cmapr > 0 else r


def xi_to_r(xi, cmapr, ascale):
    return (1.0 + xi) / (1.0 - xi) * ascale if cmapr > 0 else xi


def z_to_y(z, cmapz, hscale):
    if cmapz == 1:
        return np.sign(z) * np.arcsinh(np.abs(z / hscale))
    if cmapz == 2:
        return z / np.sqrt(z * z + hscale * hscale)
    return z


def y_to_z(y, cmapz, hscale):
    if cmapz == 1:
        return hscale * np.sinh(y)
    if cmapz == 2:
        return y * hscale / np.sqrt(1.0 - y * y)
    return y


def synthetic_cyl_table(mmax=6, norder=18, numx=64, numy=32, ascale=0.01, hscale=0.001,
                        rmin=0.001, rmax=20.0, cmapr=1, cmapz=1) -> CylTable:
    rtable = math.sqrt(0.5) * rmax
    xmin, xmax = r_to_xi(rmin * ascale, cmapr, ascale), r_to_xi(rtable * ascale, cmapr, ascale)
    ymin, ymax = z_to_y(-rtable * ascale, cmapz, hscale), z_to_y(rtable * ascale, cmapz, hscale)
    x = xmin + (xmax - xmin) / numx * np.arange(numx + 1)
    y = ymin + (ymax - ymin) / numy * np.arange(numy + 1)
    R = xi_to_r(x, cmapr, ascale)[:, None]
    Z = y_to_z(y, cmapz, hscale)[None, :]

    def f(R, Z, m, n, phase):
        u, v = R / ascale, Z / (3.0 * hscale)
        rad = u ** m / (1.0 + u) ** (m + 1)
        osc = np.cos((n + 1) * 2.0 * u / (1.0 + u) + phase + 0.3 * m)
        return -rad * osc / (1.0 + v * v) * (1.0 + 0.1 * np.tanh(v) * ((n + m) % 2))

    shape = (mmax + 1, norder, numx + 1, numy + 1)
    out = {k: np.zeros(shape) for k in ("potC", "potS", "rforceC", "rforceS", "zforceC", "zforceS")}
    hR, hZ = 1e-6 * ascale, 1e-6 * hscale
    for m in range(mmax + 1):
        for n in range(norder):
            for tag, phase in (("C", 0.0), ("S", 0.9)):
                if tag == "S" and m == 0:
                    continue
                out["pot" + tag][m, n] = f(R, Z, m, n, phase)
                out["rforce" + tag][m, n] = -(f(R + hR, Z, m, n, phase) - f(np.maximum(R - hR, 0.0), Z, m, n, phase)) / (2 * hR)
                out["zforce" + tag][m, n] = -(f(R, Z + hZ, m, n, phase) - f(R, Z - hZ, m, n, phase)) / (2 * hZ)
    return CylTable(mmax, norder, numx, numy, cmapr, cmapz, ascale, hscale, rmin, rmax, **out)


def table_from_cache(path: str) -> CylTable:
    """The six EmpCylSL grids from a reference cache file (EmpCylSL::ReadH5Cache,
    exputil/EmpCylSL.cc:7486-7680), read with exp_b200.h5lite."""
    from . import h5lite
    c = h5lite.read_cylinder_cache(path)
    return CylTable(int(c["mmax"]), int(c["nmax"]), int(c["numx"]), int(c["numy"]), int(c["cmapr"]),
                    int(c["cmapz"]), float(c["ascl"]), float(c["hscl"]), float(c["rmin"]), float(c["rmax"]),
                    c["potC"], c["potS"], c["rforceC"], c["rforceS"], c["zforceC"], c["zforceS"])
