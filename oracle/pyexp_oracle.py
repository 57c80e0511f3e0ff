"""TEST INFRASTRUCTURE: CPU restatement of the pyEXP twin of the path,
``BasisClasses::Spherical::accumulate`` / ``load_coefs`` / ``sph_eval`` / ``cyl_eval`` /
``crt_eval`` / ``computeAccel`` (reference expui/BiorthBasis.cc:583-665, 480-516, 711-816,
930-958, 818-926), in plain Python loops over the oracle's C table functions.  Small cases only.
PARITY PINNED: pyEXP itself cannot be built here, but the reference's own Spherical::accumulate /
sph_eval / cyl_eval / crt_eval / computeAccel and BiorthBasis::legendre_R are cut from
expui/BiorthBasis.cc and compiled in place (oracle/ref_extract.py, oracle/ref_path_wrap.cc), and
tests/test_oracle_ref.py::test_pyexp_twin_against_reference holds this file to them (<= 1e-13)."""
import ctypes as C
import math

import numpy as np

from . import oracle as O


def _legendre(lmax, x, deriv=False):
    p = np.zeros((lmax + 1, lmax + 1))
    dp = np.zeros((lmax + 1, lmax + 1))
    if deriv:
        O.lib().orc_dlegendre_R(lmax, float(x), O._dp(p), O._dp(dp))
    else:
        O.lib().orc_legendre_R(lmax, float(x), O._dp(p))
    return p, dp


def _factorial(l, m):
    # expui/BiorthBasis.cc:321-329: sqrt((2l+1)/(4 pi)) exp(0.5 (lgamma(1+l-m) - lgamma(1+l+m))), sqrt(2) for m>0
    f = math.sqrt((2.0 * l + 1.0) / (4.0 * math.pi)) * math.exp(0.5 * (math.lgamma(1.0 + l - m) - math.lgamma(1.0 + l + m)))
    return f * (math.sqrt(2.0) if m else 1.0)


def _dens(grid, r):
    m = np.zeros((grid.lmax + 1, grid.nmax))
    O.lib().orc_slgrid_get_dens(C.byref(grid.c), float(r), O._dp(m))
    return m


def accumulate(grid, mass, pos, scale=1.0, rmin=0.0, rmax=1e300):
    """Spherical::accumulate over all particles -> expcoef [(lmax+1)^2, nmax], used"""
    L, N = grid.lmax, grid.nmax
    expcoef = np.zeros(((L + 1) ** 2, N))
    norm = -4.0 * math.pi
    used = 0
    for (x, y, z), ms in zip(pos, mass):
        r = math.sqrt(x * x + y * y + z * z) + 1.0e-20
        if r < rmin or r > rmax:
            continue
        used += 1
        costh, phi = z / r, math.atan2(y, x)
        potd = grid.get_pot(r / scale)
        legs, _ = _legendre(L, costh)
        loffset = 0
        for l in range(L + 1):
            moffset = 0
            for m in range(l + 1):
                fac = _factorial(l, m) * legs[l, m]
                if m == 0:
                    expcoef[loffset + moffset] += potd[l] * fac * norm * ms
                    moffset += 1
                else:
                    expcoef[loffset + moffset] += fac * math.cos(phi * m) * potd[l] * norm * ms
                    expcoef[loffset + moffset + 1] += fac * math.sin(phi * m) * potd[l] * norm * ms
                    moffset += 2
            loffset += 2 * l + 1
    return expcoef, used


def load_coefs(expcoef, lmax):
    N = expcoef.shape[1]
    coefs = np.zeros(((lmax + 1) * (lmax + 2) // 2, N), dtype=np.complex128)
    L0 = L1 = 0
    for l in range(lmax + 1):
        for m in range(l + 1):
            if m == 0:
                coefs[L0] = expcoef[L1]
                L1 += 1
            else:
                coefs[L0] = expcoef[L1] + 1j * expcoef[L1 + 1]
                L1 += 2
            L0 += 1
    return coefs


def sph_eval(grid, expcoef, r, costh, phi, scale=1.0, NO_L0=False, NO_L1=False, EVEN_L=False,
             EVEN_M=False, M0_only=False):
    L = grid.lmax
    dend, potd, dpot = _dens(grid, r / scale), grid.get_pot(r / scale), grid.get_force(r / scale)
    legs, dlegs = _legendre(L, costh, deriv=True)
    sinth = math.sqrt(abs(1.0 - costh * costh))
    fac1 = _factorial(0, 0)
    den0 = pot0 = potr = 0.0
    if not NO_L0:
        den0 = fac1 * expcoef[0].dot(dend[0])
        pot0 = fac1 * expcoef[0].dot(potd[0])
        potr = fac1 * expcoef[0].dot(dpot[0])
    den1 = pot1 = pott = potp = 0.0
    loffset = 1
    for l in range(1, L + 1):
        skip = (EVEN_L and l % 2) or (NO_L1 and l == 1)
        moffset = 0
        for m in range(l + 1):
            if skip or (M0_only and m) or (EVEN_M and m % 2):
                continue                 # NB: moffset does not advance (as in the reference)
            fac1 = _factorial(l, m)
            if m == 0:
                row = expcoef[loffset + moffset]
                den1 += fac1 * legs[l, m] * row.dot(dend[l])
                pot1 += fac1 * legs[l, m] * row.dot(potd[l])
                potr += fac1 * legs[l, m] * row.dot(dpot[l])
                pott += fac1 * dlegs[l, m] * row.dot(potd[l])
                moffset += 1
            else:
                c, s = math.cos(phi * m), math.sin(phi * m)
                r0, r1 = expcoef[loffset + moffset], expcoef[loffset + moffset + 1]
                R0, R1 = r0.dot(dend[l]), r1.dot(dend[l])
                P0, P1 = r0.dot(potd[l]), r1.dot(potd[l])
                D0, D1 = r0.dot(dpot[l]), r1.dot(dpot[l])
                den1 += fac1 * legs[l, m] * (R0 * c + R1 * s)
                pot1 += fac1 * legs[l, m] * (P0 * c + P1 * s)
                potr += fac1 * legs[l, m] * (D0 * c + D1 * s)
                pott += fac1 * dlegs[l, m] * (P0 * c + P1 * s)
                potp += fac1 * legs[l, m] * (-P0 * s + P1 * c) * m
                moffset += 2
        loffset += 2 * l + 1
    densfac = 1.0 / scale ** 3 * 0.25 / math.pi
    potlfac = 1.0 / scale
    return [den0 * densfac, den1 * densfac, (den0 + den1) * densfac, pot0 * potlfac, pot1 * potlfac,
            (pot0 + pot1) * potlfac, potr * (-potlfac) / scale, pott * (-potlfac) / r,
            potp * (-potlfac) / (r * sinth)]


def crt_eval(grid, expcoef, x, y, z, **kw):
    """Spherical::crt_eval -> cyl_eval -> sph_eval: the 9 values of getFields"""
    R = math.sqrt(x * x + y * y) + 1.0e-18
    phi = math.atan2(y, x)
    r = math.sqrt(R * R + z * z) + 1.0e-18
    costh, sinth = z / r, R / r
    v = sph_eval(grid, expcoef, r, costh, phi, **kw)
    potR = v[6] * sinth - v[7] * costh * R / r
    potz = v[6] * costh + v[7] * sinth * R / r
    tpotx = potR * x / R - v[8] * y / R
    tpoty = potR * y / R + v[8] * x / R
    return [v[0], v[1], v[2], v[3], v[4], v[5], tpotx, tpoty, potz]
