"""Synthetic initial conditions for the benchmark / parity workloads
(BASELINE.md section 4).

Counter-based (Philox) streams keyed by (seed, chunk) so that any rank can
regenerate exactly its own contiguous slice [lo, hi) of the global particle
set without generating the rest: chunk k always holds global particles
[k*CHUNK, (k+1)*CHUNK).
"""
from __future__ import annotations

import math

import numpy as np

CHUNK = 1 << 20


def _chunks(lo: int, hi: int):
    k = lo // CHUNK
    while k * CHUNK < hi:
        a, b = max(lo, k * CHUNK), min(hi, (k + 1) * CHUNK)
        yield k, a - k * CHUNK, b - k * CHUNK
        k += 1


def _rng(seed: int, chunk: int) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(key=[seed, chunk]))


def _isotropic(rng, n):
    cost = rng.uniform(-1.0, 1.0, n)
    phi = rng.uniform(0.0, 2.0 * math.pi, n)
    sint = np.sqrt(1.0 - cost * cost)
    return np.stack([sint * np.cos(phi), sint * np.sin(phi), cost], axis=1)


def _assemble(gen_chunk, lo, hi, total_mass, ntot):
    pos = np.empty((hi - lo, 3))
    vel = np.empty((hi - lo, 3))
    jobs, o = [], 0
    for k, a, b in _chunks(lo, hi):
        jobs.append((k, a, b, o))
        o += b - a

    def run(job):
        k, a, b, o = job
        p, v = gen_chunk(k)
        pos[o:o + b - a] = p[a:b]
        vel[o:o + b - a] = v[a:b]

    if len(jobs) > 1:
        from concurrent.futures import ThreadPoolExecutor
        import os
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1, 16)) as ex:
            list(ex.map(run, jobs))
    else:
        for j in jobs:
            run(j)
    mass = np.full(hi - lo, total_mass / ntot)
    return pos, vel, mass


def hernquist_halo(ntot: int, lo: int = 0, hi: int | None = None, seed: int = 42,
                   a: float = 1.0, mass: float = 1.0, rmax: float = 50.0):
    """Hernquist sphere: r = a sqrt(u)/(1-sqrt(u)), u ~ U(0, u_max) (truncated at
    rmax), isotropic directions; isotropic Gaussian velocities with the
    Jeans dispersion of the untruncated model (only used by level assignment)."""
    hi = ntot if hi is None else hi
    umax = (rmax / (rmax + a)) ** 2

    def gen(k):
        rng = _rng(seed, k)
        s = np.sqrt(rng.uniform(0.0, umax, CHUNK))
        r = a * s / (1.0 - s)
        pos = r[:, None] * _isotropic(rng, CHUNK)
        x = r / a
        # Hernquist (1990) eq. 10 radial dispersion
        sig2 = (mass / (12.0 * a)) * (12.0 * x * (1 + x) ** 3 * np.log1p(1.0 / x)
                                      - x / (1 + x) * (25 + 52 * x + 42 * x * x + 12 * x ** 3))
        sig = np.sqrt(np.maximum(sig2, 1e-12))
        vel = rng.standard_normal((CHUNK, 3)) * sig[:, None]
        return pos, vel

    return _assemble(gen, lo, hi, mass, ntot)


def nfw_halo(ntot: int, lo: int = 0, hi: int | None = None, seed: int = 43,
             rs: float = 1.0, c: float = 10.0, mass: float = 1.0):
    """NFW sphere truncated at c*rs: inverse CDF of m(x)=ln(1+x)-x/(1+x) by
    table lookup + Newton; velocities Gaussian with sigma^2 ~ |Phi|/3 (level assignment only)."""
    hi = ntot if hi is None else hi
    mu = lambda x: np.log1p(x) - x / (1.0 + x)
    mc = mu(c)
    tab_x = np.concatenate([[0.0], np.geomspace(1e-6, c, 8192)])
    tab_m = mu(tab_x)

    def gen(k):
        rng = _rng(seed, k)
        u = rng.uniform(0.0, 1.0, CHUNK) * mc
        x = np.interp(u, tab_m, tab_x)              # table guess, then Newton on m(x) = u
        for _ in range(4):
            x = np.clip(x - (mu(x) - u) * (1.0 + x) ** 2 / np.maximum(x, 1e-300), 1e-12, c)
        r = rs * x
        pos = r[:, None] * _isotropic(rng, CHUNK)
        phi = mass / mc * np.log1p(r / rs) / np.maximum(r, 1e-12)
        vel = rng.standard_normal((CHUNK, 3)) * np.sqrt(phi / 3.0)[:, None]
        return pos, vel

    return _assemble(gen, lo, hi, mass, ntot)


def exponential_disk(ntot: int, lo: int = 0, hi: int | None = None, seed: int = 44,
                     a: float = 0.01, h: float = 0.001, mass: float = 0.05,
                     vcirc: float = 1.0):
    """Exponential disk: R = -a ln(u1 u2) (Gamma(2)), z = h atanh(2u-1) (sech^2),
    phi uniform; flat rotation curve + small dispersion."""
    hi = ntot if hi is None else hi

    def gen(k):
        rng = _rng(seed, k)
        u1 = rng.uniform(1e-300, 1.0, CHUNK)
        u2 = rng.uniform(1e-300, 1.0, CHUNK)
        R = -a * np.log(u1 * u2)
        z = h * np.arctanh(np.clip(2.0 * rng.uniform(0.0, 1.0, CHUNK) - 1.0, -1 + 1e-15, 1 - 1e-15))
        phi = rng.uniform(0.0, 2.0 * math.pi, CHUNK)
        pos = np.stack([R * np.cos(phi), R * np.sin(phi), z], axis=1)
        vel = np.stack([-vcirc * np.sin(phi), vcirc * np.cos(phi), np.zeros(CHUNK)], axis=1)
        vel += 0.05 * vcirc * rng.standard_normal((CHUNK, 3))
        return pos, vel

    return _assemble(gen, lo, hi, mass, ntot)


def table_model_halo(r, rho, m, u, n: int, seed: int = 11):
    """Equilibrium realisation of a tabulated spherical model (columns r, rho, M(<r), U(r) of a
    reference-format SLGridSph.model file; the job of the reference's `gensph`,
    utils/ICs): radii from the mass profile, isotropic velocities from the Eddington
    distribution function f(E) of (rho, U).  Equal masses m_tot / n."""
    r, rho, m, u = (np.asarray(a, dtype=np.float64) for a in (r, rho, m, u))
    rng = np.random.Generator(np.random.Philox(key=[seed, 0]))
    psi = -(u - u[-1])                       # relative potential, 0 at the truncation radius
    # rho(psi): psi decreases with r; work on increasing psi
    ps, rh = psi[::-1], rho[::-1]
    ps, keep = np.unique(ps, return_index=True)
    rh = rh[keep]
    drho = np.gradient(rh, ps)
    # G(e) = int_0^e rho'(psi) dpsi / sqrt(e - psi) = 2 int_0^sqrt(e) rho'(e - t^2) dt
    eg = np.linspace(0.0, ps[-1], 2000)
    tt = np.linspace(0.0, 1.0, 400)
    G = np.empty_like(eg)
    for k, e in enumerate(eg):
        t = tt * math.sqrt(e)
        G[k] = 2.0 * np.trapezoid(np.interp(e - t * t, ps, drho), t) if e > 0 else 0.0
    f = np.maximum(np.gradient(G, eg), 0.0)           # up to the constant 1/(sqrt(8) pi^2)
    # radii: inverse of M(r)
    rr = np.interp(rng.uniform(m[0], m[-1], n), m, r)
    pr = np.interp(rr, r, psi)
    # speeds: pdf(v) ~ v^2 f(psi - v^2/2) on [0, sqrt(2 psi)], inverted on a per-particle grid
    s = np.linspace(0.0, 1.0, 257)[None, :]
    v = s * np.sqrt(2.0 * pr)[:, None]
    pdf = v * v * np.interp(pr[:, None] - 0.5 * v * v, eg, f)
    cdf = np.cumsum(0.5 * (pdf[:, 1:] + pdf[:, :-1]), axis=1)
    cdf = np.concatenate([np.zeros((n, 1)), cdf], axis=1)
    tot = cdf[:, -1:]
    q = rng.uniform(0.0, 1.0, n)[:, None] * tot
    idx = np.clip((cdf < q).sum(axis=1), 1, 256)
    ar = np.arange(n)
    c0, c1 = cdf[ar, idx - 1], cdf[ar, idx]
    w = np.where(c1 > c0, (q[:, 0] - c0) / np.maximum(c1 - c0, 1e-300), 0.5)
    speed = v[ar, idx - 1] + w * (v[ar, idx] - v[ar, idx - 1])
    pos = rr[:, None] * _isotropic(rng, n)
    vel = speed[:, None] * _isotropic(rng, n)
    mass = np.full(n, m[-1] / n)
    return pos, vel, mass
