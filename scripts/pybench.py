#!/usr/bin/env python
"""BASELINE configs[4] in miniature: pyEXP-style createFromArray sweep over synthetic snapshots
(sphereSL on an NFW halo, cylinder on an exponential disk), host arrays in, coefficient
structure out.  Reports seconds per snapshot and particles/s for each basis.
  python scripts/pybench.py [n] [snapshots]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from exp_b200 import cyltable, ics, pyexp, slgrid  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
nsnap = int(sys.argv[2]) if len(sys.argv) > 2 else 5
tab = slgrid.make_slgrid(slgrid.NFWModel(), 6, 18, numr=2000, rmin=1e-3, rmax=20.0)
sph = pyexp.SphericalSL(tab)
ctab = cyltable.synthetic_cyl_table(mmax=6, norder=18, numx=256, numy=128)
cyl = pyexp.Cylindrical(ctab)
pos, _, mass = ics.nfw_halo(n, seed=43)
rng = np.random.default_rng(44)
a, h = 0.01, 0.001
R = -a * np.log(rng.random(n) * rng.random(n))
dpos = np.stack([R * np.cos(ph := 2 * np.pi * rng.random(n)), R * np.sin(ph),
                 h * np.arctanh(2.0 * rng.random(n) - 1.0)], axis=1)
for name, basis, p in (("sphereSL", sph, pos), ("cylinder", cyl, dpos)):
    times = []
    for k in range(nsnap + 1):
        q = p * (1.0 + 1e-3 * np.sin(1000.0 + k))          # a different snapshot every time
        t0 = time.perf_counter()
        coefs = basis.createFromArray(mass, q, time=float(k))
        times.append(time.perf_counter() - t0)
    warm = times[1:]
    print(f"{name}: n={n} first {times[0]*1e3:.1f} ms, then {np.mean(warm)*1e3:.1f} ms/snapshot "
          f"(min {np.min(warm)*1e3:.1f}) = {n/np.mean(warm):.3e} particles/s, coefs {coefs.coefs.shape}")
