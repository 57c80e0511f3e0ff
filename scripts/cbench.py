#!/usr/bin/env python
"""Developer micro-benchmark of the cylinder path (BASELINE configs[2] shape): coefficients,
force and, optionally, the covariance on an exponential disk, per-kernel times from the
library's profile counters.
  python scripts/cbench.py [n] [reps] [sampT]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from exp_b200 import _lib, cyltable  # noqa: E402
from exp_b200.component import Component  # noqa: E402
from exp_b200.cylinder import Cylinder  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
sampT = int(sys.argv[3]) if len(sys.argv) > 3 else 0
L = _lib.lib()
tab = cyltable.synthetic_cyl_table(mmax=6, norder=18, numx=256, numy=128)
rng = np.random.default_rng(44)
a, h = 0.01, 0.001
R = -a * np.log(rng.random(n) * rng.random(n))
z = h * np.arctanh(2.0 * rng.random(n) - 1.0)
ph = 2 * np.pi * rng.random(n)
pos = np.stack([R * np.cos(ph), R * np.sin(ph), z], axis=1)
vel = np.zeros((n, 3))
mass = np.full(n, 1.0 / n)
c = Component(n)
cy = Cylinder(tab)
c.upload(pos, vel, mass)
cy.determine_coefficients(c, to_host=False)          # warm-up (allocations)
cy.set_coefficients(None, None)
c.zero_pot_acc()
cy.get_acceleration_and_potential(c)
c.sync()
_lib.check(L.expb200_profile_reset())
_lib.check(L.expb200_profile_enable(1))
for _ in range(reps):
    c.upload(pos, vel, mass)                          # new positions: the binning is redone
    cy.determine_coefficients(c, to_host=False)
    cy.set_coefficients(None, None)
    c.zero_pot_acc()
    cy.get_acceleration_and_potential(c)
c.sync()
_lib.check(L.expb200_profile_enable(0))
prof = _lib.profile_read()
tot = sum(v[1] for k, v in prof.items() if v[2] and k not in ("pack",)) / reps
print(f"cylinder n={n} kernels ms/(bin+coef+force)={tot:.4f} rate={n / tot * 1e3:.4e}")
print("  " + " ".join(f"{k}={v[1] / reps:.4f}" for k, v in prof.items() if v[2]))
if sampT:
    cy.covariance(c, sampT)                      # warm-up: table build, scratch allocation
    for rep in range(3):
        c.event_record(2)
        cy.covariance(c, sampT)
        c.event_record(3)
        c.sync()
        print(f"  covariance sampT={sampT}: {c.event_elapsed_ms(2, 3):.3f} ms (incl. D2H of the result)")
