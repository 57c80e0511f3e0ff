#!/usr/bin/env python
"""Developer micro-benchmark: per-kernel times of the fused sphere step (profile counters),
plus a parity check of coefficients/forces against the CPU checker on a sub-sample.
  python scripts/kbench.py [n] [steps]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from exp_b200 import _lib, ics, slgrid  # noqa: E402
from exp_b200.component import Component  # noqa: E402
from exp_b200.sphere import Sphere  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
L = _lib.lib()
tab = slgrid.make_slgrid(slgrid.NFWModel(), 6, 18, numr=2000, rmin=1e-3, rmax=20.0)
pos, vel, mass = ics.nfw_halo(n, seed=43)
c = Component(n)
s = Sphere(tab)
c.upload(pos, vel, mass)
s.determine_coefficients(c, to_host=False)
s.set_coefficients(None)
c.zero_pot_acc()
s.get_acceleration_and_potential(c)
for _ in range(3):
    s.step(c, 1e-3)
c.sync()
_lib.check(L.expb200_profile_reset())
_lib.check(L.expb200_profile_enable(1))
c.event_record(0)
if os.environ.get("KBENCH_SINGLE"):
    for _ in range(steps):
        s.step(c, 1e-3)
else:
    s.steps(c, 1e-3, steps)
c.event_record(1)
c.sync()
ms = c.event_elapsed_ms(0, 1) / steps
_lib.check(L.expb200_profile_enable(0))
prof = _lib.profile_read()
print(f"variant={os.environ.get('EXPB200_DEPOSIT_VARIANT', '0')} n={n} ms/step={ms:.4f} rate={n / ms * 1e3:.4e}")
print("  " + " ".join(f"{k}={v[1] / steps:.4f}" for k, v in prof.items() if v[2]))
if os.environ.get("KBENCH_COV"):
    sampT = int(os.environ["KBENCH_COV"])
    s.covariance(c, sampT)                      # warm-up (allocations)
    for rep in range(3):
        c.event_record(2)
        s.covariance(c, sampT)
        c.event_record(3)
        c.sync()
        print(f"  sphere covariance sampT={sampT}: {c.event_elapsed_ms(2, 3):.3f} ms (incl. D2H of the result)")
