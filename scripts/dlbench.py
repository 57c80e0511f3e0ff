#!/usr/bin/env python
"""Developer micro-benchmark: Component.upload / download of 1e7 particles from and to pageable
numpy arrays (the pinned bounce-buffer path); download() allocates fresh arrays, so its time
includes their first-touch page faults."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from exp_b200 import ics
from exp_b200.component import Component
n = 10_000_000
pos, vel, mass = ics.nfw_halo(n, seed=1)
c = Component(n)
c.upload(pos, vel, mass)
for k in range(3):
    t0 = time.perf_counter(); out = c.download(); t1 = time.perf_counter()
    t2 = time.perf_counter(); c.upload(pos, vel, mass); t3 = time.perf_counter()
    print(f"download {1e3*(t1-t0):.1f} ms ({800e6/(t1-t0)/1e9:.1f} GB/s), upload {1e3*(t3-t2):.1f} ms ({560e6/(t3-t2)/1e9:.1f} GB/s)")
assert np.array_equal(out["pos"], pos) and np.array_equal(out["vel"], vel)
print("round trip exact")
