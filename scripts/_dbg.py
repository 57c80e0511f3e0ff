import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import numpy as np
from conftest import get_table
from exp_b200 import ics
from exp_b200.component import Component
from exp_b200.sphere import Sphere
tab4 = get_table("hernquist", 2, 8, numr=4000)
n = 20000
pos, vel, mass = ics.hernquist_halo(n, seed=78)
pos = pos + np.array([0.01, -0.02, 0.005])
def fresh():
    c = Component(n); s = Sphere(tab4)
    c.upload(pos, vel, mass); s.determine_coefficients(c); s.set_coefficients(None); c.zero_pot_acc(); s.get_acceleration_and_potential(c)
    return c, s
for nst in (1, 2, 3):
    c2, s2 = fresh(); c3, s3 = fresh()
    for _ in range(nst): s2.step(c2, 1e-3)
    s3.steps(c3, 1e-3, nst)
    a, b = c2.download(), c3.download()
    for k in ("pos", "vel", "acc", "pot"):
        d = a[k] != b[k]
        if d.ndim > 1: d = d.any(axis=1)
        idx = np.nonzero(d)[0]
        print(nst, k, "mismatch", len(idx), idx[:8], (np.abs(a[k]-b[k]).max()))
# run-to-run determinism of steps(3)
c4, s4 = fresh(); s4.steps(c4, 1e-3, 3); c5, s5 = fresh(); s5.steps(c5, 1e-3, 3)
a, b = c4.download(), c5.download()
print("steps3 twice identical:", all(a[k].tobytes() == b[k].tobytes() for k in a))
c4, s4 = fresh(); [s4.step(c4, 1e-3) for _ in range(3)]; c5, s5 = fresh(); [s5.step(c5, 1e-3) for _ in range(3)]
a, b = c4.download(), c5.download()
print("step x3 twice identical:", all(a[k].tobytes() == b[k].tobytes() for k in a))
