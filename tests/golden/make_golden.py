"""Generates the golden fixtures of tests/golden/*.npz.

The reference ships no golden vectors for this path and cannot be built
here, so these vectors are produced by the CPU oracle (oracle/exp_oracle.cc) and serve
as a regression pin of the oracle itself plus a fixed input/output pair for the CUDA
path.  Each file holds the SLGridSph table image, the particles and the oracle's
coefficients / accelerations / potentials / levels.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from exp_b200 import ics, slgrid  # noqa: E402
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def make(name, model, gen, lmax, nmax, numr, rmax, n, seed):
    tab = slgrid.make_slgrid(model, lmax, nmax, numr=numr, rmin=1e-3, rmax=rmax)
    pos, vel, mass = gen(n, seed=seed)
    pos = pos + np.array([0.01, -0.02, 0.005])
    pos[:3] *= 100.0
    pos[3:5] *= 1e-7
    g = O.Grid(tab)
    S = O.ParticleSet(pos, vel, mass)
    P = O.sph_params(rmin=tab.rmin, rmax=tab.rmax)
    coef, used = O.sph_coefficients(g, P, S)
    O.sph_accel(g, P, S, coef)
    acc, pot = S.acc.copy(), S.pot.copy()
    LP = O.level_params(dtime=0.01, multistep=4)
    O.adjust_levels(S, LP)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"), lmax=lmax, nmax=nmax, numr=numr, cmap=tab.cmap,
        rmin=tab.rmin, rmax=tab.rmax, rmap=tab.rmap, ev=tab.ev, ef=tab.ef.astype(np.float64),
        p0=tab.p0, d0=tab.d0, pos=pos, vel=vel, mass=mass, coef=coef, used=used, acc=acc, pot=pot,
        level=S.level.copy(), dtreq=S.dtreq.copy())
    print(name, "coef[0,:3]", coef[0, :3], "used", used)


if __name__ == "__main__":
    make("sph_A_hernquist_l4n10", slgrid.HernquistModel(), ics.hernquist_halo, 4, 10, 256, 50.0, 1500, 101)
    make("sph_B_nfw_l6n18", slgrid.NFWModel(), ics.nfw_halo, 6, 18, 128, 20.0, 1500, 102)
