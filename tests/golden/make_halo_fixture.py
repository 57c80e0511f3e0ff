#!/usr/bin/env python
"""Writes tests/golden/halo_fixture.npz from the reference's own N-body test fixture
(/root/reference/tests/Halo): the 1000-row SLGridSph.model table (r, rho, M, U) and the run
parameters of tests/Halo/config.yml (Lmax=2 nmax=10 numr=4000 rmin=1e-4 rmax=1.95
rmapping=0.0667 multistep=4 dtime=0.002), plus the acceptance threshold of tests/Halo/check.py
((mean 2T/VC - 1)^2 <= 0.003).  Run in the build container (the GPU box has no /root/reference):
    python tests/golden/make_halo_fixture.py
"""
import os
import sys

import numpy as np
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from exp_b200 import slgrid  # noqa: E402

REF = "/root/reference/tests/Halo"
model = slgrid.TableModel(os.path.join(REF, "SLGridSph.model"))
cfg = yaml.safe_load(open(os.path.join(REF, "config.yml")))
g = cfg["Global"]
fp = cfg["Components"][0]["force"]["parameters"]
assert cfg["Components"][0]["force"]["id"] == "sphereSL"
np.savez_compressed(
    os.path.join(ROOT, "tests", "golden", "halo_fixture.npz"),
    r=model.r, rho=model.rho, m=model.m, u=model.u,
    Lmax=fp["Lmax"], nmax=fp["nmax"], numr=fp["numr"], rmin=fp["rmin"], rmax=fp["rmax"],
    rmapping=fp["rmapping"], multistep=g["multistep"], dtime=g["dtime"], nsteps=g["nsteps"],
    dynfracV=g["dynfracV"], dynfracA=g["dynfracA"], nbodies=10000,      # gensph -N 10000 (tests/CMakeLists.txt:62)
    check_tol=0.003)                                                     # tests/Halo/check.py:19
print("wrote halo_fixture.npz:", {k: fp[k] for k in ("Lmax", "nmax", "numr", "rmin", "rmax", "rmapping")}, g)
