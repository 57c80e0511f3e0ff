"""The reference's own N-body test case (tests/Halo: SLGridSph.model + config.yml + check.py),
from the committed fixture tests/golden/halo_fixture.npz (made by tests/golden/make_halo_fixture.py)."""
import os

import numpy as np

from exp_b200 import ics, slgrid

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "halo_fixture.npz")
_CACHE = {}


def load():
    if "fx" not in _CACHE:
        z = np.load(_PATH)
        fx = {k: (z[k] if z[k].ndim else z[k].item()) for k in z.files}
        model = slgrid.TableModel.from_arrays(fx["r"], fx["rho"], fx["m"], fx["u"])
        fx["table"] = slgrid.make_slgrid(model, int(fx["Lmax"]), int(fx["nmax"]), numr=int(fx["numr"]),
                                         rmin=float(fx["rmin"]), rmax=float(fx["rmax"]), cmap=1,
                                         rmap=float(fx["rmapping"]), refine=4, richardson=False)
        fx["pos"], fx["vel"], fx["mass"] = ics.table_model_halo(fx["r"], fx["rho"], fx["m"], fx["u"],
                                                                int(fx["nbodies"]), seed=11)
        _CACHE["fx"] = fx
    return _CACHE["fx"]


def virial_2T_VC(pos, vel, acc, mass):
    """-2T/VC as OutLog forms it (src/OutLog.cc:493-494, 598): T = sum m v^2/2, VC = sum m x.a"""
    T = 0.5 * float((mass[:, None] * vel * vel).sum())
    VC = float((mass[:, None] * pos * acc).sum())
    return -2.0 * T / VC
