"""The reference's own N-body test (tests/Halo: `gensph -N 10000 -i SLGridSph.model`, then `exp
config.yml` with Lmax=2 nmax=10 numr=4000 multistep=4, then check.py: (mean 2T/VC - 1)^2 <= 0.003)
run from the committed fixture: on the CPU with the oracle-driven stepper, and (GPU) through the
C++ host driver over the C ABI, which must also track the oracle run step for step."""
import numpy as np
import pytest

import halo_fixture as H
from conftest import acc_rel_err, rel_rms
from ref_stepper import RefSphereStepper


def _dyn(fx):
    return dict(dynfracV=float(fx["dynfracV"]), dynfracA=float(fx["dynfracA"]))


def test_fixture_is_the_reference_case():
    fx = H.load()
    assert (fx["Lmax"], fx["nmax"], fx["numr"], fx["multistep"]) == (2, 10, 4000, 4)
    assert fx["rmin"] == 1e-4 and fx["rmax"] == 1.95 and fx["rmapping"] == 0.0667 and fx["dtime"] == 0.002
    assert len(fx["r"]) == 1000 and abs(fx["m"][-1] - 1.008) < 1e-12
    # biorthogonality of the table built from the reference's model (SLGridSph::orthoCheck,
    # exputil/SLGridMP2.cc:1775-1836; the reference's orthoTest tolerance is 1e-4)
    from exp_b200 import slgrid
    for l in range(3):
        assert np.abs(slgrid.ortho_check(fx["table"], l) - np.eye(10)).max() < 1e-4


def test_oracle_run_passes_reference_check():
    fx = H.load()
    ref = RefSphereStepper(fx["table"], fx["pos"], fx["vel"], fx["mass"], float(fx["dtime"]),
                           int(fx["multistep"]), **_dyn(fx))
    ref.initialize()
    vals = [H.virial_2T_VC(ref.S.pos, ref.S.vel, ref.S.acc, ref.S.mass)]
    for k in range(20):
        ref.step()
        if (k + 1) % 10 == 0:                       # outlog nint: 10
            vals.append(H.virial_2T_VC(ref.S.pos, ref.S.vel, ref.S.acc, ref.S.mass))
    mean = float(np.mean(vals))
    assert (mean - 1.0) ** 2 <= float(fx["check_tol"]), vals
    assert len(np.unique(ref.S.level)) >= 3


@pytest.mark.gpu
def test_host_driver_runs_reference_case():
    from exp_b200.exphost import Simulation
    fx = H.load()
    tab, n = fx["table"], int(fx["nbodies"])
    dtime, multistep = float(fx["dtime"]), int(fx["multistep"])
    ref = RefSphereStepper(tab, fx["pos"], fx["vel"], fx["mass"], dtime, multistep, **_dyn(fx))
    ref.initialize()
    sim = Simulation(dtime, multistep, **_dyn(fx))
    h = sim.add_component("halo", fx["pos"], fx["vel"], fx["mass"])
    sim.set_sphere(h, tab)
    sim.begin_run()
    out = sim.particles(h, n)
    vals = [H.virial_2T_VC(out["pos"], out["vel"], out["acc"], fx["mass"])]
    nsteps, ncheck = 120, 10
    for k in range(nsteps // 10):
        sim.do_steps(10)
        out = sim.particles(h, n)
        vals.append(H.virial_2T_VC(out["pos"], out["vel"], out["acc"], fx["mass"]))
        if (k + 1) * 10 == ncheck:                  # step-for-step agreement with the oracle run
            for _ in range(ncheck):
                ref.step()
            assert np.array_equal(out["level"], ref.S.level)
            assert np.abs(out["pos"] - ref.S.pos).max() <= 1e-10 * np.abs(ref.S.pos).max()
            assert np.abs(out["vel"] - ref.S.vel).max() <= 1e-9 * np.abs(ref.S.vel).max()
            assert acc_rel_err(out["acc"], ref.S.acc).max() <= 1e-8
            assert rel_rms(sim.sphere_coefs(h), ref.expcoef) <= 1e-10
    mean = float(np.mean(vals))
    assert (mean - 1.0) ** 2 <= float(fx["check_tol"]), vals          # tests/Halo/check.py:19
    assert abs(sim.state()["tnow"] - nsteps * dtime) < 1e-9
