"""The C++ host driver (exp_b200/host/: Component, PotAccel -> Sphere / Cylinder,
ComponentContainer, begin_run, do_step, adjust_multistep_level) over the C ABI, against the
same ordering driven with the CPU oracle (tests/ref_stepper.py): single level, multistep,
and a disk+halo pair with mutual interactions (BASELINE configs[3] in miniature)."""
import numpy as np
import pytest

from conftest import acc_rel_err, get_table, rel_rms
from exp_b200 import cyltable, ics
from exp_b200.exphost import Simulation
from ref_stepper import RefSim, RefSphereStepper
from test_emul_cyl import disk

pytestmark = pytest.mark.gpu


def _halo(n, seed=31):
    pos, vel, mass = ics.hernquist_halo(n, seed=seed)
    return pos + np.array([0.01, -0.02, 0.005]), vel, mass


@pytest.mark.parametrize("multistep,dtime,nsteps", [(0, 0.005, 3), (3, 0.04, 2)])
def test_sphere_do_step(multistep, dtime, nsteps):
    tab = get_table("hernquist", 4, 10)
    n = 30000
    pos, vel, mass = _halo(n)
    dyn = dict(dynfracV=0.02, dynfracA=0.06)
    ref = RefSphereStepper(tab, pos, vel, mass, dtime, multistep, **dyn)
    ref.initialize()
    sim = Simulation(dtime, multistep, **dyn)
    h = sim.add_component("halo", pos, vel, mass)
    sim.set_sphere(h, tab)
    sim.begin_run()
    for _ in range(nsteps):
        ref.step()
    sim.do_steps(nsteps)
    out = sim.particles(h, n)
    assert np.array_equal(out["level"], ref.S.level)
    if multistep:
        assert len(np.unique(ref.S.level)) >= 3
        st = sim.state()
        assert np.allclose(st["pos_check"], dtime, rtol=1e-12)      # CHK_STEP identity
        assert np.allclose(st["vel_check"], dtime, rtol=1e-12)
        assert sim.level_report(h)["counts"].sum() == n
    assert np.abs(out["pos"] - ref.S.pos).max() <= 1e-10 * np.abs(ref.S.pos).max()
    assert np.abs(out["vel"] - ref.S.vel).max() <= 1e-9 * np.abs(ref.S.vel).max()
    assert acc_rel_err(out["acc"], ref.S.acc).max() <= 1e-8
    assert rel_rms(sim.sphere_coefs(h), ref.expcoef) <= 1e-10
    assert abs(sim.state()["tnow"] - nsteps * dtime) < 1e-12


def _check_coef_output(sim, h, d, tnow):
    """dump_coefs / coef_struct of both forces against the restated formats"""
    import coef_format as CF
    ex = sim.sphere_coefs(h)
    lmax, nmax = int(round(np.sqrt(ex.shape[0]))) - 1, ex.shape[1]
    buf = sim.dump_coefs(h, 0)
    magic, node, off = CF.parse_header(buf)
    assert magic == 0xc0a57a2
    assert node["id"] == "SphereSL" and node["nmax"] == nmax and node["lmax"] == lmax and node["normed"] is True
    assert node["time"] == tnow and node["scale"] == 1.0
    assert np.array_equal(np.frombuffer(buf[off:]), CF.sphere_payload(ex, lmax, nmax))
    assert sim.dump_coefs(h, 1) == CF.sphere_old_stream(ex, "SphereSL", tnow, 1.0, lmax, nmax)
    cof = np.frombuffer(sim.dump_coefs(h, 2), dtype=np.complex128).reshape(-1, nmax)
    assert np.array_equal(cof, CF.sphere_struct(ex, lmax, nmax))
    cs, sn, _ = sim.cylinder_coefs(d)
    buf = sim.dump_coefs(d, 0)
    magic, node, off = CF.parse_header(buf)
    assert magic == 0xc0a57a3
    assert node == {"time": tnow, "mmax": cs.shape[0] - 1, "nmax": cs.shape[1]}
    assert np.array_equal(np.frombuffer(buf[off:]), CF.cylinder_payload(cs, sn))
    cof = np.frombuffer(sim.dump_coefs(d, 2), dtype=np.complex128).reshape(cs.shape)
    assert np.array_equal(cof.real, cs) and np.array_equal(cof.imag[1:], sn[1:]) and not cof.imag[0].any()
    # dump_coefs_h5: first call creates the file, the second appends a snapshot
    import os
    import tempfile
    from exp_b200 import h5lite
    with tempfile.TemporaryDirectory() as td:
        for idx, geom, fid in ((h, "sphere", "SphereSL"), (d, "cylinder", "Cylinder")):
            path = os.path.join(td, f"coefs.{geom}.h5")
            assert sim.dump_coefs_h5(idx, path, config="id: test\n") == 1
            assert sim.dump_coefs_h5(idx, path) == 2
            r = h5lite.read_coefs_h5(path)
            assert (r["geometry"], r["forceID"], r["config"]) == (geom, fid, "id: test\n")
            assert r["name"] == ("halo" if geom == "sphere" else "disk")
            assert np.array_equal(r["times"], [tnow, tnow]) and r["coefs"].shape[0] == 2
            assert np.array_equal(r["coefs"][1], sim.coef_struct(idx))
            assert np.array_equal(r["rotations"][0], np.eye(3))


@pytest.mark.parametrize("multistep", [0, 2])
def test_disk_halo_interactions(multistep):
    """two components, each force acting on both (the reference's interaction list)"""
    tabS = get_table("hernquist", 4, 10)
    tabC = cyltable.synthetic_cyl_table(mmax=4, norder=6, numx=48, numy=24)
    nh, nd = 20000, 12000
    ph, vh, mh = _halo(nh, seed=7)
    ph *= 0.05                                   # bring part of the halo onto the disk's grid
    pd, vd, md = disk(nd, seed=8)
    md *= 0.05
    dtime = 0.002
    dyn = dict(dynfracV=0.05, dynfracA=0.2, dynfracP=0.3)
    ref = RefSim(dtime, multistep, **dyn)
    rh = ref.add("sphere", tabS, ph, vh, mh)
    rd = ref.add("cylinder", tabC, pd, vd, md)
    ref.add_interaction(rh, rd)
    ref.add_interaction(rd, rh)
    ref.initialize()

    sim = Simulation(dtime, multistep, **dyn)
    h = sim.add_component("halo", ph, vh, mh)
    d = sim.add_component("disk", pd, vd, md)
    sim.set_sphere(h, tabS)
    sim.set_cylinder(d, tabC)
    sim.add_interaction(h, d)
    sim.add_interaction(d, h)
    sim.begin_run()

    nsteps = 2
    for _ in range(nsteps):
        ref.step()
    sim.do_steps(nsteps)
    for idx, S, n in ((h, ref.S[rh], nh), (d, ref.S[rd], nd)):
        out = sim.particles(idx, n)
        assert np.array_equal(out["level"], S.level)
        assert np.abs(out["pos"] - S.pos).max() <= 1e-10 * np.abs(S.pos).max()
        assert np.abs(out["vel"] - S.vel).max() <= 1e-9 * np.abs(S.vel).max()
        assert acc_rel_err(out["acc"], S.acc).max() <= 1e-8
        assert np.abs(out["pot"] - S.pot).max() <= 1e-8 * np.abs(S.pot).max()
        # centring pass (Component::fix_positions): mtot, com, cov, coa of each component
        got, lev = sim.com(idx)
        want = ref.com[{h: rh, d: rd}[idx]]
        assert abs(got[0] - want[0]) <= 1e-12 * want[0]
        assert np.abs(got[1:4] - want[1:4]).max() <= 1e-10 * np.abs(S.pos).max()
        assert np.abs(got[4:7] - want[4:7]).max() <= 1e-9 * np.abs(S.vel).max()
        assert np.abs(got[7:10] - want[7:10]).max() <= 1e-8 * np.abs(S.acc).max()
        assert abs(lev[:, 0].sum() - want[0]) <= 1e-12 * want[0]
    if multistep:
        lv = np.concatenate([ref.S[rh].level, ref.S[rd].level])
        assert len(np.unique(lv)) >= 2
        # sub-steps re-bin only the tail of their active levels (csrc: ensure_sorted): after the
        # whole-component sorts of sub-step 0 the later sub-steps take the tail path
        for idx in (h, d):
            full, tail, tail_items, full_items = (int(x) for x in sim.sort_counters(idx))
            assert tail > 0 and full > 0
            assert tail_items < tail * (full_items // full)
    assert rel_rms(sim.sphere_coefs(h), ref.F[rh].use[0]) <= 1e-10
    _check_coef_output(sim, h, d, nsteps * dtime)
    cs, sn, cm = sim.cylinder_coefs(d)
    assert rel_rms(cs, ref.F[rd].use[0]) <= 1e-10
    assert rel_rms(sn[1:], ref.F[rd].use[1][1:]) <= 1e-10
    assert abs(cm - ref.F[rd].cylmass) <= 1e-11 * cm
