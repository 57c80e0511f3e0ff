"""bench_workloads.py -- the other BASELINE.json configurations as `bench.py --workload ...`:

  cylinder  configs[2]: Cylinder (EmpCylSL) exponential disk, 1e7 particles, mmax=6 norder=18,
            257x129 tables, covariance accumulation (sampT bins) every step
  diskhalo  configs[3]: disk + halo (Cylinder + sphereSL) with mutual interactions, multistep
            levels, through the C++ host driver (exp_b200/host: do_step, adjust_multistep_level),
            sharded over the GPUs with the coefficient all-reduces
  pyexp     configs[4]: basis.createFromArray coefficient sweep over synthetic snapshots
            (sphereSL + cylinder), host numpy arrays in, coefficient structures out

Same JSON line contract as bench.py (one line on rank 0); the default driver run only uses
the sphere workload, these lines are kept under profiles/."""
from __future__ import annotations

import ctypes as C
import os
import sys
import time

import numpy as np


def run(args, B):
    if args.impl == "reference":
        return {"cylinder": ref_cylinder, "diskhalo": ref_diskhalo, "pyexp": ref_pyexp}[args.workload](args, B)
    return {"cylinder": run_cylinder, "diskhalo": run_diskhalo, "pyexp": run_pyexp}[args.workload](args, B)


def cyl_table():
    from exp_b200 import cyltable
    return cyltable.synthetic_cyl_table(mmax=6, norder=18, numx=256, numy=128)


def disk_particles(ntot, lo, hi):
    from exp_b200 import ics
    return ics.exponential_disk(ntot, lo=lo, hi=hi, seed=44)


# Algorithmic HBM bytes per particle of the cylinder kernels: key (pos 32 R, key 4 W), sort
# (2 passes x (key+perm 8 R + 8 W)), deposit (perm 4 + gathered pos 32), force (perm 4 + pos 32 +
# acc 32 R + 32 W); the force also gathers 4 nodes x 42 table values (1344 B) from L2.
CYL_BYTES = {"key": 36, "sort": 32, "cyl_deposit": 36, "cyl_force": 100}
CYL_L2_GATHER = {"cyl_force": 4 * 42 * 8, "cyl_deposit": 0}


def _cyl_config(ntot, n, world, sampT):
    return {"workload": f"Cylinder (EmpCylSL) exponential disk, {ntot} particles total, mmax=6 norder=18, "
                        f"257x129 tables (synthetic smooth tables), covariance accumulation every step "
                        f"(sampT={sampT}) (BASELINE configs[2]); {n} per GPU on {world} GPU(s)",
            "particles_total": ntot, "particles_per_gpu": n, "mmax": 6, "norder": 18, "numx": 256, "numy": 128,
            "sampT": sampT, "dt": B_DT,
            "cache": "inputs larger than L2 (112 B/particle resident state + 186 MB of tables)"}


B_DT = 1.0e-3


def run_cylinder(args, B):
    R = B.Ranks()
    from exp_b200 import _lib
    from exp_b200.component import Component
    from exp_b200.cylinder import Cylinder
    L = _lib.lib()
    R.init()
    ntot = args.particles_per_gpu * R.world if args.particles_per_gpu else (
        args.particles_total if args.particles_total != 100_000_000 else 10_000_000)
    lo, hi = B.slice_of(ntot, R.rank, R.world)
    n = hi - lo
    tab = cyl_table()
    pos, vel, mass = disk_particles(ntot, lo, hi)
    hpos, hvel, hmass = B.pinned_empty((n, 3)), B.pinned_empty((n, 3)), B.pinned_empty((n,))
    hpos[:], hvel[:], hmass[:] = pos, vel, mass
    hout = dict(pos=B.pinned_empty((n, 3)), vel=B.pinned_empty((n, 3)), acc=B.pinned_empty((n, 3)),
                pot=B.pinned_empty((n,)))
    c = Component(n, device=R.local)
    cyl = Cylinder(tab, device=R.local)
    c.upload(hpos, hvel, hmass)
    sampT = args.sampT

    def allreduce():
        if R.comm:
            p, cnt, st = cyl.coef_device()
            _lib.check(L.expb200_allreduce_sum(R.comm, p, cnt, st))

    def step(cov=True):
        c.incr_velocity(0.5 * B_DT)
        c.incr_position(B_DT)
        cyl.determine_coefficients(c, to_host=False)
        allreduce()
        if cov:
            cyl.covariance(c, sampT, indx_base=lo)
        cyl.set_coefficients(None, None)
        c.zero_pot_acc()
        cyl.get_acceleration_and_potential(c)
        c.incr_velocity(0.5 * B_DT)

    cyl.determine_coefficients(c, to_host=False)
    allreduce()
    cyl.set_coefficients(None, None)
    c.zero_pot_acc()
    cyl.get_acceleration_and_potential(c)
    clocks = B.ClockSampler(R.local) if R.rank == 0 else None
    if clocks:
        clocks.start()
    for _ in range(args.warmup):
        step()
    c.sync()
    _lib.check(L.expb200_profile_reset())
    _lib.check(L.expb200_profile_enable(1))
    R.barrier()
    c.sync()
    if clocks:
        clocks.mark_begin()
    c.event_record(0)
    for _ in range(args.steps):
        step()
    c.event_record(1)
    c.sync()
    if clocks:
        clocks.mark_end()
    R.barrier()
    ms = R.max(c.event_elapsed_ms(0, 1))
    clk = clocks.stop() if clocks else None
    _lib.check(L.expb200_profile_enable(0))
    prof = _lib.profile_read()
    launches = sum(v[0] for v in prof.values())
    # the same step without the covariance, for the split
    c.event_record(2)
    for _ in range(args.steps):
        step(cov=False)
    c.event_record(3)
    c.sync()
    ms_nocov = R.max(c.event_elapsed_ms(2, 3))

    def e2e_step():
        c.upload_range(0, hpos, hvel, hmass)
        step()
        c.download_range(0, n, hout["pos"], hout["vel"], hout["acc"], hout["pot"])

    e2e_step()
    c.sync()
    R.barrier()
    c.event_record(4)
    for _ in range(args.e2e_steps):
        e2e_step()
    c.event_record(5)
    c.sync()
    ms_e2e = R.max(c.event_elapsed_ms(4, 5))
    h2d, d2h = int(R.sum(n * 56)), int(R.sum(n * 80 + sampT * 7 * (18 + 324) * 16))
    if R.rank != 0:
        R.close()
        return
    per = {k: v[1] / args.steps for k, v in prof.items() if v[2]}
    dom = max((k for k in per if k in CYL_BYTES or k.startswith("cyl")), key=lambda k: per[k])
    bytes_dom = CYL_BYTES.get(dom)
    roof = {"bound": "hbm", "kernel": dom, "avg_launch_ms": per[dom],
            "kernel_ms_per_step": {k: round(v, 4) for k, v in per.items()},
            "peak": B.HBM_PEAK_GBS, "unit": "GB/s", "peak_source": B.HBM_PEAK_SRC, "traffic": None,
            "step_ms_without_covariance": ms_nocov / args.steps,
            "covariance_ms_per_step": (ms - ms_nocov) / args.steps,
            "covariance_flop_per_particle": 8 * 18 * 18 * 7,
            "covariance_tflops": 8 * 18 * 18 * 7 * n / max((ms - ms_nocov) / args.steps * 1e-3, 1e-12) / 1e12}
    if bytes_dom:
        ach = bytes_dom * n / (per[dom] * 1e-3) / 1e9
        roof.update(achieved=ach, frac=ach / B.HBM_PEAK_GBS, algorithmic_bytes_per_particle=bytes_dom,
                    l2_gather_bytes_per_particle=CYL_L2_GATHER.get(dom, 0),
                    l2_gather_gbs=CYL_L2_GATHER.get(dom, 0) * n / (per[dom] * 1e-3) / 1e9)
    else:
        roof.update(achieved=None, frac=None)
    cpu = None
    if R.world == 1 and not args.no_cpu:
        cpu = _cpu_cylinder(tab, np.array(hpos[:min(n, 500_000)]), np.array(hvel[:min(n, 500_000)]),
                            np.array(hmass[:min(n, 500_000)]), sampT)
    line = {"metric": "particle-steps/sec (coef+covariance+force+kick-drift) cylinder", "value": ntot * args.steps / (ms * 1e-3),
            "unit": B.UNIT, "n_gpus": R.world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak" if args.particles_per_gpu else "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": dict(_cyl_config(ntot, n, R.world, sampT), allreduce=R.allreduce_desc()),
            "clocks": clk,
            "e2e": {"value": ntot * args.e2e_steps / (ms_e2e * 1e-3), "unit": B.UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.e2e_steps, "steps": args.e2e_steps,
                    "path": "Component.upload_range -> incr_velocity/position, Cylinder.determine_coefficients, "
                            "covariance (results to the host), get_acceleration_and_potential -> download_range"},
            "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu}
    B.emit(line)
    R.close()


def _cpu_cylinder(tab, pos, vel, mass, sampT, steps=1):
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    g = O.CylGridPy(tab)
    P = O.cyl_params(rcylmax=20.0, acyl=tab.ascale)
    S = O.ParticleSet(pos, vel, mass)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.incr_velocity(S, 0.5 * B_DT)
        O.incr_position(S, B_DT)
        cs, sn, used, cm = O.cyl_coefficients(g, P, S, nthreads=cores)
        m = min(len(pos), 50_000)            # the covariance runs under a mutex in the reference: one thread
        S2 = O.ParticleSet(pos[:m], vel[:m], mass[:m])
        tc = time.perf_counter()
        O.cyl_coefficients_cov(g, P, S2, sampT)
        tcov = (time.perf_counter() - tc) * len(pos) / m
        S.acc[:] = 0.0
        S.pot[:] = 0.0
        O.cyl_accel(g, P, S, cs, sn, cm, nthreads=cores)
        O.incr_velocity(S, 0.5 * B_DT)
    dt = time.perf_counter() - t0 - tcov * m / len(pos) + tcov
    return {"value": len(pos) * steps / dt, "unit": "particle-steps/s", "cores": cores, "kind": "port",
            "sample": f"first {len(pos)} disk particles, {steps} step(s), {cores} threads for coefficients and force; "
                      f"covariance single-threaded as in the reference (global mutex), timed on {m} particles and "
                      f"scaled (oracle/cyl_oracle.cc)"}


def ref_cylinder(args, B):
    R = B.Ranks()
    if R.rank != 0:
        return
    ntot = args.particles_total if args.particles_total != 100_000_000 else 10_000_000
    n = min(500_000, ntot)
    tab = cyl_table()
    pos, vel, mass = disk_particles(ntot, 0, n)
    cpu = _cpu_cylinder(tab, pos, vel, mass, args.sampT, steps=max(1, min(args.steps, 3)))
    cfg = _cyl_config(ntot, ntot // max(args.gpus, 1), max(args.gpus, 1), args.sampT)
    cfg["reference_sample_particles_per_step"] = n
    cfg["extrapolated"] = True
    B.emit({"impl": "reference", "metric": "particle-steps/sec (coef+covariance+force+kick-drift) cylinder",
            "value": cpu["value"], "unit": B.UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": n / cpu["value"] * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg, "cpu_baseline": cpu,
            "e2e": {"value": cpu["value"], "unit": B.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0})


# ---- disk + halo through the C++ host driver ------------------------------------------------
def _diskhalo_sizes(args, world):
    ntot = args.particles_per_gpu * world if args.particles_per_gpu else args.particles_total
    nd = ntot // 10
    return ntot, nd, ntot - nd


def _diskhalo_config(args, world, ntot, nd, nh):
    return {"workload": f"disk+halo (Cylinder + sphereSL) with mutual interactions, {ntot} particles total "
                        f"({nd} disk + {nh} halo), multistep={args.multistep} (levels assigned by "
                        f"adjust_multistep_level), sphere lmax=6 nmax=18, cylinder mmax=6 norder=18 257x129, "
                        f"C++ host driver do_step over the C ABI (BASELINE configs[3]); {world} GPU(s)",
            "particles_total": ntot, "disk": nd, "halo": nh, "multistep": args.multistep, "dtime": 0.004}


def run_diskhalo(args, B):
    R = B.Ranks()
    from exp_b200 import _lib, ics
    from exp_b200.exphost import Simulation
    L = _lib.lib()
    R.init()
    ntot, nd, nh = _diskhalo_sizes(args, R.world)
    dlo, dhi = B.slice_of(nd, R.rank, R.world)
    hlo, hhi = B.slice_of(nh, R.rank, R.world)
    tabS = B.make_table()
    tabC = cyl_table()
    ph, vh, mh = ics.nfw_halo(nh, lo=hlo, hi=hhi, seed=43)
    ph *= 0.02                               # halo scale radius 0.02: the disk (a = 0.01) sits inside it
    vh *= 7.0
    pd, vd, md = ics.exponential_disk(nd, lo=dlo, hi=dhi, seed=44)
    dtime = 0.004
    sim = Simulation(dtime, args.multistep, dynfracV=0.01, dynfracA=0.03, dynfracP=0.05,
                     comm=R.comm, myid=R.rank, numprocs=R.world, device=R.local)
    h = sim.add_component("halo", ph, vh, mh)
    d = sim.add_component("disk", pd, vd, md)
    sim.set_sphere(h, tabS, scale=0.02, rmin=tabS.rmin * 0.02, rmax=tabS.rmax * 0.02)
    sim.set_cylinder(d, tabC)
    sim.add_interaction(h, d)
    sim.add_interaction(d, h)
    sim.begin_run()
    # nvidia-smi polled every 500 ms here: this workload is launch- and sync-bound on the host side
    # (2000 kernels and ~130 stream synchronisations per master step) and the 50 ms polling of the
    # other workloads slowed the timed region by up to 2x and made it erratic
    clocks = B.ClockSampler(R.local, period_ms=500) if R.rank == 0 else None
    if clocks:
        clocks.start()
    sim.do_steps(max(1, args.warmup // 3))
    # timed region without the per-kernel events (2000 kernels per master step: the event pairs
    # keep consecutive launches from pipelining and cost 25 %); the breakdown comes from the same
    # number of master steps repeated with the events on
    _lib.check(L.expb200_profile_reset())
    _lib.check(L.expb200_profile_enable(0))
    R.barrier()
    if clocks:
        clocks.mark_begin()
    steps = max(1, args.steps // 4)
    ms = R.max(sim.do_steps_timed(steps))
    if clocks:
        clocks.mark_end()
    clk = clocks.stop() if clocks else None
    launches = sum(v[0] for v in _lib.profile_read().values())
    _lib.check(L.expb200_profile_reset())
    _lib.check(L.expb200_profile_enable(1))
    ms_profiled = R.max(sim.do_steps_timed(steps))
    _lib.check(L.expb200_profile_enable(0))
    prof = _lib.profile_read()
    counts_h = sim.level_report(h)["counts"].astype(np.float64)
    counts_d = sim.level_report(d)["counts"].astype(np.float64)
    sub = float(sum((counts_h[l] + counts_d[l]) * (1 << l) for l in range(args.multistep + 1)))
    sub = R.sum(sub)
    lv = [R.sum(float(counts_h[l] + counts_d[l])) for l in range(args.multistep + 1)]
    if R.comm:
        _lib.check(L.expb200_comm_check(R.comm))
    if R.rank != 0:
        R.close()
        return
    per = {k: round(v[1] / steps, 4) for k, v in prof.items() if v[2]}
    line = {"metric": "particle-steps/sec (master steps; coef+force+kick-drift, multistep) disk+halo",
            "value": ntot * steps / (ms * 1e-3), "unit": B.UNIT, "n_gpus": R.world, "steps": steps,
            "warmup": max(1, args.warmup // 3), "ms_per_step": ms / steps, "higher_is_better": True,
            "scaling": "weak" if args.particles_per_gpu else "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": dict(_diskhalo_config(args, R.world, ntot, nd, nh), allreduce=R.allreduce_desc(),
                           level_populations=lv),
            "particle_substeps_per_sec": sub * steps / (ms * 1e-3),
            "binning": {name: dict(zip(("whole_component_sorts", "active_tail_sorts", "tail_items", "whole_items"),
                                       (int(x) for x in sim.sort_counters(i))))
                        for name, i in (("disk", d), ("halo", h))},
            "clocks": clk, "e2e": None, "gpu_launches": launches,
            "ms_per_step_with_kernel_events": ms_profiled / steps,
            "roofline": {"bound": "fp64", "kernel": "whole master step (see kernel_ms_per_step)", "achieved": None,
                         "peak": None, "unit": "TFLOP/s", "frac": None, "traffic": None,
                         "kernel_ms_per_step": per},
            "cpu_baseline": None if (R.world > 1 or args.no_cpu) else _cpu_diskhalo(args, B)}
    B.emit(line)
    R.close()


def _cpu_diskhalo(args, B, n=100_000, steps=1):
    """one master step of the same disk+halo set-up on `n` particles with the oracle-driven stepper
    (tests/ref_stepper.py: the reference's do_step ordering over the oracle's C functions), one thread"""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests"))
    from ref_stepper import RefSim
    from exp_b200 import ics
    nd = n // 10
    nh = n - nd
    tabS, tabC = B.make_table(), cyl_table()
    ph, vh, mh = ics.nfw_halo(nh, seed=43)
    ph *= 0.02
    vh *= 7.0
    pd, vd, md = ics.exponential_disk(nd, seed=44)
    sim = RefSim(0.004, args.multistep, dynfracV=0.01, dynfracA=0.03, dynfracP=0.05)
    rh = sim.add("sphere", tabS, ph, vh, mh, scale=0.02)
    rd = sim.add("cylinder", tabC, pd, vd, md)
    sim.add_interaction(rh, rd)
    sim.add_interaction(rd, rh)
    sim.initialize()
    t0 = time.perf_counter()
    for _ in range(steps):
        sim.step()
    dt = time.perf_counter() - t0
    lv = np.bincount(np.concatenate([sim.S[rh].level, sim.S[rd].level]), minlength=args.multistep + 1)
    return {"value": n * steps / dt, "unit": "particle-steps/s", "cores": 1, "kind": "port",
            "sample": f"{nd} disk + {nh} halo particles of the same distributions, {steps} master step(s) of "
                      f"{1 << args.multistep} sub-steps in {dt:.1f} s, one thread: tests/ref_stepper.py (the reference's "
                      f"do_step / adjust_multistep_level ordering in Python) over the oracle's C functions; level "
                      f"populations {lv.tolist()}"}


def ref_diskhalo(args, B):
    R = B.Ranks()
    if R.rank != 0:
        return
    world = max(args.gpus, 1)
    ntot, nd, nh = _diskhalo_sizes(args, world)
    n = 100_000
    cpu = _cpu_diskhalo(args, B, n=n, steps=1)
    cfg = _diskhalo_config(args, world, ntot, nd, nh)
    cfg["reference_sample_particles_per_step"] = n
    cfg["extrapolated"] = True
    B.emit({"impl": "reference",
            "metric": "particle-steps/sec (master steps; coef+force+kick-drift, multistep) disk+halo",
            "value": cpu["value"], "unit": B.UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": n / cpu["value"] * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg, "cpu_baseline": cpu,
            "e2e": {"value": cpu["value"], "unit": B.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0})


# ---- pyEXP createFromArray sweep --------------------------------------------------------------
def run_pyexp(args, B):
    R = B.Ranks()
    from exp_b200 import _lib, ics, pyexp
    L = _lib.lib()
    n = args.particles_per_gpu or (args.particles_total if args.particles_total != 100_000_000 else 10_000_000)
    nsnap = max(args.steps, 1)
    tab = B.make_table()
    sph = pyexp.SphericalSL(tab)
    ctab = cyl_table()
    cyl = pyexp.Cylindrical(ctab)
    pos, _, mass = ics.nfw_halo(n, seed=43)
    dpos, _, dmass = ics.exponential_disk(n, seed=44)
    out = {}
    _lib.check(L.expb200_profile_reset())
    for name, basis, p, m in (("sphereSL", sph, pos, mass), ("cylinder", cyl, dpos, dmass)):
        times = []
        for k in range(nsnap + args.warmup):
            q = p * (1.0 + 1e-3 * np.sin(1000.0 + k))          # a different snapshot every time
            t0 = time.perf_counter()
            basis.createFromArray(m, q, time=float(k))
            times.append(time.perf_counter() - t0)
        out[name] = float(np.mean(times[args.warmup:]))
    per_snap = out["sphereSL"] + out["cylinder"]
    cpu = None
    if not args.no_cpu:
        from oracle import pyexp_oracle as PO
        m = 20000
        t0 = time.perf_counter()
        PO.sph_create_from_array(tab, mass[:m], pos[:m]) if hasattr(PO, "sph_create_from_array") else None
        dt = time.perf_counter() - t0
        if dt > 0 and hasattr(PO, "sph_create_from_array"):
            cpu = {"value": m / dt, "unit": "particles/s", "cores": 1, "kind": "port",
                   "sample": f"{m} particles through oracle/pyexp_oracle.py (a Python restatement of the serial "
                             f"expui accumulate loop, expui/BiorthBasis.cc:4712-4735): not a performance baseline"}
    B.emit({"metric": "particles/sec (basis.createFromArray, sphereSL + cylinder, host arrays in, coefficients out)",
            "value": 2 * n / per_snap, "unit": "particles/s", "n_gpus": 1, "steps": nsnap, "warmup": args.warmup,
            "ms_per_step": per_snap * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"pyEXP basis.createFromArray sweep over {nsnap} synthetic {n}-particle snapshots "
                                   f"(sphereSL lmax=6 nmax=18 + cylinder mmax=6 norder=18), pageable numpy arrays "
                                   f"(BASELINE configs[4])", "particles_per_snapshot": n, "snapshots": nsnap},
            "snapshot_ms": {k: v * 1e3 for k, v in out.items()},
            "e2e": {"value": 2 * n / per_snap, "unit": "particles/s", "h2d_bytes_per_step": 2 * n * 32,
                    "d2h_bytes_per_step": (28 * 18 + 7 * 18) * 16, "path": "the value itself is end to end"},
            "gpu_launches": sum(v[0] for v in _lib.profile_read().values()),
            "roofline": {"bound": "pcie", "achieved": 2 * n * 32 / per_snap / 1e9, "peak": 55.0, "unit": "GB/s",
                         "frac": 2 * n * 32 / per_snap / 1e9 / 55.0, "traffic": None,
                         "peak_source": "measured pinned H2D on this pool (gpurun_out/pcie_probe.txt): 55.6 GB/s"},
            "cpu_baseline": cpu})


def ref_pyexp(args, B):
    R = B.Ranks()
    if R.rank != 0:
        return
    B.emit({"impl": "reference", "unavailable": "pyEXP cannot be built here (needs Eigen3, HDF5, MPI, pybind11 build of "
            "the reference); oracle/pyexp_oracle.py is a Python restatement used for parity only"})
