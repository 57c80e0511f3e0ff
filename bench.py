#!/usr/bin/env python
"""bench.py -- particle-steps/s of the sphereSL basis-function-expansion force loop.

One "step" = one single-level KDK leapfrog step of a self-gravitating sphereSL
component (reference src/step.cc:272-322): kick(dt/2), drift(dt), coefficient
accumulation (+ all-reduce of the coefficient vector across GPUs), force
evaluation, kick(dt/2).

Workload (BASELINE.json configs[1]): NFW halo, 1e7 particles per GPU, lmax=6,
nmax=18, numr=2000; weak scaling (every rank owns its own 1e7-particle slice of
one global N*1e7-body halo, regenerated from the counter-based IC stream).

    python bench.py --gpus N --steps K --warmup W           # ours
    python bench.py --impl reference --steps K --warmup W    # reference CPU path (oracle)

Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# Algorithmic HBM bytes per particle-step: binning 100 (histogram 4,
# rank 8, apply 88), deposit 2 x 36, force + kicks + drift + key 168.
STEP_BYTES = 340


def _hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0          # B200_PROFILING.md fallback


HBM_PEAK_GBS = _hbm_peak()
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full
# capture of this workload (profiles/r01_ncu_*): bytes per launch at 1e7 particles
NCU_TRAFFIC = {"force": 1.645e9, "deposit": 7.42e8}   # profiles/r01f_ncu_full_summary.txt

METRIC = "particle-steps/sec (coef+force+kick-drift) sphereSL"
UNIT = "particle-steps/s"
LMAX, NMAX, NUMR = 6, 18, 2000
RMIN, RMAX = 1.0e-3, 20.0
DT = 1.0e-3


# ---- algorithmic work per particle (DESIGN.md section "Work model") ----------
def flops_force(L):
    """sph_force2p_kernel: 31 flop per (l,m>0) term (Legendre recursion 4,
    derivative 3, Horner forms 12, column sums 12), 19 per (l,0) term, 21 per m>0 column
    (azimuthal recursion + combination), 120 for geometry, coordinate map and the Cartesian
    projection; FMA = 2."""
    return 31 * L * (L + 1) // 2 + 19 * (L + 1) + 21 * L + 120


def flops_deposit(L):
    """sph_deposit_reg_kernel: 14 flop per (l,m>0) term (recursion 4, 2 products, 4 node FMAs),
    8 per (l,0) term, 4 per m>0 column, 60 for geometry and weights (counted once although the
    two register sweeps both evaluate it)."""
    return 14 * L * (L + 1) // 2 + 8 * (L + 1) + 4 * L + 60


def flops_reference(L, N):
    """the reference's own arithmetic per particle-step"""
    return 6 * (L + 1) ** 2 * N + 14 * (L + 1) * N + 11.5 * (L + 1) * (L + 2) + 218


def make_table():
    from exp_b200 import slgrid
    return slgrid.make_slgrid(slgrid.NFWModel(), LMAX, NMAX, numr=NUMR, rmin=RMIN, rmax=RMAX)


def make_particles(n_per_gpu, rank, world):
    from exp_b200 import ics
    ntot = n_per_gpu * world
    return ics.nfw_halo(ntot, lo=rank * n_per_gpu, hi=(rank + 1) * n_per_gpu, seed=43)


# ---- clocks ------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)),
                "power_w_max": float(max(power)), "samples": len(sm), "reasons": sorted(reasons)}


# ---- pinned host arrays --------------------------------------------------------
def pinned_empty(shape, dtype=np.float64):
    from exp_b200 import _lib
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    _lib.check(_lib.lib().expb200_host_alloc(C.byref(p), max(nbytes, 8)))
    buf = (C.c_char * max(nbytes, 8)).from_address(p.value)
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


# ---- CPU reference arm -----------------------------------------------------------
def cpu_step_rate(tab, pos, vel, mass, cores, steps, warmup):
    """The oracle (line-faithful C++ restatement of the reference CPU path) running the same
    step on `cores` host threads with the reference's thread structure: contiguous slices,
    per-thread partials, ordered sum."""
    from oracle import oracle as O
    g = O.Grid(tab)
    P = O.sph_params(rmin=tab.rmin, rmax=tab.rmax)
    S = O.ParticleSet(pos, vel, mass)
    coef, _ = O.sph_coefficients(g, P, S, nthreads=cores)
    O.sph_accel(g, P, S, coef, nthreads=cores)

    def one():
        O.incr_velocity(S, 0.5 * DT)
        O.incr_position(S, DT)
        coef, _ = O.sph_coefficients(g, P, S, nthreads=cores)
        S.acc[:] = 0.0
        S.pot[:] = 0.0
        O.sph_accel(g, P, S, coef, nthreads=cores)
        O.incr_velocity(S, 0.5 * DT)

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = time.perf_counter() - t0
    return len(pos) * steps / dt, dt / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    tab = make_table()
    n = args.cpu_sample
    pos, vel, mass = make_particles(n, 0, 1)
    rate, sec = cpu_step_rate(tab, pos, vel, mass, cores, args.steps, max(args.warmup, 1))
    sample = (f"{n} particles of the NFW workload per step (bounded sample of the "
              f"{args.particles_per_gpu}-particle/GPU job), {cores} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args, max(args.gpus, 1)),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(args, world):
    return {"workload": f"sphereSL NFW halo, {args.particles_per_gpu} particles/GPU, lmax={LMAX} "
                        f"nmax={NMAX} numr={NUMR}, single-level KDK leapfrog (BASELINE configs[1])",
            "particles_per_gpu": args.particles_per_gpu, "particles_total": args.particles_per_gpu * world,
            "lmax": LMAX, "nmax": NMAX, "numr": NUMR, "dt": DT,
            "parallelism": f"particles sharded over {world} GPU(s), coefficient all-reduce",
            "cache": "inputs larger than L2 (>= 140 B/particle resident state >> 126 MB)"}


# ---- our arm ------------------------------------------------------------------------
def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    n = args.particles_per_gpu

    from exp_b200 import _lib
    from exp_b200.component import Component
    from exp_b200.sphere import Sphere
    L = _lib.lib()          # raises if the CUDA library is missing: no CPU fallback

    dist = None
    comm = None
    peer = False
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            raw = C.create_string_buffer(128)
            _lib.check(L.expb200_comm_unique_id(raw))
            idbuf.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
        dist.broadcast(idbuf, 0)
        raw = bytes(idbuf.cpu().numpy().tobytes())
        comm = C.c_void_p()
        _lib.check(L.expb200_comm_create(local, world, rank, raw, C.byref(comm)))
        # the 882-double coefficient all-reduce by one kernel over NVLink peer memory (rank-ordered,
        # bitwise identical on every rank) instead of ncclAllReduce; EXPB200_NO_PEER=1 keeps NCCL
        if not os.environ.get("EXPB200_NO_PEER"):
            try:
                _lib.connect_peers(comm, dist, world)
                peer = True
            except Exception as e:          # e.g. no P2P between two devices: NCCL does it
                print(f"[bench] peer all-reduce unavailable ({e}); using NCCL", file=sys.stderr)

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    tab = make_table()
    pos0, vel0, mass0 = make_particles(n, rank, world)
    # pinned host images of the particle fields (the host side of the plugin boundary)
    hpos, hvel, hmass = pinned_empty((n, 3)), pinned_empty((n, 3)), pinned_empty((n,))
    hacc, hpot = pinned_empty((n, 3)), pinned_empty((n,))
    hpos[:], hvel[:], hmass[:] = pos0, vel0, mass0
    del pos0, vel0, mass0

    c = Component(n, device=local)
    s = Sphere(tab, device=local)
    c.upload(hpos, hvel, hmass)
    # initial force so that the first half-kick has accelerations
    s.determine_coefficients(c, to_host=False)
    if comm:
        p, st = s.coef_device()
        _lib.check(L.expb200_allreduce_sum(comm, p, s.nrow * s.nmax, st))
    s.set_coefficients(None)
    c.zero_pot_acc()
    s.get_acceleration_and_potential(c)
    c.sync()

    # ---------------- resident throughput: K fused steps, device-timed ----------------
    s.steps(c, DT, args.warmup, comm=comm)
    c.sync()
    clocks = ClockSampler(local) if rank == 0 else None
    _lib.check(L.expb200_profile_reset())
    _lib.check(L.expb200_profile_enable(1))
    barrier()
    c.sync()
    if clocks:
        clocks.start()
    c.event_record(0)
    s.steps(c, DT, args.steps, comm=comm)          # K KDK steps (expb200_sphere_steps)
    c.event_record(1)
    c.sync()
    barrier()
    ms = max_over_ranks(c.event_elapsed_ms(0, 1))
    clk = clocks.stop() if clocks else None
    _lib.check(L.expb200_profile_enable(0))
    prof = _lib.profile_read()
    launches = sum(v[0] for v in prof.values())
    value = n * world * args.steps / (ms * 1e-3)

    # ---------------- end to end through the plugin calls, host buffers ----------------
    def e2e_step():
        c.upload_range(0, hpos, hvel, hmass)                 # H2D of the step's inputs
        s.step(c, DT, comm=comm)
        c.download_range(0, n, hpos, hvel, hacc, hpot)       # D2H of the step's result

    for _ in range(max(1, min(args.warmup, 2))):
        e2e_step()
    c.sync()
    barrier()
    c.event_record(2)
    for _ in range(args.e2e_steps):
        e2e_step()
    c.event_record(3)
    c.sync()
    barrier()
    ms_e2e = max_over_ranks(c.event_elapsed_ms(2, 3))
    e2e_value = n * world * args.e2e_steps / (ms_e2e * 1e-3)
    h2d = n * (24 + 24 + 8)
    d2h = n * (24 + 24 + 24 + 8)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel ----------------
    peak_tf = C.c_double(0.0)
    hint = C.c_double(0.0)
    _lib.check(L.expb200_fp64_peak(local, C.byref(peak_tf), C.byref(hint)))
    kern_ms = {k: v[1] / max(v[2], 1) for k, v in prof.items() if v[2]}
    tot = sum(v[1] for v in prof.values()) or 1.0
    shares = {k: round(v[1] / tot, 4) for k, v in prof.items() if v[2]}
    dom = max(("force", "deposit"), key=lambda k: kern_ms.get(k, 0.0))
    # the force launch of a multi-step call also does both half kicks and the drift (18 flop)
    fl = flops_force(LMAX) + 18 if dom == "force" else flops_deposit(LMAX)
    ms_dom = prof[dom][1] / args.steps          # per step (the deposit is two launches per step)
    achieved = fl * n / (ms_dom * 1e-3) / 1e12
    step_fl = flops_force(LMAX) + flops_deposit(LMAX) + 18
    roofline = {
        "bound": "fp64", "kernel": {"force": f"sph_force2p_kernel<{LMAX},64,6,4,2>",
                                    "deposit": f"sph_deposit_reg_kernel<{LMAX},*> (2 sweeps)"}[dom],
        "achieved": achieved, "peak": peak_tf.value, "unit": "TFLOP/s", "frac": achieved / peak_tf.value,
        "peak_source": "measured: dependency-free DFMA microbenchmark on this device "
                       "(MEASURED_PEAKS.json has no FP64 figure)",
        "flop_per_particle": fl, "avg_launch_ms": ms_dom, "traffic": NCU_TRAFFIC.get(dom),
        "flop_model": "flops of the factorised formulation this kernel executes (DESIGN.md section 4), "
                      "not SURVEY 8d's reference-formulation count",
        "frac_survey_8d_formula": flops_reference(LMAX, NMAX) * n / (ms / args.steps * 1e-3) / 1e12 / peak_tf.value,
        "kernel_ms_per_step": {k: round(v[1] / args.steps, 4) for k, v in prof.items() if v[2]},
        "kernel_share": shares,
        "step_flop_per_particle": step_fl,
        "step_fp64_frac": step_fl * n / (ms / args.steps * 1e-3) / 1e12 / peak_tf.value,
        "reference_flop_per_particle": flops_reference(LMAX, NMAX),
        "step_hbm_bytes_per_particle": STEP_BYTES,
        "step_hbm_frac": STEP_BYTES * n / (ms / args.steps * 1e-3) / 1e9 / HBM_PEAK_GBS,
    }

    # ---------------- CPU baseline (oracle port of the reference CPU path) ----------------
    cpu = None
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        m = min(args.cpu_sample, n)
        rate, sec = cpu_step_rate(tab, np.array(hpos[:m]), np.array(hvel[:m]), np.array(hmass[:m]),
                                  cores, 2, 1)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"first {m} particles of the workload, 2 timed steps after 1 warm-up, "
                         f"{cores} host threads (oracle/exp_oracle.cc)"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(workload_config(args, world),
                       allreduce=("none" if world == 1 else
                                  "one kernel over NVLink peer memory, rank-ordered sum" if peer else "ncclAllReduce")),
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d * world,
                "d2h_bytes_per_step": d2h * world, "ms_per_step": ms_e2e / args.e2e_steps,
                "steps": args.e2e_steps,
                "path": "Component.upload_range -> Sphere.step (kick, drift, coefficients, force, kick) "
                        "-> Component.download_range, pinned host buffers"},
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


def emit(line):
    """the ONE JSON line, on the process's original stdout"""
    os.write(_STDOUT_FD, (json.dumps(line) + "\n").encode())


# Libraries (NCCL's version banner, torchrun) write to fd 1; keep the contract's single JSON
# line clean by pointing fd 1 at stderr for everything except emit().
_STDOUT_FD = os.dup(1)
os.dup2(2, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--particles-per-gpu", type=int, default=10_000_000)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-sample", type=int, default=1_000_000)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
