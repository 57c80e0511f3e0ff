#!/usr/bin/env python
"""bench.py -- particle-steps/s of the basis-function-expansion force loop on B200.

Default workload = the configuration BASELINE.json's metric is quoted on: a sphereSL NFW halo
of 1e8 bodies TOTAL (lmax=6, nmax=18, numr=2000), single-level KDK leapfrog, STRONG scaling:
with N GPUs every rank owns a contiguous 1e8/N slice of the same global particle stream and
the ranks all-reduce the 882-double coefficient vector every step.

One "step" = one KDK leapfrog step of the self-gravitating component (reference
src/step.cc:272-322): kick(dt/2), drift(dt), coefficient accumulation (+ all-reduce), force
evaluation, kick(dt/2).

    python bench.py --gpus N --steps K --warmup W                 # ours (torchrun for N > 1)
    python bench.py --impl reference --steps K --warmup W         # reference CPU path (oracle port)
    python bench.py --workload cylinder|diskhalo|pyexp            # BASELINE configs[2..4]
    python bench.py --particles-per-gpu 10000000                  # weak scaling (round-1 layout)

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

# Algorithmic HBM bytes per particle-step of the sphere step: the binning pass moves
# keys and slots (28 B), the first deposit sweep gathers the 32-byte position records through the
# permutation and writes the sorted stream (72 B), the second streams it (36 B), the force pass 168 B
STEP_BYTES = {"sort": 28, "deposit": 108, "force": 168}


def _hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


HBM_PEAK_GBS, HBM_PEAK_SRC = _hbm_peak()

METRIC = "particle-steps/sec (coef+force+kick-drift) sphereSL"
UNIT = "particle-steps/s"
LMAX, NMAX, NUMR = 6, 18, 2000
RMIN, RMAX = 1.0e-3, 20.0
DT = 1.0e-3


# ---- algorithmic work per particle (DESIGN.md section "Work model") ----------
def flops_force(L):
    """sph_force kernel: 31 flop per (l,m>0) term (Legendre recursion 4,
    derivative 3, Horner forms 12, column sums 12), 19 per (l,0) term, 21 per m>0 column
    (azimuthal recursion + combination), 120 for geometry, coordinate map and the Cartesian
    projection; FMA = 2."""
    return 31 * L * (L + 1) // 2 + 19 * (L + 1) + 21 * L + 120


def flops_deposit(L):
    """sph_deposit kernel: 14 flop per (l,m>0) term (recursion 4, 2 products, 4 node FMAs),
    8 per (l,0) term, 4 per m>0 column, 60 for geometry and weights."""
    return 14 * L * (L + 1) // 2 + 8 * (L + 1) + 4 * L + 60


def flops_reference(L, N):
    """the reference's own arithmetic per particle-step"""
    return 6 * (L + 1) ** 2 * N + 14 * (L + 1) * N + 11.5 * (L + 1) * (L + 2) + 218


def ncu_traffic(kernel_key, n):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernel family, from the
    committed ncu --set full capture of this workload (profiles/ncu_traffic.json, written by
    profiles/ncu_summary.py from the .ncu-rep of the same command).  Captures are per particle
    count; the nearest capture is scaled linearly in n (streaming kernels).  None when no
    capture is committed."""
    try:
        tab = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        ent = tab.get(kernel_key)
        if not ent:
            return None, None
        best = min(ent, key=lambda e: abs(np.log(e["n"] / n)))
        return best["dram_bytes_per_launch"] * (n / best["n"]), best["source"]
    except Exception:
        return None, None


def make_table():
    from exp_b200 import slgrid
    return slgrid.make_slgrid(slgrid.NFWModel(), LMAX, NMAX, numr=NUMR, rmin=RMIN, rmax=RMAX)


def slice_of(ntot, rank, world):
    return ntot * rank // world, ntot * (rank + 1) // world


def make_particles(ntot, lo, hi, out=None):
    """global particles [lo, hi) of the ntot-body NFW stream, generated 16 Mi at a time"""
    from exp_b200 import ics
    if out is None:
        return ics.nfw_halo(ntot, lo=lo, hi=hi, seed=43)
    hpos, hvel, hmass = out
    step = 1 << 24
    for a in range(lo, hi, step):
        b = min(hi, a + step)
        p, v, m = ics.nfw_halo(ntot, lo=a, hi=b, seed=43)
        hpos[a - lo:b - lo], hvel[a - lo:b - lo], hmass[a - lo:b - lo] = p, v, m
    return hpos, hvel, hmass


# ---- clocks ------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, period_ms=50):
        self.gpu = gpu_index
        self.period = period_ms
        self.proc = None
        self.lines = []          # (arrival time, text)
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", str(self.period)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, bufsize=1)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            deadline = time.time() + 3.0          # first sample = the tool is up
            while not self.lines and time.time() < deadline:
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def parse(rows):
            sm, mx, reasons, power = [], [], set(), []
            for _, ln in rows:
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
            return sm, mx, reasons, power

        inside = [r for r in self.lines if self.t0 is not None and self.t0 <= r[0] <= (self.t1 or 1e300) + 0.06]
        window = "timed region"
        if len(inside) < 2:                       # a region shorter than two sampling periods
            inside, window = self.lines, "warm-up + timed region"
        sm, mx, reasons, power = parse(inside)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_min_mhz": float(min(sm)), "sm_max_mhz": float(max(mx)),
                "power_w_max": float(max(power)), "samples": len(sm), "window": window,
                "period_ms": self.period, "reasons": sorted(reasons)}


# ---- pinned host arrays --------------------------------------------------------
def pinned_empty(shape, dtype=np.float64):
    from exp_b200 import _lib
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    _lib.check(_lib.lib().expb200_host_alloc(C.byref(p), max(nbytes, 8)))
    buf = (C.c_char * max(nbytes, 8)).from_address(p.value)
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


# ---- distributed plumbing ---------------------------------------------------------
class Ranks:
    def __init__(self):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None
        self.comm = None
        self.peer = False

    def init(self):
        if self.world == 1:
            return
        from exp_b200 import _lib
        import torch
        import torch.distributed as dist
        L = _lib.lib()
        torch.cuda.set_device(self.local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.dist = dist
        idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if self.rank == 0:
            raw = C.create_string_buffer(128)
            _lib.check(L.expb200_comm_unique_id(raw))
            idbuf.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
        dist.broadcast(idbuf, 0)
        raw = bytes(idbuf.cpu().numpy().tobytes())
        comm = C.c_void_p()
        _lib.check(L.expb200_comm_create(self.local, self.world, self.rank, raw, C.byref(comm)))
        self.comm = comm
        # the 882-double coefficient all-reduce by one kernel over NVLink peer memory (rank-ordered,
        # bitwise identical on every rank) instead of ncclAllReduce; EXPB200_NO_PEER=1 keeps NCCL.
        # Every rank must take the same path: agree on it collectively.
        ok = 0
        raw = C.create_string_buffer(64)
        if not os.environ.get("EXPB200_NO_PEER"):
            try:
                _lib.check(L.expb200_comm_peer_handle(comm, raw))
                ok = 1
            except Exception as e:
                print(f"[bench] rank {self.rank}: peer buffers unavailable ({e})", file=sys.stderr)
        flag = torch.tensor([ok], dtype=torch.int32, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 1:
            got = [None] * self.world
            dist.all_gather_object(got, bytes(raw.raw))
            ok2 = 1
            try:
                _lib.check(L.expb200_comm_peer_connect(comm, b"".join(got)))
            except Exception as e:
                ok2 = 0
                print(f"[bench] rank {self.rank}: peer connect failed ({e})", file=sys.stderr)
            flag = torch.tensor([ok2], dtype=torch.int32, device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) != 1:
                raise RuntimeError("peer all-reduce connected on some ranks only; rerun with EXPB200_NO_PEER=1")
            self.peer = True
        dist.barrier()

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def max(self, x):
        if self.dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, x):
        if self.dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def allreduce_desc(self):
        return ("none" if self.world == 1 else
                "one kernel over NVLink peer memory, rank-ordered sum" if self.peer else "ncclAllReduce")

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


# ---- CPU reference arm -----------------------------------------------------------
def cpu_step_rate(tab, pos, vel, mass, cores, steps, warmup):
    """-> (rate, seconds per step, kind).  kind "reference": the reference's OWN member functions
    (incr_velocity/incr_position threads, SphericalBasis coefficient and force threads, SLGridSph
    interpolation, Basis recursions) compiled from /root/reference where they lie
    (oracle/_ref/libexpref_path.so, built in the build container and shipped with the snapshot) and
    driven in do_step's order on `cores` host threads with the reference's fork/join structure.
    kind "port" (only when that library is absent): the line-faithful restatement in
    oracle/exp_oracle.cc, which tests/test_oracle_ref.py pins to it bit for bit."""
    from oracle import ref_path as RP
    if RP.available():
        st = RP.RefSphereStepper(tab, pos, vel, mass, nthreads=cores)
        st.force()
        for _ in range(warmup):
            st.step(DT)
        t0 = time.perf_counter()
        for _ in range(steps):
            st.step(DT)
        dt = time.perf_counter() - t0
        return len(pos) * steps / dt, dt / steps, "reference"
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
    return len(pos) * steps / dt, dt / steps, "port"


CPU_KIND_NOTE = {"reference": "the reference's own SphericalBasis/SLGridSph/Basis/leapfrog thread functions compiled in "
                              "place (oracle/_ref/libexpref_path.so) over a flat particle store",
                 "port": "oracle/exp_oracle.cc (restatement pinned bit for bit to the reference's functions)"}


def workload_config(args, world, n_local):
    ntot = args.particles_total
    mode = "weak" if args.particles_per_gpu else "strong"
    return {"workload": f"sphereSL NFW halo, {ntot} bodies total, lmax={LMAX} nmax={NMAX} numr={NUMR}, "
                        f"single-level KDK leapfrog (the configuration of BASELINE.json's metric; "
                        f"{mode} scaling: {n_local} bodies per GPU on {world} GPU(s))",
            "particles_total": ntot, "particles_per_gpu": n_local,
            "lmax": LMAX, "nmax": NMAX, "numr": NUMR, "dt": DT,
            "parallelism": f"particles sharded over {world} GPU(s), coefficient all-reduce",
            "cache": "inputs larger than L2 (>= 112 B/particle resident state >> 126 MB)"}


def resolve_sizes(args, world):
    """particles_total and the per-rank slice for the run"""
    if args.particles_per_gpu:
        args.particles_total = args.particles_per_gpu * world
    return args.particles_total


def run_reference(args):
    R = Ranks()
    if R.rank != 0:
        return
    world = max(args.gpus, 1)
    ntot = resolve_sizes(args, world)
    cores = os.cpu_count() or 1
    tab = make_table()
    n = min(args.cpu_sample, ntot)
    pos, vel, mass = make_particles(ntot, 0, n)
    rate, sec, kind = cpu_step_rate(tab, pos, vel, mass, cores, args.steps, max(args.warmup, 1))
    sample = (f"{n} particles of the NFW workload per step (the first {n} of the {ntot}-body stream; the "
              f"rate is per particle and extrapolates linearly), {cores} threads; {CPU_KIND_NOTE[kind]}")
    cfg = workload_config(args, world, slice_of(ntot, 0, world)[1])
    cfg["reference_sample_particles_per_step"] = n      # what this arm actually timed
    cfg["extrapolated"] = n != ntot
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak" if args.particles_per_gpu else "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---- our arm: sphere (default) ---------------------------------------------------------
def run_sphere(args):
    R = Ranks()
    from exp_b200 import _lib
    from exp_b200.component import Component
    from exp_b200.sphere import Sphere
    L = _lib.lib()          # raises if the CUDA library is missing: no CPU fallback
    R.init()
    rank, world, local = R.rank, R.world, R.local
    ntot = resolve_sizes(args, world)
    lo, hi = slice_of(ntot, rank, world)
    n = hi - lo

    tab = make_table()
    # pinned host images of the particle fields (the host side of the plugin boundary)
    hpos, hvel, hmass = pinned_empty((n, 3)), pinned_empty((n, 3)), pinned_empty((n,))
    make_particles(ntot, lo, hi, out=(hpos, hvel, hmass))
    nbuf = 2 if args.e2e_pipeline else 1
    hout = [dict(pos=pinned_empty((n, 3)), vel=pinned_empty((n, 3)), acc=pinned_empty((n, 3)),
                 pot=pinned_empty((n,))) for _ in range(nbuf)]

    comps = [Component(n, device=local) for _ in range(nbuf)]
    sphs = [Sphere(tab, device=local) for _ in range(nbuf)]
    for c, s in zip(comps, sphs):
        c.upload(hpos, hvel, hmass)
        # initial force so that the first half-kick has accelerations
        s.determine_coefficients(c, to_host=False)
        if R.comm:
            p, st = s.coef_device()
            _lib.check(L.expb200_allreduce_sum(R.comm, p, s.nrow * s.nmax, st))
        s.set_coefficients(None)
        c.zero_pot_acc()
        s.get_acceleration_and_potential(c)
        c.sync()
    c, s = comps[0], sphs[0]

    # ---------------- resident throughput: K fused steps, device-timed ----------------
    clocks = ClockSampler(local) if rank == 0 else None
    if clocks:
        clocks.start()
    s.steps(c, DT, args.warmup, comm=R.comm)
    c.sync()
    # The timed region runs without the per-kernel events (a pair of timing events around every
    # kernel keeps consecutive launches from pipelining: ~1 % here, 25 % in the launch-heavy
    # multistep workload); the launch counters count regardless.  The per-kernel times of the
    # roofline come from the same K steps repeated right after it with the events on.
    _lib.check(L.expb200_profile_reset())
    _lib.check(L.expb200_profile_enable(0))
    R.barrier()
    c.sync()
    if clocks:
        clocks.mark_begin()
    c.event_record(0)
    s.steps(c, DT, args.steps, comm=R.comm)          # K KDK steps (expb200_sphere_steps)
    c.event_record(1)
    c.sync()
    if clocks:
        clocks.mark_end()
    R.barrier()
    ms = R.max(c.event_elapsed_ms(0, 1))
    clk = clocks.stop() if clocks else None
    launches = sum(v[0] for v in _lib.profile_read().values())
    value = ntot * args.steps / (ms * 1e-3)
    _lib.check(L.expb200_profile_reset())
    _lib.check(L.expb200_profile_enable(1))
    c.event_record(2)
    s.steps(c, DT, args.steps, comm=R.comm)          # the same K steps again, per-kernel events on
    c.event_record(3)
    c.sync()
    _lib.check(L.expb200_profile_enable(0))
    prof = _lib.profile_read()
    ms_profiled = R.max(c.event_elapsed_ms(2, 3))

    # ---------------- end to end through the plugin calls, host buffers ----------------
    # Every step: H2D of pos/vel/mass from pinned host memory, the step, D2H of pos/vel/acc/pot.
    # Two independent batches are kept in flight on two device stores (double buffering): the
    # host thread issues upload+step of batch B (asynchronous), then blocks in the download of
    # batch A, so A's D2H overlaps B's H2D and kernels (PCIe is full duplex, two copy engines).
    def issue(k):
        comps[k].upload_range(0, hpos, hvel, hmass)                 # H2D of the step's inputs
        sphs[k].step(comps[k], DT, comm=R.comm)

    def collect(k):
        o = hout[k]
        comps[k].download_range(0, n, o["pos"], o["vel"], o["acc"], o["pot"])   # D2H of the step's result

    # One continuous run of W + K steps; the clock starts when the W-th result has landed on the
    # host and stops when the (W+K)-th has: K complete steps (inputs up, results down) of the
    # pipeline in steady state.  collect() returns only when the D2H copies are complete.
    def e2e_run(warm, steps):
        t0 = None
        total = warm + steps
        if nbuf == 1:
            for i in range(total):
                if i == warm:
                    c.sync(); R.barrier(); t0 = time.perf_counter()
                issue(0)
                collect(0)
            return (time.perf_counter() - t0) * 1e3
        issue(0)
        for i in range(1, total + 1):
            if i < total:
                issue(i & 1)
            collect((i - 1) & 1)               # result of step i-1 is on the host
            if i == warm:
                t0 = time.perf_counter()
        return (time.perf_counter() - t0) * 1e3

    R.barrier()
    wall_e2e = e2e_run(max(2, min(args.warmup, 3)), args.e2e_steps)
    for cc in comps:
        cc.sync()
    R.barrier()
    ms_e2e = R.max(wall_e2e)
    e2e_value = ntot * args.e2e_steps / (ms_e2e * 1e-3)
    h2d = n * (24 + 24 + 8)
    d2h = n * (24 + 24 + 24 + 8)
    h2d_tot, d2h_tot = int(R.sum(h2d)), int(R.sum(d2h))
    if R.comm:
        _lib.check(L.expb200_comm_check(R.comm))

    if rank != 0:
        R.close()
        return

    # ---------------- roofline of the dominant kernel ----------------
    peak_tf = C.c_double(0.0)
    hint = C.c_double(0.0)
    _lib.check(L.expb200_fp64_peak(local, C.byref(peak_tf), C.byref(hint)))
    kern_ms = {k: v[1] / max(v[2], 1) for k, v in prof.items() if v[2]}
    tot = sum(v[1] for v in prof.values()) or 1.0
    shares = {k: round(v[1] / tot, 4) for k, v in prof.items() if v[2]}
    dom = max(("force", "deposit"), key=lambda k: prof.get(k, (0, 0.0, 0))[1])
    # the force launch of a multi-step call also does both half kicks and the drift (18 flop)
    fl = flops_force(LMAX) + 18 if dom == "force" else flops_deposit(LMAX)
    ms_dom = prof[dom][1] / args.steps          # per step
    achieved = fl * n / (ms_dom * 1e-3) / 1e12
    step_fl = flops_force(LMAX) + flops_deposit(LMAX) + 18
    nominal_tf = 148 * 64 * 2 * 1.965e9 / 1e12          # 148 SMs x 64 DFMA/clk x 2 flop x 1.965 GHz
    traffic, traffic_src = ncu_traffic(dom, n)
    step_bytes = sum(STEP_BYTES.values())
    roofline = {
        "bound": "fp64", "kernel": {"force": f"sph_force2p_kernel<{LMAX},...> (force + both half kicks + drift + next key)",
                                    "deposit": f"sph_deposit_reg_kernel<{LMAX},*>"}[dom],
        "achieved": achieved, "peak": peak_tf.value, "unit": "TFLOP/s", "frac": achieved / peak_tf.value,
        "peak_source": "measured: dependency-free DFMA microbenchmark on this device in this run "
                       "(MEASURED_PEAKS.json has no FP64 figure)",
        "peak_nominal": nominal_tf, "frac_of_nominal": achieved / nominal_tf,
        "flop_per_particle": fl, "avg_launch_ms": ms_dom, "traffic": traffic, "traffic_source": traffic_src,
        "algorithmic_bytes_per_launch": STEP_BYTES[dom] * n,
        "flop_model": "flops of the factorised formulation this kernel executes (DESIGN.md section 4), "
                      "not SURVEY 8d's reference-formulation count",
        "frac_survey_8d_formula": flops_reference(LMAX, NMAX) * n / (ms / args.steps * 1e-3) / 1e12 / peak_tf.value,
        "kernel_ms_per_step": {k: round(v[1] / args.steps, 4) for k, v in prof.items() if v[2]},
        "kernel_times_from": "the K steps repeated right after the timed region with a pair of CUDA events "
                             "around every kernel, on the launching stream",
        "ms_per_step_with_kernel_events": ms_profiled / args.steps,
        "kernel_share": shares,
        "step_flop_per_particle": step_fl,
        "step_fp64_frac": step_fl * n / (ms / args.steps * 1e-3) / 1e12 / peak_tf.value,
        "reference_flop_per_particle": flops_reference(LMAX, NMAX),
        "step_hbm_bytes_per_particle": step_bytes,
        "step_hbm_frac": step_bytes * n / (ms / args.steps * 1e-3) / 1e9 / HBM_PEAK_GBS,
        "hbm_peak_gbs": HBM_PEAK_GBS, "hbm_peak_source": HBM_PEAK_SRC,
    }

    # ---------------- CPU baseline (oracle port of the reference CPU path) ----------------
    cpu = None
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        m = min(args.cpu_sample, n)
        ksteps = args.cpu_steps
        rate, sec, kind = cpu_step_rate(tab, np.array(hpos[:m]), np.array(hvel[:m]), np.array(hmass[:m]),
                                        cores, ksteps, 1)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"first {m} particles of the workload, {ksteps} timed steps after 1 warm-up "
                         f"({sec * ksteps:.1f} s of CPU work), {cores} host threads; {CPU_KIND_NOTE[kind]}"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak" if args.particles_per_gpu else "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": dict(workload_config(args, world, n), allreduce=R.allreduce_desc()),
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_tot,
                "d2h_bytes_per_step": d2h_tot, "ms_per_step": ms_e2e / args.e2e_steps,
                "steps": args.e2e_steps,
                "path": "Component.upload_range -> Sphere.step (kick, drift, coefficients, force, kick) "
                        "-> Component.download_range, pinned host buffers"
                        + ("; two batches in flight on two device stores so that the D2H of one step "
                           "overlaps the H2D + kernels of the next (full-duplex PCIe); host clock from the "
                           "arrival of result W to the arrival of result W+K" if nbuf == 2 else ""),
                "timer": "host perf_counter around blocking downloads (each returns after its D2H completed)"},
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    emit(line)
    R.close()


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
    ap.add_argument("--workload", default="sphere", choices=["sphere", "cylinder", "diskhalo", "pyexp"])
    ap.add_argument("--particles-total", type=int, default=100_000_000,
                    help="bodies in the whole job (strong scaling: split over the GPUs)")
    ap.add_argument("--particles-per-gpu", type=int, default=0,
                    help="weak scaling: this many bodies on every GPU (overrides --particles-total)")
    ap.add_argument("--e2e-steps", type=int, default=8)
    ap.add_argument("--no-e2e-pipeline", dest="e2e_pipeline", action="store_false",
                    help="one device store: upload, step, download strictly in sequence")
    ap.add_argument("--cpu-sample", type=int, default=4_000_000)
    ap.add_argument("--cpu-steps", type=int, default=8)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--multistep", type=int, default=4)
    ap.add_argument("--sampT", type=int, default=100)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.workload != "sphere":
        import bench_workloads
        return bench_workloads.run(args, sys.modules[__name__])
    if args.impl == "reference":
        run_reference(args)
    else:
        run_sphere(args)


if __name__ == "__main__":
    main()
