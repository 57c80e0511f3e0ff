"""ctypes loader for libexpb200.so (the C ABI declared in include/expb200.h).

The product path has no CPU fallback: if the CUDA library is missing or a call
fails, this module raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libexpb200.so")
CSRC = os.path.join(_HERE, "csrc")
_LIB = None

c_dp = C.POINTER(C.c_double)
c_up = C.POINTER(C.c_uint)
c_fp = C.POINTER(C.c_float)
c_u64p = C.POINTER(C.c_uint64)

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


class Expb200Error(RuntimeError):
    pass


class SphereConfig(C.Structure):
    _fields_ = [("lmax", C.c_int), ("nmax", C.c_int), ("numr", C.c_int), ("cmap", C.c_int),
                ("rmin_grid", C.c_double), ("rmax_grid", C.c_double), ("rmap", C.c_double),
                ("scale", C.c_double), ("rmin", C.c_double), ("rmax", C.c_double),
                ("NO_L0", C.c_int), ("NO_L1", C.c_int), ("EVEN_L", C.c_int), ("EVEN_M", C.c_int),
                ("M0_only", C.c_int), ("multistep", C.c_int)]


class CylinderConfig(C.Structure):
    _fields_ = [("mmax", C.c_int), ("norder", C.c_int), ("numx", C.c_int), ("numy", C.c_int),
                ("cmapr", C.c_int), ("cmapz", C.c_int), ("ascale", C.c_double),
                ("hscale", C.c_double), ("rmin", C.c_double), ("rmax", C.c_double),
                ("rcylmax", C.c_double), ("acyl", C.c_double), ("EVEN_M", C.c_int),
                ("MMIN", C.c_int), ("MLIM", C.c_int), ("NMIN", C.c_int), ("NLIM", C.c_int),
                ("multistep", C.c_int)]


class LevelParams(C.Structure):
    _fields_ = [("dtime", C.c_double), ("dynfracS", C.c_double), ("dynfracD", C.c_double),
                ("dynfracV", C.c_double), ("dynfracA", C.c_double), ("dynfracP", C.c_double),
                ("multistep", C.c_int), ("shiftlevl", C.c_int), ("mfirst_mdrft", C.c_int),
                ("minlev", C.c_int), ("noswitch", C.c_int), ("reset_dtreq", C.c_int),
                ("apply", C.c_int), ("dim", C.c_int)]


class LevelStats(C.Structure):
    _fields_ = [("examined", C.c_uint64), ("switched", C.c_uint64), ("offlo", C.c_uint64),
                ("offhi", C.c_uint64), ("mindt", C.c_double), ("maxdt", C.c_double)]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into exp_b200/lib/libexpb200.so: one object
    per .cu (in parallel, only when stale), then one link."""
    from concurrent.futures import ThreadPoolExecutor
    srcs = sources()
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(_HERE, "..", "include", "expb200.h"))
    hdr_time = max(os.path.getmtime(h) for h in hdrs)
    objdir = os.path.join(_HERE, "lib", "obj")
    os.makedirs(objdir, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != "-shared"]

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        if (not force and os.path.exists(obj)
                and os.path.getmtime(obj) >= max(os.path.getmtime(src), hdr_time)):
            return obj, None
        cmd = ["nvcc"] + flags + ["-c", "-o", obj, src]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        return obj, (r.stdout + r.stderr if r.returncode != 0 else None)

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(compile_one, srcs))
    errs = [e for _, e in results if e]
    if errs:
        raise Expb200Error("nvcc failed:\n" + "\n".join(errs))
    objs = [o for o, _ in results]
    if (force or not os.path.exists(LIB_PATH)
            or any(os.path.getmtime(o) > os.path.getmtime(LIB_PATH) for o in objs)):
        cmd = ["nvcc", "-shared", "-o", LIB_PATH] + objs + ["-ldl"]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise Expb200Error("link failed:\n" + r.stdout + r.stderr)
    return LIB_PATH


HOST_DIR = os.path.join(_HERE, "host")
HOST_LIB_PATH = os.path.join(_HERE, "lib", "libexphost.so")


def build_host(force: bool = False, verbose: bool = False) -> str:
    """Compile the C++ host driver (exp_b200/host/*.cc: Component, PotAccel, SphericalBasis,
    Cylinder, ComponentContainer, do_step ...) into exp_b200/lib/libexphost.so, linked
    against libexpb200.so ($ORIGIN rpath)."""
    srcs = sorted(os.path.join(HOST_DIR, f) for f in os.listdir(HOST_DIR) if f.endswith(".cc"))
    deps = srcs + [os.path.join(HOST_DIR, f) for f in os.listdir(HOST_DIR) if f.endswith((".H", ".h"))]
    deps += [os.path.join(_HERE, "..", "include", "expb200.h"), LIB_PATH]
    if (not force and os.path.exists(HOST_LIB_PATH)
            and os.path.getmtime(HOST_LIB_PATH) >= max(os.path.getmtime(d) for d in deps)):
        return HOST_LIB_PATH
    cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-fPIC", "-shared", "-I" + os.path.join(_HERE, "..", "include"),
           "-o", HOST_LIB_PATH] + srcs + ["-L" + os.path.dirname(LIB_PATH), "-lexpb200",
                                          "-Wl,-rpath,$ORIGIN"]
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise Expb200Error("host build failed:\n" + r.stdout + r.stderr)
    return HOST_LIB_PATH


# symbol -> (restype, argtypes); every symbol include/expb200.h declares
_H = C.c_void_p
SIGNATURES = {
    "expb200_last_error": (C.c_char_p, []),
    "expb200_version": (C.c_int, []),
    "expb200_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "expb200_device_info": (C.c_int, [C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int),
                                      C.POINTER(C.c_size_t)]),
    "expb200_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "expb200_host_free": (C.c_int, [C.c_void_p]),
    "expb200_host_register": (C.c_int, [C.c_void_p, C.c_size_t]),
    "expb200_host_unregister": (C.c_int, [C.c_void_p]),
    "expb200_component_create": (C.c_int, [C.c_int, C.c_size_t, C.POINTER(_H)]),
    "expb200_component_destroy": (C.c_int, [_H]),
    "expb200_component_upload": (C.c_int, [_H, C.c_size_t, c_dp, c_dp, c_dp, c_up, c_fp, c_dp]),
    "expb200_component_upload_range": (C.c_int, [_H, C.c_size_t, C.c_size_t, c_dp, c_dp, c_dp]),
    "expb200_component_set_count": (C.c_int, [_H, C.c_size_t]),
    "expb200_component_sync": (C.c_int, [_H]),
    "expb200_component_download": (C.c_int, [_H, c_dp, c_dp, c_dp, c_dp, c_up, c_fp]),
    "expb200_component_download_range": (C.c_int, [_H, C.c_size_t, C.c_size_t, c_dp, c_dp, c_dp, c_dp]),
    "expb200_component_count": (C.c_size_t, [_H]),
    "expb200_component_set_center": (C.c_int, [_H, c_dp, C.c_double]),
    "expb200_component_zero_pot_acc": (C.c_int, [_H, C.c_int]),
    "expb200_component_set_levels": (C.c_int, [_H, c_up]),
    "expb200_component_set_acc_pot": (C.c_int, [_H, c_dp, c_dp]),
    "expb200_component_event_record": (C.c_int, [_H, C.c_int]),
    "expb200_component_event_elapsed": (C.c_int, [_H, C.c_int, C.c_int, c_dp]),
    "expb200_component_level_counts": (C.c_int, [_H, C.c_int, c_u64p]),
    "expb200_component_sort_counters": (C.c_int, [_H, c_u64p]),
    "expb200_component_fix_positions": (C.c_int, [_H, C.c_int, C.c_int, C.c_int, c_dp, C.c_void_p]),
    "expb200_incr_position": (C.c_int, [_H, C.c_double, C.c_int]),
    "expb200_incr_velocity": (C.c_int, [_H, C.c_double, C.c_int]),
    "expb200_move_levels": (C.c_int, [_H, c_dp, c_dp, C.c_int, C.c_int]),
    "expb200_adjust_multistep_level": (C.c_int, [_H, C.POINTER(LevelParams), C.POINTER(LevelStats)]),
    "expb200_sphere_create": (C.c_int, [C.c_int, C.POINTER(SphereConfig), c_dp, c_dp, c_dp,
                                        C.POINTER(_H)]),
    "expb200_sphere_destroy": (C.c_int, [_H]),
    "expb200_sphere_set_owner": (C.c_int, [_H, _H]),
    "expb200_cylinder_set_owner": (C.c_int, [_H, _H]),
    "expb200_sphere_determine_coefficients": (C.c_int, [_H, _H, C.c_int, C.c_double, c_dp, c_u64p]),
    "expb200_sphere_coef_device": (C.c_int, [_H, C.POINTER(c_dp), C.POINTER(C.c_void_p)]),
    "expb200_sphere_set_coefficients": (C.c_int, [_H, c_dp]),
    "expb200_sphere_determine_acceleration": (C.c_int, [_H, _H, C.c_int]),
    "expb200_sphere_multistep_commit": (C.c_int, [_H, C.c_int]),
    "expb200_sphere_compute_multistep_coefficients": (C.c_int, [_H, C.c_int]),
    "expb200_sphere_multistep_update": (C.c_int, [_H, _H, C.c_double, C.c_int]),
    "expb200_sphere_multistep_update_finish": (C.c_int, [_H]),
    "expb200_sphere_differ_device": (C.c_int, [_H, C.POINTER(c_dp), C.POINTER(C.c_size_t), C.POINTER(C.c_void_p)]),
    "expb200_sphere_get_level_coefficients": (C.c_int, [_H, C.c_int, C.c_int, c_dp]),
    "expb200_sphere_set_level_coefficients": (C.c_int, [_H, C.c_int, C.c_int, c_dp]),
    "expb200_sphere_covariance": (C.c_int, [_H, _H, C.c_int, C.c_double, C.c_int, C.c_uint64, c_up, c_dp,
                                            c_dp, c_dp]),
    "expb200_sphere_covariance_eof": (C.c_int, [_H, _H, C.c_int, C.c_double, c_dp]),
    "expb200_sphere_step": (C.c_int, [_H, _H, C.c_double, C.c_double, C.c_void_p]),
    "expb200_sphere_steps": (C.c_int, [_H, _H, C.c_double, C.c_int, C.c_double, C.c_void_p]),
    "expb200_cylinder_create": (C.c_int, [C.c_int, C.POINTER(CylinderConfig), c_dp, c_dp, c_dp, c_dp,
                                          c_dp, c_dp, C.POINTER(_H)]),
    "expb200_cylinder_destroy": (C.c_int, [_H]),
    "expb200_cylinder_set_orientation": (C.c_int, [_H, c_dp, c_dp]),
    "expb200_cylinder_determine_coefficients": (C.c_int, [_H, _H, C.c_int, C.c_double, c_dp, c_dp,
                                                          c_u64p, c_dp]),
    "expb200_cylinder_coef_device": (C.c_int, [_H, C.POINTER(c_dp), C.POINTER(C.c_size_t),
                                               C.POINTER(C.c_void_p)]),
    "expb200_cylinder_set_coefficients": (C.c_int, [_H, c_dp, c_dp, C.c_double]),
    "expb200_cylinder_determine_acceleration": (C.c_int, [_H, _H, C.c_int]),
    "expb200_cylinder_multistep_commit": (C.c_int, [_H, C.c_int]),
    "expb200_cylinder_compute_multistep_coefficients": (C.c_int, [_H, C.c_int, C.c_double]),
    "expb200_cylinder_multistep_update": (C.c_int, [_H, _H, C.c_double, C.c_int]),
    "expb200_cylinder_multistep_update_finish": (C.c_int, [_H]),
    "expb200_cylinder_differ_device": (C.c_int, [_H, C.POINTER(c_dp), C.POINTER(C.c_size_t),
                                                 C.POINTER(C.c_void_p)]),
    "expb200_cylinder_get_level_coefficients": (C.c_int, [_H, C.c_int, C.c_int, c_dp, c_dp]),
    "expb200_cylinder_get_mass": (C.c_int, [_H, C.c_int, C.c_int, c_dp]),
    "expb200_cylinder_covariance": (C.c_int, [_H, _H, C.c_int, C.c_double, C.c_int, C.c_uint64, c_up, c_dp,
                                              c_dp, c_dp]),
    "expb200_cylinder_covariance_pca": (C.c_int, [_H, _H, C.c_int, C.c_double, C.c_int, C.c_uint64, C.c_int,
                                                  c_up, c_dp, c_dp, c_dp]),
    "expb200_profile_enable": (C.c_int, [C.c_int]),
    "expb200_profile_reset": (C.c_int, []),
    "expb200_profile_kernels": (C.c_int, []),
    "expb200_profile_name": (C.c_char_p, [C.c_int]),
    "expb200_profile_read": (C.c_int, [C.POINTER(C.c_ulonglong), c_dp, C.POINTER(C.c_ulonglong)]),
    "expb200_fp64_peak": (C.c_int, [C.c_int, c_dp, c_dp]),
    "expb200_dmma_peak": (C.c_int, [C.c_int, C.c_int, c_dp]),
    "expb200_comm_unique_id": (C.c_int, [C.c_char_p]),
    "expb200_comm_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_char_p, C.POINTER(C.c_void_p)]),
    "expb200_comm_destroy": (C.c_int, [C.c_void_p]),
    "expb200_allreduce_sum": (C.c_int, [C.c_void_p, c_dp, C.c_size_t, C.c_void_p]),
    "expb200_comm_check": (C.c_int, [C.c_void_p]),
    "expb200_allreduce_sum_host": (C.c_int, [C.c_void_p, c_dp, C.c_size_t]),
    "expb200_comm_peer_handle": (C.c_int, [C.c_void_p, C.c_char_p]),
    "expb200_comm_peer_connect": (C.c_int, [C.c_void_p, C.c_char_p]),
}


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise Expb200Error(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().expb200_last_error()
        raise Expb200Error(f"expb200 error {rc}: {msg.decode() if msg else '?'}")


def profile_read():
    """{kernel: (launches, ms, timed_regions)} since the last reset."""
    L = lib()
    k = L.expb200_profile_kernels()
    la = (C.c_ulonglong * k)()
    ms = (C.c_double * k)()
    tm = (C.c_ulonglong * k)()
    check(L.expb200_profile_read(la, ms, tm))
    return {L.expb200_profile_name(i).decode(): (int(la[i]), float(ms[i]), int(tm[i])) for i in range(k)}


def connect_peers(comm, dist, world):
    """Switch the small all-reduces of `comm` to the peer-memory kernel: gather every rank's
    IPC handle with torch.distributed (any backend) and connect.  Ranks must share a node."""
    L = lib()
    raw = C.create_string_buffer(64)
    check(L.expb200_comm_peer_handle(comm, raw))
    got = [None] * world
    dist.all_gather_object(got, bytes(raw.raw))
    check(L.expb200_comm_peer_connect(comm, b"".join(got)))
    dist.barrier()
