"""Host-side mirror of the reference's ``Component`` particle accessors for the
device-resident store (reference src/Component.H:1082-1139 ``ParticlesToCuda`` /
``CudaToParticles`` / ``ZeroPotAccel``; leapfrog src/incpos.cc, src/incvel.cc;
level assignment src/multistep.cc).  Thin wrapper over the C ABI."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import c_dp, c_fp, c_up, check


def _dp(a):
    return a.ctypes.data_as(c_dp) if a is not None else None


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


class Component:
    def __init__(self, capacity: int, device: int = 0):
        self._h = C.c_void_p()
        self.device = device
        self.capacity = int(capacity)
        check(_lib.lib().expb200_component_create(device, self.capacity, C.byref(self._h)))

    def close(self):
        if self._h:
            _lib.lib().expb200_component_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- ParticlesToCuda / CudaToParticles --------------------------------
    def upload(self, pos, vel=None, mass=None, level=None, scale=None, potext=None):
        pos = _f64(pos)
        n = len(pos)
        vel, mass, potext = _f64(vel), _f64(mass), _f64(potext)
        if mass is None:
            mass = np.full(n, 1.0 / max(n, 1))
        lev = None if level is None else np.ascontiguousarray(level, dtype=np.uint32)
        sc = None if scale is None else np.ascontiguousarray(scale, dtype=np.float32)
        check(_lib.lib().expb200_component_upload(
            self._h, n, _dp(pos), _dp(vel), _dp(mass),
            lev.ctypes.data_as(c_up) if lev is not None else None,
            sc.ctypes.data_as(c_fp) if sc is not None else None, _dp(potext)))

    def upload_range(self, first, pos, vel, mass):
        """Asynchronous chunk upload; arrays must stay alive until sync()."""
        check(_lib.lib().expb200_component_upload_range(self._h, first, len(pos), _dp(pos),
                                                        _dp(vel), _dp(mass)))

    def set_count(self, n):
        check(_lib.lib().expb200_component_set_count(self._h, n))

    def sync(self):
        check(_lib.lib().expb200_component_sync(self._h))

    @property
    def n(self):
        return _lib.lib().expb200_component_count(self._h)

    def download(self, pos=True, vel=True, acc=True, pot=True, level=False, dtreq=False):
        n = self.n
        out = {}
        if pos: out["pos"] = np.empty((n, 3))
        if vel: out["vel"] = np.empty((n, 3))
        if acc: out["acc"] = np.empty((n, 3))
        if pot: out["pot"] = np.empty(n)
        if level: out["level"] = np.empty(n, dtype=np.uint32)
        if dtreq: out["dtreq"] = np.empty(n, dtype=np.float32)
        check(_lib.lib().expb200_component_download(
            self._h, _dp(out.get("pos")), _dp(out.get("vel")), _dp(out.get("acc")),
            _dp(out.get("pot")),
            out["level"].ctypes.data_as(c_up) if level else None,
            out["dtreq"].ctypes.data_as(c_fp) if dtreq else None))
        return out

    def download_range(self, first, count, pos=None, vel=None, acc=None, pot=None):
        check(_lib.lib().expb200_component_download_range(self._h, first, count, _dp(pos), _dp(vel),
                                                          _dp(acc), _dp(pot)))

    def set_center(self, center=(0.0, 0.0, 0.0), rtrunc=1.0e20):
        c = (C.c_double * 3)(*[float(x) for x in center])
        check(_lib.lib().expb200_component_set_center(self._h, c, float(rtrunc)))

    def set_levels(self, level):
        lev = np.ascontiguousarray(level, dtype=np.uint32)
        check(_lib.lib().expb200_component_set_levels(self._h, lev.ctypes.data_as(c_up)))

    def set_acc_pot(self, acc, pot):
        acc, pot = _f64(acc), _f64(pot)
        check(_lib.lib().expb200_component_set_acc_pot(self._h, _dp(acc), _dp(pot)))

    # ---- ZeroPotAccel / leapfrog / levels ---------------------------------
    def zero_pot_acc(self, minlev=0):
        check(_lib.lib().expb200_component_zero_pot_acc(self._h, minlev))

    def incr_position(self, dt, mlevel=-1):
        check(_lib.lib().expb200_incr_position(self._h, float(dt), mlevel))

    def incr_velocity(self, dt, mlevel=-1):
        check(_lib.lib().expb200_incr_velocity(self._h, float(dt), mlevel))

    def move_levels(self, dt_kick, dt_drift, mlevel_lo=0):
        """the per-level kicks and drifts of one sub-step in one pass (expb200_move_levels): level M >=
        mlevel_lo gets vel += acc dt_kick[M], then pos += vel dt_drift[M]; either table may be None"""
        k = None if dt_kick is None else _f64(dt_kick)
        d = None if dt_drift is None else _f64(dt_drift)
        nlev = len(k if k is not None else d)
        check(_lib.lib().expb200_move_levels(self._h, None if k is None else _dp(k),
                                             None if d is None else _dp(d), nlev, mlevel_lo))

    def adjust_multistep_level(self, dtime, multistep, mfirst_mdrft=0, minlev=0, dynfracS=1.0,
                               dynfracD=1.0e32, dynfracV=0.01, dynfracA=0.03, dynfracP=0.05,
                               shiftlevl=0, noswitch=False, reset_dtreq=False, apply=True, dim=3):
        """Defaults as in the reference (src/global.cc:70-74)."""
        P = _lib.LevelParams(dtime, dynfracS, dynfracD, dynfracV, dynfracA, dynfracP, multistep,
                             shiftlevl, mfirst_mdrft, minlev, int(noswitch), int(reset_dtreq),
                             int(apply), dim)
        S = _lib.LevelStats()
        check(_lib.lib().expb200_adjust_multistep_level(self._h, C.byref(P), C.byref(S)))
        return dict(examined=S.examined, switched=S.switched, offlo=S.offlo, offhi=S.offhi,
                    mindt=S.mindt, maxdt=S.maxdt)

    def event_record(self, slot):
        check(_lib.lib().expb200_component_event_record(self._h, slot))

    def event_elapsed_ms(self, a, b):
        ms = C.c_double(0.0)
        check(_lib.lib().expb200_component_event_elapsed(self._h, a, b, C.byref(ms)))
        return ms.value

    def level_counts(self, multistep):
        out = np.zeros(multistep + 1, dtype=np.uint64)
        check(_lib.lib().expb200_component_level_counts(self._h, multistep,
                                                        out.ctypes.data_as(_lib.c_u64p)))
        return out

    def sort_counters(self):
        """[whole-component sorts, tail-of-active-levels sorts, particles moved by the tail sorts,
        particles moved by the whole-component sorts] since creation"""
        out = np.zeros(4, dtype=np.uint64)
        check(_lib.lib().expb200_component_sort_counters(self._h, out.ctypes.data_as(_lib.c_u64p)))
        return out

    def fix_positions(self, mlevel, multistep, lev_sums=None, dim=3, nccl_comm=None):
        """Component::fix_positions (src/Component.cc:3354-3590) on the device: refreshes the rows
        M >= mlevel of lev_sums[M] = [com_mas, com_lev[3], cov_lev[3], coa_lev[3]] and returns
        (lev_sums, [mtot, com[3], cov[3], coa[3]])."""
        if lev_sums is None:
            lev_sums = np.zeros((multistep + 1, 10))
        lev_sums = np.ascontiguousarray(lev_sums, dtype=np.float64)
        check(_lib.lib().expb200_component_fix_positions(self._h, mlevel, multistep, dim,
                                                         lev_sums.ctypes.data_as(_lib.c_dp), nccl_comm))
        out = lev_sums.sum(axis=0)
        if out[0] > 0.0:
            out[1:] /= out[0]
        return lev_sums, out
