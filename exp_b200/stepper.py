"""Host-side step driver mirroring the reference's ``do_step`` / ``begin_run`` /
``adjust_multistep_level`` ordering (src/step.cc:67-323, src/begin.cc:86-127,
src/multistep.cc:344-560) for one self-gravitating sphereSL component whose particles,
coefficient tables and level bookkeeping all stay on the GPU.  The reference's engine keeps
this logic as it is; this module exists so that the path can be driven (and tested) without EXP itself."""
from __future__ import annotations

import ctypes as C

from . import _lib
from .component import Component
from .sphere import Sphere


def multistep_tables(multistep: int):
    """initialize_multistep (src/multistep.cc:630-679): Mstep, mintvl[], mfirst[]"""
    Mstep = 1 << multistep
    mintvl = [Mstep >> n for n in range(multistep + 1)]
    mfirst = []
    for ms in range(Mstep + 1):
        for lev in range(multistep + 1):
            if ms == 0 or ms % (1 << (multistep - lev)) == 0:
                mfirst.append(lev)
                break
    return Mstep, mintvl, mfirst


class SphereStepper:
    def __init__(self, comp: Component, sphere: Sphere, dtime: float, multistep: int = 0,
                 comm=None, adb: float = 1.0, **dyn):
        assert sphere.multistep == multistep
        self.c, self.s, self.dtime, self.multistep = comp, sphere, float(dtime), multistep
        self.comm, self.adb, self.dyn = comm, adb, dyn
        self.Mstep, self.mintvl, self.mfirst = multistep_tables(multistep)
        self.this_step = 0
        self.tnow = 0.0
        self.stats = []

    # -- ComponentContainer::compute_expansion (src/ComponentContainer.cc:1173-1233)
    def compute_expansion(self, M: int):
        self.s.set_multistep_level(M)
        self.s.determine_coefficients(self.c, adb=self.adb, to_host=False)
        if self.comm:
            p, st = self.s.coef_device()
            _lib.check(_lib.lib().expb200_allreduce_sum(self.comm, p, self.s.nrow * self.s.nmax, st))
        if self.multistep:
            self.s.multistep_commit(M)

    # -- ComponentContainer::compute_potential (src/ComponentContainer.cc:580-1170)
    def compute_potential(self, mlevel: int, mdrft: int):
        self.c.zero_pot_acc(mlevel)
        self.s.set_multistep_level(mlevel)
        if self.multistep:
            self.s.compute_multistep_coefficients(mdrft)
        else:
            self.s.set_coefficients(None)
        self.s.get_acceleration_and_potential(self.c)

    # -- adjust_multistep_level (src/multistep.cc:344-560)
    def adjust_multistep_level(self, mdrft: int, mstep: int):
        first_call = self.this_step == 0 and mdrft == 0
        first = 0 if (self.this_step == 0 and mstep == 0) else self.mfirst[mdrft]
        st = self.c.adjust_multistep_level(self.dtime, self.multistep, mfirst_mdrft=self.mfirst[mdrft],
                                           minlev=first, reset_dtreq=first_call, **self.dyn)
        self.s.multistep_update(self.c, adb=self.adb, finish=self.comm is None)
        if self.comm:
            p, n, stream = self.s.differ_device()
            _lib.check(_lib.lib().expb200_allreduce_sum(self.comm, p, n, stream))
            self.s.multistep_update_finish()
        self.stats.append(st)
        return st

    # -- begin_run (src/begin.cc:86-127)
    def initialize(self):
        if self.multistep:
            for M in range(self.multistep + 1):
                self.compute_expansion(M)
            self.compute_potential(0, 0)
            self.adjust_multistep_level(0, 0)
        for M in range(self.multistep + 1):
            self.compute_expansion(M)
        self.compute_potential(0, 0)
        self.this_step = 1          # the main loop counts from 1 (src/expand.cc:422)

    # -- do_step (src/step.cc:67-323)
    def step(self):
        c, dtime = self.c, self.dtime
        if self.multistep:
            dt = dtime / self.Mstep
            for mstep in range(self.Mstep):
                for M in range(self.mfirst[mstep], self.multistep + 1):
                    DT = dt * self.mintvl[M]
                    c.incr_velocity(0.5 * DT, M)
                    c.incr_position(DT, M)
                    self.compute_expansion(M)
                self.tnow += dt
                mdrft = mstep + 1
                self.compute_potential(self.mfirst[mstep], mdrft)
                for M in range(self.mfirst[mdrft], self.multistep + 1):
                    c.incr_velocity(0.5 * dt * self.mintvl[M], M)
                self.adjust_multistep_level(mdrft, mstep)
        else:
            self.tnow += dtime
            c.incr_velocity(0.5 * dtime)
            c.incr_position(dtime)
            self.compute_expansion(0)
            self.compute_potential(0, 0)
            c.incr_velocity(0.5 * dtime)
        self.this_step += 1
