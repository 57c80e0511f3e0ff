"""TEST INFRASTRUCTURE: the reference's do_step / begin_run / adjust_multistep_level
ordering (src/step.cc:67-323, src/begin.cc:86-127, src/multistep.cc:344-560) driven with
the CPU oracle's functions, for one self-gravitating sphereSL component.  levlist[lev] is
kept as ascending particle indices (the reference's order is hash-map order, i.e. not
canonical; sums are order-independent to rounding)."""
import numpy as np

from oracle import oracle as O


class RefSphereStepper:
    def __init__(self, tab, pos, vel, mass, dtime, multistep, **dyn):
        self.g = O.Grid(tab)
        self.P = O.sph_params(rmin=tab.rmin, rmax=tab.rmax)
        self.S = O.ParticleSet(pos, vel, mass)
        self.dtime, self.multistep, self.dyn = dtime, multistep, dyn
        self.mintvl, self.mfirst, self.dL, self.dN = O.multistep_tables(multistep)
        self.Mstep = 1 << multistep
        shape = ((tab.lmax + 1) ** 2, tab.nmax)
        self.coefN = np.zeros((multistep + 1,) + shape)
        self.coefL = np.zeros((multistep + 1,) + shape)
        self.expcoef = np.zeros(shape)
        self.this_step = 0
        self.levlist = [self.S.levlist(l) for l in range(multistep + 1)]

    def compute_expansion(self, M):
        coef, _ = O.sph_coefficients(self.g, self.P, self.S, lst=self.levlist[M])
        if self.multistep:
            self.coefL[M], self.coefN[M] = self.coefN[M].copy(), coef      # src/SphericalBasis.cc:785-788
        else:
            self.expcoef = coef

    def compute_potential(self, mlevel, mdrft):
        act = np.concatenate([self.levlist[l] for l in range(mlevel, self.multistep + 1)]).astype(np.int32)
        self.S.acc[act] = 0.0
        self.S.pot[act] = 0.0
        if self.multistep:
            self.expcoef = O.multistep_interp(self.multistep, mdrft, self.coefL, self.coefN)
        O.sph_accel(self.g, self.P, self.S, self.expcoef, lst=act)

    def adjust_multistep_level(self, mdrft, mstep):
        first_call = self.this_step == 0 and mdrft == 0
        first = 0 if (self.this_step == 0 and mstep == 0) else int(self.mfirst[mdrft])
        lst = np.concatenate([self.levlist[l] for l in range(first, self.multistep + 1)]).astype(np.int32)
        old = self.S.level.copy()
        LP = O.level_params(self.dtime, self.multistep, mfirst_mdrft=int(self.mfirst[mdrft]),
                            reset_dtreq=first_call, **self.dyn)
        nsw, which, offlo, offhi = O.adjust_levels(self.S, LP, lst=lst)
        # multistep_update per switching particle + finish (src/SphericalBasis.cc:1013-1228)
        differ = np.zeros((self.multistep + 1, self.expcoef.size))
        for i in np.nonzero(old != self.S.level)[0]:
            O.sph_multistep_update(self.g, self.P, self.S, int(i), int(old[i]), int(self.S.level[i]), differ)
        for M in range(int(self.mfirst[mdrft]), self.multistep + 1):
            self.coefN[M] += differ[M].reshape(self.expcoef.shape)
        self.levlist = [self.S.levlist(l) for l in range(self.multistep + 1)]     # reset_level_lists
        return nsw

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

    def step(self):
        S = self.S
        if self.multistep:
            dt = self.dtime / self.Mstep
            for mstep in range(self.Mstep):
                for M in range(int(self.mfirst[mstep]), self.multistep + 1):
                    DT = dt * int(self.mintvl[M])
                    O.incr_velocity(S, 0.5 * DT, lst=self.levlist[M])
                    O.incr_position(S, DT, lst=self.levlist[M])
                    self.compute_expansion(M)
                mdrft = mstep + 1
                self.compute_potential(int(self.mfirst[mstep]), mdrft)
                for M in range(int(self.mfirst[mdrft]), self.multistep + 1):
                    O.incr_velocity(S, 0.5 * dt * int(self.mintvl[M]), lst=self.levlist[M])
                self.adjust_multistep_level(mdrft, mstep)
        else:
            O.incr_velocity(S, 0.5 * self.dtime)
            O.incr_position(S, self.dtime)
            self.compute_expansion(0)
            self.compute_potential(0, 0)
            O.incr_velocity(S, 0.5 * self.dtime)
        self.this_step += 1
