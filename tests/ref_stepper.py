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


# ---------------------------------------------------------------------------------------
# Several components (sphereSL and/or cylinder forces) with mutual interactions: the
# reference's ComponentContainer::compute_expansion / compute_potential
# (src/ComponentContainer.cc:580-1233) around the same do_step ordering.
# ---------------------------------------------------------------------------------------
class RefForce:
    """one force + its per-level coefficient tables, oracle-driven"""

    def __init__(self, kind, tab, multistep, **kw):
        self.kind, self.multistep = kind, multistep
        if kind == "sphere":
            self.g = O.Grid(tab)
            sc = kw.get("scale", 1.0)
            self.P = O.sph_params(scale=sc, rmin=kw.get("rmin", tab.rmin * sc), rmax=kw.get("rmax", tab.rmax * sc))
            shape = (1, (tab.lmax + 1) ** 2, tab.nmax)
        else:
            self.g = O.CylGridPy(tab)
            self.P = O.cyl_params(rcylmax=kw.get("rcylmax", 20.0), acyl=kw.get("acyl", tab.ascale))
            shape = (2, tab.mmax + 1, tab.norder)          # [cos | sin]
        self.N = np.zeros((multistep + 1,) + shape)
        self.L = np.zeros((multistep + 1,) + shape)
        self.use = np.zeros(shape)
        self.cylmass, self.resetT, self.used = 0.0, 0.0, 0

    def reset(self, tnow):                                  # multistep_reset
        self.cylmass, self.used, self.resetT = 0.0, 0, tnow

    def coefficients(self, S, lst, M, tnow):
        if self.kind == "sphere":
            coef, used = O.sph_coefficients(self.g, self.P, S, lst=lst)
            new = coef[None]
        else:
            cs, sn, used, mass = O.cyl_coefficients(self.g, self.P, S, lst=lst)
            new = np.stack([cs, sn])
            if tnow == self.resetT:                          # src/Cylinder.cc:1091-1099
                self.cylmass += mass
        if self.multistep:
            self.L[M], self.N[M] = self.N[M].copy(), new
        else:
            self.use = new

    def interpolate(self, mdrft):
        self.use = O.multistep_interp(self.multistep, mdrft,
                                      self.L.reshape(self.multistep + 1, -1),
                                      self.N.reshape(self.multistep + 1, -1)).reshape(self.use.shape)

    def accel(self, S, lst):
        if self.kind == "sphere":
            O.sph_accel(self.g, self.P, S, self.use[0], lst=lst)
        else:
            O.cyl_accel(self.g, self.P, S, self.use[0], self.use[1], self.cylmass, lst=lst)

    def switch(self, S, old, mfirst_mdrft):
        differ = np.zeros_like(self.N)
        for i in np.nonzero(old != S.level)[0]:
            if self.kind == "sphere":
                d = differ.reshape(self.multistep + 1, -1)
                O.sph_multistep_update(self.g, self.P, S, int(i), int(old[i]), int(S.level[i]), d)
            else:
                dC = np.ascontiguousarray(differ[:, 0])
                dS = np.ascontiguousarray(differ[:, 1])
                O.cyl_multistep_update(self.g, S, int(i), int(old[i]), int(S.level[i]), dC, dS)
                differ[:, 0], differ[:, 1] = dC, dS
        for M in range(mfirst_mdrft, self.multistep + 1):
            self.N[M] += differ[M]


class RefSim:
    def __init__(self, dtime, multistep, **dyn):
        self.dtime, self.multistep, self.dyn = dtime, multistep, dyn
        self.mintvl, self.mfirst, _, _ = O.multistep_tables(multistep)
        self.Mstep = 1 << multistep
        self.S, self.F, self.inter = [], [], []
        self.this_step, self.tnow = 0, 0.0
        self.com = {}                         # component -> [mtot, com, cov, coa] (fix_positions)
        self.centerlevl = multistep // 2      # src/ComponentContainer.cc:42-45

    def add(self, kind, tab, pos, vel, mass, **kw):
        self.S.append(O.ParticleSet(pos, vel, mass))
        self.F.append(RefForce(kind, tab, self.multistep, **kw))
        return len(self.S) - 1

    def add_interaction(self, frm, on):
        self.inter.append((frm, on))

    def _lev(self, S, lo):
        return np.nonzero(S.level >= lo)[0].astype(np.int32)

    def compute_expansion(self, M):
        for S, F in zip(self.S, self.F):
            F.coefficients(S, S.levlist(M), M, self.tnow)

    def fix_positions(self):
        # Component::fix_positions_cpu, default mlevel = 0 at both call sites
        # (src/ComponentContainer.cc:1380, src/multistep.cc:543)
        for k, S in enumerate(self.S):
            _, self.com[k] = O.fix_positions(S, 0, self.multistep)

    def compute_potential(self, mlevel, mdrft, mstep=0):
        for S, F in zip(self.S, self.F):
            act = self._lev(S, mlevel)
            S.acc[act] = 0.0
            S.pot[act] = 0.0
            if self.multistep:
                F.interpolate(mdrft)
            F.accel(S, act)
        for frm, on in self.inter:
            self.F[frm].accel(self.S[on], self._lev(self.S[on], mlevel))
        # mactive[mstep][centerlevl] (src/ComponentContainer.cc:955, src/multistep.cc:646-654)
        if mstep % (1 << (self.multistep - self.centerlevl)) == 0:
            self.fix_positions()

    def adjust_multistep_level(self, mdrft, mstep):
        first_call = self.this_step == 0 and mdrft == 0
        first = 0 if (self.this_step == 0 and mstep == 0) else int(self.mfirst[mdrft])
        nsw = 0
        for S, F in zip(self.S, self.F):
            old = S.level.copy()
            LP = O.level_params(self.dtime, self.multistep, mfirst_mdrft=int(self.mfirst[mdrft]),
                                reset_dtreq=first_call or mstep == 0, **self.dyn)
            n, _, _, _ = O.adjust_levels(S, LP, lst=self._lev(S, first))
            F.switch(S, old, int(self.mfirst[mdrft]))
            nsw += n
        self.fix_positions()
        return nsw

    def reset(self):
        for F in self.F:
            F.reset(self.tnow)

    def initialize(self):
        if self.multistep:
            self.reset()
            for M in range(self.multistep + 1):
                self.compute_expansion(M)
            self.compute_potential(0, 0)
            self.adjust_multistep_level(0, 0)
        self.reset()
        for M in range(self.multistep + 1):
            self.compute_expansion(M)
        self.compute_potential(0, 0)
        self.this_step = 1

    def step(self):
        self.reset()
        if self.multistep:
            dt = self.dtime / self.Mstep
            for mstep in range(self.Mstep):
                for M in range(int(self.mfirst[mstep]), self.multistep + 1):
                    DT = dt * int(self.mintvl[M])
                    for S in self.S:
                        lst = S.levlist(M)
                        O.incr_velocity(S, 0.5 * DT, lst=lst)
                        O.incr_position(S, DT, lst=lst)
                    self.compute_expansion(M)
                self.tnow += dt
                mdrft = mstep + 1
                self.compute_potential(int(self.mfirst[mstep]), mdrft, mstep)
                for M in range(int(self.mfirst[mdrft]), self.multistep + 1):
                    for S in self.S:
                        O.incr_velocity(S, 0.5 * dt * int(self.mintvl[M]), lst=S.levlist(M))
                self.adjust_multistep_level(mdrft, mstep)
        else:
            self.tnow += self.dtime
            for S in self.S:
                O.incr_velocity(S, 0.5 * self.dtime)
                O.incr_position(S, self.dtime)
            self.compute_expansion(0)
            self.compute_potential(0, 0)
            for S in self.S:
                O.incr_velocity(S, 0.5 * self.dtime)
        self.this_step += 1
