// step.cc -- do_step (src/step.cc:67-323): one K(1/2) D K(1/2) leapfrog step; with
// multistep > 0 the hierarchy of 2^multistep sub-steps in which level M moves every
// mintvl[M] sub-steps and the inactive levels' coefficients are interpolated.
#include <cmath>

#include "expand.H"

#include <algorithm>
#include <vector>

namespace exphost {

std::vector<double> pos_check, vel_check;

void do_step(int n)
{
  (void)n;
  comp->multistep_reset();

  if (multistep) {
    const double dt = dtime / Mstep;                   // smallest time step
    pos_check.assign(multistep + 1, 0.0);
    vel_check.assign(multistep + 1, 0.0);
    std::vector<double> dtk(multistep + 1), dtd(multistep + 1);

    for (mstep = 0; mstep < Mstep; mstep++) {
      mdrft = mstep;

      // The levels that move on this sub-step: first half kick and drift, then their coefficients
      // at the drifted positions.  The reference interleaves the three per level
      // (src/step.cc:215-237); level M's kick, drift and coefficients touch level-M particles only,
      // so doing all the moves first -- in one pass over the active levels, expb200_move_levels --
      // gives the same numbers and lets the device bin the moved levels once per sub-step instead
      // of once per level (csrc/sphere.cu: ensure_sorted).
      std::fill(dtk.begin(), dtk.end(), 0.0);
      std::fill(dtd.begin(), dtd.end(), 0.0);
      for (unsigned M = (unsigned)mfirst[mstep]; M <= multistep; M++) {
        const double DT = dt * mintvl[M];
        dtk[M] = 0.5 * DT;                             // first half kick
        vel_check[M] += 0.5 * DT;
        dtd[M] = DT;                                   // drift
        pos_check[M] += DT;
      }
      move_levels(dtk.data(), dtd.data(), mfirst[mstep]);
      for (unsigned M = (unsigned)mfirst[mstep]; M <= multistep; M++)
        comp->compute_expansion(M);                    // coefficients at the drifted position

      tnow += dt;

      mdrft = mstep + 1;                               // drifted position in the level tables
      comp->compute_potential((unsigned)mfirst[mstep]);

      // second half kick for the levels active at the next sub-step
      std::fill(dtk.begin(), dtk.end(), 0.0);
      for (unsigned M = (unsigned)mfirst[mdrft]; M <= multistep; M++) {
        dtk[M] = 0.5 * dt * mintvl[M];
        vel_check[M] += 0.5 * dt * mintvl[M];
      }
      move_levels(dtk.data(), nullptr, mfirst[mdrft]);

      adjust_multistep_level();
    }
  } else {
    tnow += dtime;
    incr_velocity(0.5 * dtime);
    incr_position(dtime);
    comp->compute_expansion(0);
    comp->compute_potential();
    incr_velocity(0.5 * dtime);
  }
  this_step++;
}

} // namespace exphost
