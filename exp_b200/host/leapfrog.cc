// leapfrog.cc -- incr_position / incr_velocity (src/incpos.cc:73-150, src/incvel.cc:91-170):
// the GPU branch (incr_position_cuda, src/incpos.cc:81-86) for every component.
#include "expand.H"

namespace exphost {

void incr_position(double dt, int mlevel)
{
  for (auto c : comp->components) EXPHOST_CHECK(expb200_incr_position(c->b200, dt, mlevel));
}

void move_levels(const double *dt_kick, const double *dt_drift, int mlevel_lo)
{
  for (auto c : comp->components)
    EXPHOST_CHECK(expb200_move_levels(c->b200, dt_kick, dt_drift, (int)multistep + 1, mlevel_lo));
}

void incr_velocity(double dt, int mlevel)
{
  for (auto c : comp->components) EXPHOST_CHECK(expb200_incr_velocity(c->b200, dt, mlevel));
}

} // namespace exphost
