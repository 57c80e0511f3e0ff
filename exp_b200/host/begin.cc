// begin.cc -- begin_run (src/begin.cc:60-135): level tables, first coefficients, first
// force, first level assignment, then coefficients and force again at the new levels.
#include "expand.H"

namespace exphost {

void begin_run()
{
  initialize_multistep();

  initializing = true;

  if (multistep) {
    comp->multistep_reset();
    for (unsigned M = 0; M <= multistep; M++) comp->compute_expansion(M);   // every level
    comp->compute_potential(0);                                            // every particle
    adjust_multistep_level();                                              // all on first call
  }

  if (multistep) comp->multistep_reset();
  for (unsigned M = 0; M <= multistep; M++) comp->compute_expansion(M);
  comp->compute_potential(0);

  initializing = false;
  this_step = 1;          // the main loop counts steps from 1 (src/expand.cc:422)
}

} // namespace exphost
