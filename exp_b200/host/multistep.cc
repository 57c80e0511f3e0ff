// multistep.cc -- level bookkeeping of the multiple-time-step driver:
// initialize_multistep (src/multistep.cc:630-679) and adjust_multistep_level (:344-560).
// The per-particle criterion (adjust_multistep_level_thread, :52-236) runs on the device
// behind expb200_adjust_multistep_level, bit-exact with the CPU path.
#include "expand.H"

namespace exphost {

std::map<Component *, LevelReport> level_report;

void initialize_multistep()
{
  Mstep = 1 << multistep;
  mintvl.assign(1, Mstep);                                   // sub-steps per level
  for (unsigned n = 1; n <= multistep; n++) mintvl.push_back(mintvl.back() / 2);

  // level `lev` is active at sub-step ms when ms is a multiple of its interval
  mactive.assign(Mstep + 1, std::vector<bool>(multistep + 1, false));
  for (unsigned lev = 0; lev <= multistep; lev++) mactive[0][lev] = true;
  for (int ms = 1; ms <= Mstep; ms++)
    for (unsigned lev = 0; lev <= multistep; lev++)
      mactive[ms][lev] = (ms % (1 << (multistep - lev))) == 0;

  mfirst.assign(Mstep + 1, 0);                               // lowest active level per sub-step
  for (int ms = 0; ms <= Mstep; ms++)
    for (unsigned lev = 0; lev <= multistep; lev++)
      if (mactive[ms][lev]) { mfirst[ms] = (int)lev; break; }

  // bracketing sub-steps for the coefficient interpolation of an inactive level
  dstepL.assign(multistep + 1, std::vector<int>(Mstep, 0));
  dstepN.assign(multistep + 1, std::vector<int>(Mstep, 0));
  for (unsigned lev = 0; lev <= multistep; lev++) {
    const int width = 1 << (multistep - lev);
    for (int n = 0; n < Mstep; n++) {
      dstepL[lev][n] = (n / width) * width;
      dstepN[lev][n] = dstepL[lev][n] + width;
    }
  }
}

void adjust_multistep_level()
{
  if (!multistep) return;

  const bool firstCall = this_step == 0 && mdrft == 0;

  for (auto c : comp->components) c->force->multistep_update_begin();

  for (auto c : comp->components) {
    int first = mfirst[mdrft];                       // first active level at the drifted step
    if (this_step == 0 && mstep == 0) first = 0;     // everything on the very first call

    // src/multistep.cc:518-524: level changes are applied only when switching is allowed
    bool apply = !c->NoSwitch() || mdrft == Mstep || firstCall;
    if (!firstCall && c->FreezeLev()) apply = false;

    expb200_level_params p;
    p.dtime = dtime;
    p.dynfracS = dynfracS; p.dynfracD = dynfracD; p.dynfracV = dynfracV;
    p.dynfracA = dynfracA; p.dynfracP = dynfracP;
    p.multistep = (int)multistep;
    p.shiftlevl = shiftlevl;
    p.mfirst_mdrft = mfirst[mdrft];
    p.minlev = first;
    p.noswitch = c->NoSwitch();
    p.reset_dtreq = (c->DTreset() && mstep == 0) || firstCall;          // :91-96
    p.apply = apply;
    p.dim = c->dim;

    expb200_level_stats st;
    EXPHOST_CHECK(expb200_adjust_multistep_level(c->b200, &p, &st));

    LevelReport &r = level_report[c];
    if (mdrft == Mstep || firstCall) r = LevelReport();
    r.examined += st.examined; r.switched += st.switched;
    r.offlo += st.offlo; r.offhi += st.offhi;
    if (st.mindt < r.mindt) r.mindt = st.mindt;
    if (st.maxdt > r.maxdt) r.maxdt = st.maxdt;

    if (apply) {
      // every switching particle's contribution moves between the level tables
      c->force->multistep_update(c);
      c->force->multistep_update_finish();
      // reset_level_lists: the device store re-bins by level on its next sort
    }
    c->fix_positions();                   // src/multistep.cc:543
  }
}

} // namespace exphost
