// Cylinder.cc -- see Cylinder.H
#include "Cylinder.H"

#include <cmath>

#include "global.H"

namespace exphost {

Cylinder::Cylinder(const EmpCylSLTables &t, int device, double rcylmax_, double acyl_, bool even_m,
                   int MMIN, int MLIM, int NMIN, int NLIM)
  : mmax(t.mmax), ncylorder(t.norder), rcylmax(rcylmax_), EVEN_M(even_m)
{
  id = "Cylinder";             // src/Cylinder.cc:99
  acyl = acyl_ < 0.0 ? t.ascale : acyl_;
  expb200_cylinder_config cfg = {t.mmax, t.norder, t.numx, t.numy, t.cmapr, t.cmapz, t.ascale, t.hscale,
                                 t.rmin, t.rmax, rcylmax, acyl, EVEN_M, MMIN, MLIM, NMIN, NLIM,
                                 (int)multistep};
  EXPHOST_CHECK(expb200_cylinder_create(device, &cfg, t.potC, t.potS, t.rforceC, t.rforceS, t.zforceC,
                                        t.zforceS, &b200c));
}

Cylinder::~Cylinder()
{
  if (b200c) expb200_cylinder_destroy(b200c);
}

void Cylinder::set_orientation(const double *body, const double *orig)
{
  EXPHOST_CHECK(expb200_cylinder_set_orientation(b200c, body, orig));
}

void Cylinder::determine_coefficients(Component *c)
{
  if (!self_consistent && !firstime_coef && !initializing) return;

  // subsample / PCA gating (src/Cylinder.cc:1004-1025)
  compute = false;
  covar = false;
  if (nint) {
    if (mstep == 0 && this_step % nint == 0) { compute = true; covar = true; }
  }
  if (pcavar || pcaeof) {
    if (this_step >= npca0) compute = (mstep == 0) && !((this_step - npca0) % npca);
  }

  uint64_t use1 = 0;
  double mass1 = 0.0;
  EXPHOST_CHECK(expb200_cylinder_determine_coefficients(b200c, c->b200, (int)mlevel, c->Adiabatic(),
                                                        nullptr, nullptr, &use1, &mass1));
  if (compute) accumulate_covariance(c);
  // EmpCylSL::make_coefficients' two MPI_Allreduce calls and the cylmass one
  // (exputil/EmpCylSL.cc:4192-4214, src/Cylinder.cc:1093-1095) -> one device all-reduce of
  // [cos | sin | mass]
  if (nccl_comm) {
    double *d = nullptr; size_t n = 0; void *st = nullptr;
    EXPHOST_CHECK(expb200_cylinder_coef_device(b200c, &d, &n, &st));
    EXPHOST_CHECK(expb200_allreduce_sum(nccl_comm, d, n, st));
    EXPHOST_CHECK(expb200_cylinder_get_mass(b200c, 4, 0, &mass1));
  }
  // used / cylmass only grow on the first sub-step after multistep_reset (src/Cylinder.cc:1091-1099)
  if (tnow == resetT) { used += (int)use1; cylmass += mass1; }

  if (multistep == 0 || !self_consistent) {
    EXPHOST_CHECK(expb200_cylinder_set_coefficients(b200c, nullptr, nullptr, cylmass));
  } else {
    EXPHOST_CHECK(expb200_cylinder_multistep_commit(b200c, (int)mlevel));
  }
  firstime_coef = false;
}

// The compute branch of EmpCylSL::accumulate for this level's particles, added to the per-bin
// arrays (zeroed at mlevel == 0), and at the last level their reduction over ranks
// (EmpCylSL::make_coefficients / make_covar).
void Cylinder::accumulate_covariance(Component *c)
{
  const size_t M1 = mmax + 1, N = ncylorder;
  if (sampT == 0) {
    sampT = defSampT ? defSampT : (unsigned)std::floor(std::sqrt((double)c->CurTotal()));
    if (sampT < 1) sampT = 1;
  }
  const size_t T = sampT;
  if (mlevel == 0) {
    numbT.assign(T, 0); massT.assign(T, 0.0);
    if (covar) { VC.assign(T * M1 * N * 2, 0.0); MV.assign(T * M1 * N * N * 2, 0.0); }
    if (pcavar) { covV.assign(T * M1 * N, 0.0); covM.assign(T * M1 * N * N, 0.0); }
    if (pcaeof) tvar.assign(M1 * N * N, 0.0);
  }
  std::vector<unsigned> nb(T);
  std::vector<double> ms(T);
  bool counted = false;
  auto add_counts = [&]() {
    if (counted) return;
    for (size_t k = 0; k < T; k++) { numbT[k] += nb[k]; massT[k] += ms[k]; }
    counted = true;
  };
  if (covar) {
    std::vector<double> v(VC.size()), m(MV.size());
    EXPHOST_CHECK(expb200_cylinder_covariance(b200c, c->b200, (int)mlevel, c->Adiabatic(), (int)T, c->indx_base,
                                              nb.data(), ms.data(), v.data(), m.data()));
    for (size_t k = 0; k < v.size(); k++) VC[k] += v[k];
    for (size_t k = 0; k < m.size(); k++) MV[k] += m[k];
    add_counts();
  }
  if (pcavar) {
    std::vector<double> v(covV.size()), m(covM.size());
    EXPHOST_CHECK(expb200_cylinder_covariance_pca(b200c, c->b200, (int)mlevel, c->Adiabatic(), (int)T,
                                                  c->indx_base, 1, nb.data(), ms.data(), v.data(), m.data()));
    for (size_t k = 0; k < v.size(); k++) covV[k] += v[k];
    for (size_t k = 0; k < m.size(); k++) covM[k] += m[k];
    add_counts();
  }
  if (pcaeof) {
    std::vector<double> m(tvar.size());
    EXPHOST_CHECK(expb200_cylinder_covariance_pca(b200c, c->b200, (int)mlevel, c->Adiabatic(), 1, 0, 2,
                                                  nullptr, nullptr, nullptr, m.data()));
    for (size_t k = 0; k < m.size(); k++) tvar[k] += m[k];
  }
  if (mlevel == multistep && nccl_comm) {
    auto red = [&](std::vector<double> &a) {
      if (!a.empty()) EXPHOST_CHECK(expb200_allreduce_sum_host(nccl_comm, a.data(), a.size()));
    };
    std::vector<double> cnt(numbT.begin(), numbT.end());
    red(cnt);
    for (size_t k = 0; k < cnt.size(); k++) numbT[k] = (unsigned)std::llround(cnt[k]);
    red(massT); red(VC); red(MV); red(covV); red(covM); red(tvar);
  }
}

void Cylinder::compute_multistep_coefficients()
{
  EXPHOST_CHECK(expb200_cylinder_compute_multistep_coefficients(b200c, mdrft, cylmass));
}

void Cylinder::get_acceleration_and_potential(Component *cC)
{
  if (!use_external) {
    if (multistep && (self_consistent || initializing)) compute_multistep_coefficients();
  }
  EXPHOST_CHECK(expb200_cylinder_determine_acceleration(b200c, cC->b200, (int)mlevel));
}

void Cylinder::RegisterComponent()
{
  EXPHOST_CHECK(expb200_cylinder_set_owner(b200c, component ? component->b200 : nullptr));
}

void Cylinder::multistep_reset()
{
  used = 0;
  cylmass = 0.0;
  resetT = tnow;
}

void Cylinder::multistep_update(Component *c)
{
  if (!self_consistent) return;
  EXPHOST_CHECK(expb200_cylinder_multistep_update(b200c, c->b200, c->Adiabatic(), 0));
  pending_update = true;
}

void Cylinder::multistep_update_finish()
{
  if (!pending_update) return;
  if (nccl_comm) {
    double *d = nullptr; size_t n = 0; void *st = nullptr;
    EXPHOST_CHECK(expb200_cylinder_differ_device(b200c, &d, &n, &st));
    EXPHOST_CHECK(expb200_allreduce_sum(nccl_comm, d, n, st));
  }
  EXPHOST_CHECK(expb200_cylinder_multistep_update_finish(b200c));
  pending_update = false;
}

void Cylinder::get_coefs(std::vector<double> &cosv, std::vector<double> &sinv)
{
  cosv.assign(ncoef(), 0.0);
  sinv.assign(ncoef(), 0.0);
  EXPHOST_CHECK(expb200_cylinder_get_level_coefficients(b200c, 3, 0, cosv.data(), sinv.data()));
}

} // namespace exphost
