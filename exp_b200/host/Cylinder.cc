// Cylinder.cc -- see Cylinder.H
#include "Cylinder.H"
#include "AxisymmetricBasis.H"

#include <algorithm>
#include <cmath>
#include <limits>

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
  // Hall smoothing (src/Cylinder.cc:1115-1134): the analysis at the last level of a compute step,
  // then the installed coefficients are smoothed in the decorrelated basis
  if (npca0 < (1 << 30) && mlevel == multistep && compute && pcavar) {
    // (EmpCylSL::pca_hall then cleans the partition storage, :4941-4957; here the arrays are
    // re-zeroed at the next compute step's level 0 and stay readable until then)
    pca_hall(mmax, ncylorder, sampT, subsamp, (double)c->CurTotal(), numbT, massT, covV, covM, pb);
  }
  if (pcavar && used && npca0 < (1 << 30) && pb.valid && tk_type != None &&
      (multistep == 0 || !self_consistent)) {
    std::vector<double> cosv, sinv;
    get_coefs(cosv, sinv);
    raw_cos = cosv; raw_sin = sinv;
    set_trimmed(pb, tk_type, snr, hexp, cosv, sinv);
    EXPHOST_CHECK(expb200_cylinder_set_coefficients(b200c, cosv.data(), sinv.data(), cylmass));
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

// EmpCylSL::pca_hall, PCAVAR branch (exputil/EmpCylSL.cc:4617-4631, 4684-4742, 4845-4884)
bool Cylinder::pca_hall(int mmax, int norder, unsigned sampT, bool subsamp, double nbodstot,
                        const std::vector<unsigned> &numbT, const std::vector<double> &massT,
                        const std::vector<double> &covV, const std::vector<double> &covM, PCA &pb)
{
  const size_t M1 = mmax + 1, N = norder;
  pb = PCA();
  pb.mmax = mmax; pb.norder = norder;
  for (auto v : numbT) pb.Tnumb += v;
  for (auto v : massT) pb.Tmass += v;
  if (pb.Tmass <= 0.0) return false;                               // no data
  pb.meanJK.assign(M1 * N, 0.0); pb.evalJK.assign(M1 * N, 0.0); pb.ratio.assign(M1 * N, 0.0);
  pb.covrJK.assign(M1 * N * N, 0.0); pb.evecJK.assign(M1 * N * N, 0.0);
  std::vector<double> sym(N * N), eval, evec;
  for (size_t mm = 0; mm < M1; mm++) {
    double *mean = &pb.meanJK[mm * N], *cov = &pb.covrJK[mm * N * N];
    for (unsigned T = 0; T < sampT; T++) {                         // data partitions
      if (massT[T] <= 0.0) continue;                               // skip empty partition
      const double *v = &covV[((size_t)T * M1 + mm) * N];
      const double *m = &covM[((size_t)T * M1 + mm) * N * N];
      for (size_t nn = 0; nn < N; nn++) {
        mean[nn] += v[nn] / massT[T] / sampT;
        for (size_t oo = 0; oo < N; oo++)
          cov[nn * N + oo] += subsamp ? v[nn] / massT[T] * v[oo] / massT[T] / sampT
                                      : m[nn * N + oo] / massT[T] / sampT;
      }
    }
    for (size_t nn = 0; nn < N; nn++)
      for (size_t oo = 0; oo < N; oo++) cov[nn * N + oo] -= mean[nn] * mean[oo];
    // the reference solves the (symmetric) matrix with Eigen::EigenSolver, whose eigenvalue order
    // is implementation-defined; here: symmetric Jacobi, eigenvalues descending
    for (size_t nn = 0; nn < N; nn++)
      for (size_t oo = 0; oo < N; oo++) sym[nn * N + oo] = 0.5 * (cov[nn * N + oo] + cov[oo * N + nn]);
    jacobi_eigen((int)N, sym, eval, evec);
    const size_t nfid = std::min<size_t>(4, N) - 1;                 // sign convention (:4735-4739)
    for (size_t j = 0; j < N; j++)
      if (evec[nfid * N + j] < 0.0)
        for (size_t k = 0; k < N; k++) evec[k * N + j] = -evec[k * N + j];
    std::copy(eval.begin(), eval.end(), &pb.evalJK[mm * N]);
    std::copy(evec.begin(), evec.end(), &pb.evecJK[mm * N * N]);
    for (size_t nn = 0; nn < N; nn++) {
      double dd = 0.0;                                             // projected mean
      for (size_t k = 0; k < N; k++) dd += evec[k * N + nn] * mean[k];
      double var = std::max(eval[nn], std::numeric_limits<double>::min());
      if (subsamp) var *= nbodstot / (double)sampT;                // CLT estimate of the N-particle variance
      pb.ratio[mm * N + nn] = var / (dd * dd) / nbodstot;
    }
  }
  pb.valid = true;
  return true;
}

// EmpCylSL::set_trimmed (exputil/EmpCylSL.cc:5141-5207) without the `rem` branch
void Cylinder::set_trimmed(const PCA &pb, TKType tk, double snr, double hexp, std::vector<double> &cosv,
                           std::vector<double> &sinv)
{
  if (!pb.valid || tk == None) return;
  const size_t M1 = pb.mmax + 1, N = pb.norder;
  std::vector<double> smth(N), dd(N), wrk(N);
  for (size_t mm = 0; mm < M1; mm++) {
    const double *evec = &pb.evecJK[mm * N * N];
    for (size_t i = 0; i < N; i++) {
      const double b = pb.ratio[mm * N + i];
      if (tk == Hall) smth[i] = 1.0 / (1.0 + std::pow(snr * b, hexp));
      else            smth[i] = 1.0 / b > snr ? 1.0 : 0.0;
    }
    auto apply = [&](double *c) {
      for (size_t nn = 0; nn < N; nn++) {                          // project to the decorrelated basis
        double t = 0.0;
        for (size_t k = 0; k < N; k++) t += evec[k * N + nn] * c[k];
        wrk[nn] = t * smth[nn];                                    // smooth
      }
      for (size_t k = 0; k < N; k++) {                             // deproject
        double t = 0.0;
        for (size_t nn = 0; nn < N; nn++) t += evec[k * N + nn] * wrk[nn];
        dd[k] = t;
      }
      std::copy(dd.begin(), dd.end(), c);
    };
    apply(&cosv[mm * N]);
    if (mm) apply(&sinv[mm * N]);
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
