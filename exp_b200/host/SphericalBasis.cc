// SphericalBasis.cc -- see SphericalBasis.H
#include "SphericalBasis.H"

#include <limits>

#include "global.H"

namespace exphost {

SphericalBasis::SphericalBasis(const SLGridSphTables &t, int device, double scale_, double rmin_,
                               double rmax_, bool no_l0, bool no_l1, bool even_l, bool even_m,
                               bool m0_only)
  : Lmax(t.lmax), nmax(t.nmax), scale(scale_), NO_L0(no_l0), NO_L1(no_l1), EVEN_L(even_l),
    EVEN_M(even_m), M0_only(m0_only)
{
  id = "SphereSL";             // src/Sphere.cc:31
  // rmin/rmax default to the table's range, as the Sphere constructor does (src/Sphere.cc:65-66)
  rmin = rmin_ < 0.0 ? t.rmin : rmin_;
  rmax = rmax_ < 0.0 ? t.rmax : rmax_;
  expb200_sphere_config cfg = {t.lmax, t.nmax, t.numr, t.cmap, t.rmin, t.rmax, t.rmap, scale, rmin, rmax,
                               NO_L0, NO_L1, EVEN_L, EVEN_M, M0_only, (int)multistep};
  EXPHOST_CHECK(expb200_sphere_create(device, &cfg, t.ev, t.ef, t.p0, &b200s));
  pca.Lmax = Lmax;
  pca.nmax = nmax;
}

SphericalBasis::~SphericalBasis()
{
  if (b200s) expb200_sphere_destroy(b200s);
}

void SphericalBasis::determine_coefficients(Component *c)
{
  // src/SphericalBasis.cc:694: a fixed (not self-consistent) basis is computed once
  if (!self_consistent && !firstime_coef && !initializing) return;

  // covariance gating (src/SphericalBasis.cc:696-781)
  pca.update_compute(mstep, this_step);
  if (pca.compute) {
    pca.allocate(c->CurTotal());
    if (mlevel == 0) pca.zero_accumulation();
  }

  uint64_t use1 = 0;
  EXPHOST_CHECK(expb200_sphere_determine_coefficients(b200s, c->b200, (int)mlevel, c->Adiabatic(),
                                                      nullptr, &use1));
  if (pca.compute) {
    // the `compute` branch of determine_coefficients_thread (:501-509, 525-545, 563-581) for
    // this level's particles; bins by the particle's global index
    const int T = (int)pca.sampT, L = pca.ldim(), N = nmax;
    if (pca.pcavar) {
      std::vector<unsigned> cnt(T);
      std::vector<double> mass(T), cT((size_t)T * L * N * 2), cM((size_t)T * L * N * N);
      EXPHOST_CHECK(expb200_sphere_covariance(b200s, c->b200, (int)mlevel, c->Adiabatic(), T, c->indx_base,
                                              cnt.data(), mass.data(), cT.data(), cM.data()));
      pca.add_level(cnt.data(), mass.data(), cT.data(), cM.data());
    }
    if (pca.pcaeof) {
      std::vector<double> tv((size_t)L * N * N);
      EXPHOST_CHECK(expb200_sphere_covariance_eof(b200s, c->b200, (int)mlevel, c->Adiabatic(), tv.data()));
      pca.add_tvar(tv.data());
    }
  }
  // the reference's (Lmax+1)^2 host MPI_Allreduce calls (:864-903) -> one device all-reduce
  if (nccl_comm) {
    double *d = nullptr; void *st = nullptr;
    EXPHOST_CHECK(expb200_sphere_coef_device(b200s, &d, &st));
    EXPHOST_CHECK(expb200_allreduce_sum(nccl_comm, d, ncoef(), st));
  }
  if (multistep) {
    // swap expcoefL <-> expcoefN[mlevel], store the new vector as expcoefN[mlevel] (:785-788)
    EXPHOST_CHECK(expb200_sphere_multistep_commit(b200s, (int)mlevel));
  } else {
    EXPHOST_CHECK(expb200_sphere_set_coefficients(b200s, nullptr));
  }
  if (multistep == 0 || tnow == resetT) used += (int)use1;          // :905-907

  // last level (:909-931): gather the subsample arrays over ranks, then PCA / Hall smoothing of
  // the installed coefficients
  if (mlevel == multistep) {
    if (pca.compute) pca.gather(nccl_comm);
    if (pca.npca0 < std::numeric_limits<int>::max()) {
      int used_all = pca.used_cov;
      if (nccl_comm) { double u = used_all; EXPHOST_CHECK(expb200_allreduce_sum_host(nccl_comm, &u, 1)); used_all = (int)u; }
      pca.pca_hall(pca.compute, used_all, nullptr);
      if (pca.pcavar && pca.tk_type != AxisymmetricBasis::None && !pca.evec.empty()) {
        if (multistep) compute_multistep_coefficients();
        std::vector<double> ex = get_coefs();
        pca.smooth(ex.data());
        EXPHOST_CHECK(expb200_sphere_set_coefficients(b200s, ex.data()));
      }
    }
  }
  firstime_coef = false;
}

void SphericalBasis::compute_multistep_coefficients()
{
  EXPHOST_CHECK(expb200_sphere_compute_multistep_coefficients(b200s, mdrft));
}

void SphericalBasis::get_acceleration_and_potential(Component *cC)
{
  if (!use_external) {
    // (as in the reference, a multistep run re-forms expcoef from the level tables here, which
    // discards a PCA-smoothed table installed by determine_coefficients: src/SphericalBasis.cc:1690-1695)
    if (multistep && (self_consistent || initializing)) compute_multistep_coefficients();
  }
  EXPHOST_CHECK(expb200_sphere_determine_acceleration(b200s, cC->b200, (int)mlevel));
}

void SphericalBasis::RegisterComponent()
{
  EXPHOST_CHECK(expb200_sphere_set_owner(b200s, component ? component->b200 : nullptr));
}

void SphericalBasis::multistep_reset()
{
  used = 0;
  resetT = tnow;
}

void SphericalBasis::multistep_update(Component *c)
{
  if (!self_consistent) return;
  // differences stay on the device; with several ranks they are all-reduced in _finish
  EXPHOST_CHECK(expb200_sphere_multistep_update(b200s, c->b200, c->Adiabatic(), 0));
  pending_update = true;
}

void SphericalBasis::multistep_update_finish()
{
  if (!pending_update) return;
  if (nccl_comm) {
    double *d = nullptr; size_t n = 0; void *st = nullptr;
    EXPHOST_CHECK(expb200_sphere_differ_device(b200s, &d, &n, &st));
    EXPHOST_CHECK(expb200_allreduce_sum(nccl_comm, d, n, st));
  }
  EXPHOST_CHECK(expb200_sphere_multistep_update_finish(b200s));
  pending_update = false;
}

std::vector<double> SphericalBasis::get_coefs()
{
  std::vector<double> v(ncoef());
  EXPHOST_CHECK(expb200_sphere_get_level_coefficients(b200s, 3, 0, v.data()));
  return v;
}

std::vector<double> SphericalBasis::get_level_coefs(int which, int M)
{
  std::vector<double> v(ncoef());
  EXPHOST_CHECK(expb200_sphere_get_level_coefficients(b200s, which, M, v.data()));
  return v;
}

} // namespace exphost
