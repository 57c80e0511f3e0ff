// Cylinder.cc -- see Cylinder.H
#include "Cylinder.H"

#include "global.H"

namespace exphost {

Cylinder::Cylinder(const EmpCylSLTables &t, int device, double rcylmax_, double acyl_, bool even_m,
                   int MMIN, int MLIM, int NMIN, int NLIM)
  : mmax(t.mmax), ncylorder(t.norder), rcylmax(rcylmax_), EVEN_M(even_m)
{
  id = "cylinder";
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

  uint64_t use1 = 0;
  double mass1 = 0.0;
  EXPHOST_CHECK(expb200_cylinder_determine_coefficients(b200c, c->b200, (int)mlevel, c->Adiabatic(),
                                                        nullptr, nullptr, &use1, &mass1));
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
