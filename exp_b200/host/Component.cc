// Component.cc -- see Component.H
#include "Component.H"

#include <cmath>

#include "global.H"

namespace exphost {

Component::Component(const std::string &nm, size_t nbodies, int device)
  : name(nm), particles(nbodies), cudaDevice(device)
{
  for (size_t i = 0; i < nbodies; i++) particles[i].indx = i + 1;    // reference indices start at 1
  EXPHOST_CHECK(expb200_component_create(device, nbodies, &b200));
}

Component::~Component()
{
  if (b200) expb200_component_destroy(b200);
}

double Component::Adiabatic() const
{
  if (!adiabatic) return 1.0;
  return 0.25 * (1.0 + std::erf((tnow - ton) / twid)) * (1.0 + std::erf((toff - tnow) / twid));
}

void Component::update_center()
{
  double ctr[3] = {com0[0] + center[0], com0[1] + center[1], com0[2] + center[2]};
  EXPHOST_CHECK(expb200_component_set_center(b200, ctr, rtrunc));
}

// Flatten the Particle fields (include/Particle.H:31-67) into the ABI's arrays and upload
void Component::ParticlesToCuda()
{
  const size_t n = particles.size();
  h_pos.resize(3 * n); h_vel.resize(3 * n); h_mass.resize(n); h_potext.resize(n);
  h_lev.resize(n); h_scale.resize(n);
  for (size_t i = 0; i < n; i++) {
    const Particle &p = particles[i];
    for (int k = 0; k < 3; k++) { h_pos[3 * i + k] = p.pos[k]; h_vel[3 * i + k] = p.vel[k]; }
    h_mass[i] = p.mass; h_potext[i] = p.potext; h_lev[i] = p.level; h_scale[i] = p.scale;
  }
  EXPHOST_CHECK(expb200_component_upload(b200, n, h_pos.data(), h_vel.data(), h_mass.data(),
                                         h_lev.data(), h_scale.data(), h_potext.data()));
  update_center();
}

void Component::CudaToParticles()
{
  const size_t n = particles.size();
  h_pos.resize(3 * n); h_vel.resize(3 * n); h_acc.resize(3 * n); h_pot.resize(n);
  h_lev.resize(n); h_dtreq.resize(n);
  EXPHOST_CHECK(expb200_component_download(b200, h_pos.data(), h_vel.data(), h_acc.data(),
                                           h_pot.data(), h_lev.data(), h_dtreq.data()));
  for (size_t i = 0; i < n; i++) {
    Particle &p = particles[i];
    for (int k = 0; k < 3; k++) {
      p.pos[k] = h_pos[3 * i + k]; p.vel[k] = h_vel[3 * i + k]; p.acc[k] = h_acc[3 * i + k];
    }
    p.pot = h_pot[i]; p.level = h_lev[i]; p.dtreq = h_dtreq[i];
  }
}

void Component::ZeroPotAccel(int minlev)
{
  EXPHOST_CHECK(expb200_component_zero_pot_acc(b200, minlev));
}

void Component::fix_positions(unsigned mlevel)
{
  const size_t nlev = multistep + 1;
  if (com_mas.size() != nlev) {
    com_lev.assign(3 * nlev, 0.0); cov_lev.assign(3 * nlev, 0.0); coa_lev.assign(3 * nlev, 0.0);
    com_mas.assign(nlev, 0.0);
    mlevel = 0;                           // no per-level history yet: sum every level
  }
  mtot = 0.0;
  for (int k = 0; k < dim; k++) com[k] = cov[k] = coa[k] = 0.0;

  // rows M >= mlevel: {com_mas, com_lev, cov_lev, coa_lev}, already summed over ranks
  std::vector<double> sums(10 * nlev, 0.0);
  EXPHOST_CHECK(expb200_component_fix_positions(b200, (int)mlevel, (int)multistep, dim, sums.data(), nccl_comm));
  for (size_t mm = mlevel; mm < nlev; mm++) {
    com_mas[mm] = sums[10 * mm];
    for (int k = 0; k < 3; k++) {
      com_lev[3 * mm + k] = sums[10 * mm + 1 + k];
      cov_lev[3 * mm + k] = sums[10 * mm + 4 + k];
      coa_lev[3 * mm + k] = sums[10 * mm + 7 + k];
    }
  }
  // sum levels (:3484-3497); the rank sum was done per level on the device
  for (size_t mm = 0; mm < nlev; mm++) {
    for (int k = 0; k < 3; k++) {
      com[k] += com_lev[3 * mm + k];
      cov[k] += cov_lev[3 * mm + k];
      coa[k] += coa_lev[3 * mm + k];
    }
    mtot += com_mas[mm];
  }
  if (mtot > 0.0) {
    for (int k = 0; k < dim; k++) com[k] /= mtot;
    for (int k = 0; k < dim; k++) cov[k] /= mtot;
    for (int k = 0; k < dim; k++) coa[k] /= mtot;
  }
  if (com_system) {                       // :3555-3563 (consp is not mirrored)
    for (int k = 0; k < dim; k++) com0[k] = com[k];
    for (int k = 0; k < dim; k++) cov0[k] = cov[k];
    for (int k = 0; k < dim; k++) acc0[k] = coa[k];
    update_center();
  } else {
    for (int k = 0; k < dim; k++) acc0[k] = 0.0;
  }
}

std::vector<unsigned long long> Component::level_counts() const
{
  std::vector<uint64_t> c(multistep + 1, 0);
  EXPHOST_CHECK(expb200_component_level_counts(b200, (int)multistep, c.data()));
  return std::vector<unsigned long long>(c.begin(), c.end());
}

} // namespace exphost
