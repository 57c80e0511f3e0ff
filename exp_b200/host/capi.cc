// capi.cc -- see capi.h
#include "capi.h"

#include <complex>
#include <sstream>

#include <cstring>
#include <string>

#include "Cylinder.H"
#include "SphericalBasis.H"
#include "expand.H"

using namespace exphost;

namespace {

thread_local std::string g_err;
ComponentContainer *g_comp = nullptr;

Component *component_at(int index)
{
  if (!g_comp || index < 0 || index >= (int)g_comp->owned.size())
    throw GenericError("no such component", __FILE__, __LINE__);
  return g_comp->owned[index].get();
}

template <class F> int guarded(F &&f)
{
  try { f(); return 0; }
  catch (const GenericError &e) { g_err = e.what(); return e.errcode ? e.errcode : -1; }
  catch (const std::exception &e) { g_err = e.what(); return -1; }
}

} // namespace

extern "C" {

const char *exphost_last_error(void) { return g_err.c_str(); }

int exphost_reset(void)
{
  return guarded([] {
    delete g_comp;
    g_comp = new ComponentContainer();
    comp = g_comp;
    level_report.clear();
    pos_check.clear(); vel_check.clear();
    reset_globals();
  });
}

int exphost_set_params(double dtime_, int multistep_, double S, double D, double V, double A, double P,
                       int shift)
{
  return guarded([&] {
    if (multistep_ < 0 || multistep_ > 20) throw GenericError("multistep out of range", __FILE__, __LINE__);
    dtime = dtime_; multistep = (unsigned)multistep_;
    dynfracS = S; dynfracD = D; dynfracV = V; dynfracA = A; dynfracP = P; shiftlevl = shift;
  });
}

int exphost_set_comm(void *c, int id, int np)
{
  nccl_comm = c; myid = id; numprocs = np;
  return 0;
}

int exphost_add_component(const char *name, size_t n, const double *pos, const double *vel,
                          const double *mass, int device, int *index)
{
  return guarded([&] {
    if (!g_comp) throw GenericError("call exphost_reset first", __FILE__, __LINE__);
    auto c = std::make_unique<Component>(name, n, device);
    for (size_t i = 0; i < n; i++) {
      Particle &p = c->particles[i];
      p.mass = mass[i];
      for (int k = 0; k < 3; k++) { p.pos[k] = pos[3 * i + k]; p.vel[k] = vel ? vel[3 * i + k] : 0.0; }
    }
    c->ParticlesToCuda();
    g_comp->add(std::move(c));
    if (index) *index = (int)g_comp->owned.size() - 1;
  });
}

int exphost_component_options(int index, const double center[3], double rtrunc, int noswitch,
                              int freezeLev, int dtreset)
{
  return guarded([&] {
    Component *c = component_at(index);
    if (center) for (int k = 0; k < 3; k++) c->center[k] = center[k];
    c->rtrunc = rtrunc; c->noswitch = noswitch; c->freezeLev = freezeLev; c->dtreset = dtreset;
    c->update_center();
  });
}

int exphost_set_sphere(int index, const exphost_sphere_desc *d)
{
  return guarded([&] {
    Component *c = component_at(index);
    SLGridSphTables t;
    t.lmax = d->lmax; t.nmax = d->nmax; t.numr = d->numr; t.cmap = d->cmap;
    t.rmin = d->rmin; t.rmax = d->rmax; t.rmap = d->rmap; t.ev = d->ev; t.ef = d->ef; t.p0 = d->p0;
    auto f = std::make_shared<Sphere>(t, c->cudaDevice, d->scale, d->rmin_exp, d->rmax_exp, d->NO_L0,
                                      d->NO_L1, d->EVEN_L, d->EVEN_M, d->M0_only);
    f->self_consistent = d->self_consistent;
    c->set_force(f);
  });
}

int exphost_set_cylinder(int index, const exphost_cylinder_desc *d)
{
  return guarded([&] {
    Component *c = component_at(index);
    EmpCylSLTables t;
    t.mmax = d->mmax; t.norder = d->norder; t.numx = d->numx; t.numy = d->numy;
    t.cmapr = d->cmapr; t.cmapz = d->cmapz; t.ascale = d->ascale; t.hscale = d->hscale;
    t.rmin = d->rmin; t.rmax = d->rmax;
    t.potC = d->potC; t.potS = d->potS; t.rforceC = d->rforceC; t.rforceS = d->rforceS;
    t.zforceC = d->zforceC; t.zforceS = d->zforceS;
    auto f = std::make_shared<Cylinder>(t, c->cudaDevice, d->rcylmax, d->acyl, d->EVEN_M, d->MMIN,
                                        d->MLIM, d->NMIN, d->NLIM);
    f->self_consistent = d->self_consistent;
    if (d->body && d->orig) f->set_orientation(d->body, d->orig);
    c->set_force(f);
  });
}

int exphost_component_global(int index, uint64_t indx_base, uint64_t nbodies_tot)
{
  return guarded([&] {
    Component *c = component_at(index);
    c->indx_base = indx_base;
    c->nbodies_tot = nbodies_tot;
  });
}

int exphost_sphere_pca_options(int index, int npca, int npca0, int nint, int pcavar, int pcaeof,
                               int samplesz, int subsamp, int tk_type, double hexp, double snr,
                               double tksmooth, double tkcum)
{
  return guarded([&] {
    auto *s = dynamic_cast<SphericalBasis *>(component_at(index)->force.get());
    if (!s) throw GenericError("component force is not sphereSL", __FILE__, __LINE__);
    AxisymmetricBasis &p = s->pca;
    p.npca = npca; p.npca0 = npca0; p.nint = nint;
    p.pcavar = pcavar || nint > 0;                    // src/AxisymmetricBasis.cc:86-89
    p.pcaeof = pcaeof;
    p.defSampT = (unsigned)samplesz;
    p.subsamp = subsamp;
    p.tk_type = (AxisymmetricBasis::TKType)tk_type;
    p.hexp = hexp; p.snr = snr; p.tksmooth = tksmooth; p.tkcum = tkcum;
  });
}

int exphost_sphere_get_covariance(int index, int *sampT, double *massT, unsigned *countT, double *coefT,
                                  double *coefM, double *tvar, double *b_Hall, double *s_Hall,
                                  double *weight, double *muse)
{
  return guarded([&] {
    auto *s = dynamic_cast<SphericalBasis *>(component_at(index)->force.get());
    if (!s) throw GenericError("component force is not sphereSL", __FILE__, __LINE__);
    const AxisymmetricBasis &p = s->pca;
    if (sampT) *sampT = p.massT.empty() ? 0 : (int)p.sampT;
    if (muse) *muse = p.muse;
    if (massT) std::copy(p.massT.begin(), p.massT.end(), massT);
    if (countT) std::copy(p.countT.begin(), p.countT.end(), countT);
    if (coefT) memcpy(coefT, p.expcoefT.data(), p.expcoefT.size() * sizeof(std::complex<double>));
    if (coefM) std::copy(p.expcoefM.begin(), p.expcoefM.end(), coefM);
    if (tvar) std::copy(p.tvar.begin(), p.tvar.end(), tvar);
    if (b_Hall) std::copy(p.b_Hall.begin(), p.b_Hall.end(), b_Hall);
    if (s_Hall) std::copy(p.s_Hall.begin(), p.s_Hall.end(), s_Hall);
    if (weight) std::copy(p.weight.begin(), p.weight.end(), weight);
  });
}

int exphost_cylinder_cov_options(int index, int nint, int npca, int npca0, int pcavar, int pcaeof, int samplesz)
{
  return guarded([&] {
    auto *s = dynamic_cast<Cylinder *>(component_at(index)->force.get());
    if (!s) throw GenericError("component force is not cylinder", __FILE__, __LINE__);
    s->nint = nint; s->npca = npca; s->npca0 = npca0; s->pcavar = pcavar; s->pcaeof = pcaeof;
    s->defSampT = (unsigned)samplesz;
  });
}

int exphost_cylinder_get_covariance(int index, int *sampT, unsigned *numbT, double *massT, double *VC,
                                    double *MV, double *covV, double *covM, double *tvar)
{
  return guarded([&] {
    auto *s = dynamic_cast<Cylinder *>(component_at(index)->force.get());
    if (!s) throw GenericError("component force is not cylinder", __FILE__, __LINE__);
    if (sampT) *sampT = s->massT.empty() ? 0 : (int)s->sampT;
    if (numbT) std::copy(s->numbT.begin(), s->numbT.end(), numbT);
    if (massT) std::copy(s->massT.begin(), s->massT.end(), massT);
    if (VC) std::copy(s->VC.begin(), s->VC.end(), VC);
    if (MV) std::copy(s->MV.begin(), s->MV.end(), MV);
    if (covV) std::copy(s->covV.begin(), s->covV.end(), covV);
    if (covM) std::copy(s->covM.begin(), s->covM.end(), covM);
    if (tvar) std::copy(s->tvar.begin(), s->tvar.end(), tvar);
  });
}

int exphost_cylinder_pca_options(int index, int tk_type, double snr, double hexp, int subsamp)
{
  return guarded([&] {
    auto *s = dynamic_cast<Cylinder *>(component_at(index)->force.get());
    if (!s) throw GenericError("component force is not cylinder", __FILE__, __LINE__);
    if (tk_type < 0 || tk_type > 2) throw GenericError("tk_type: 0 Hall, 1 Truncate, 2 None", __FILE__, __LINE__);
    s->tk_type = (Cylinder::TKType)tk_type; s->snr = snr; s->hexp = hexp; s->subsamp = subsamp != 0;
  });
}

static void copy_pca(const Cylinder::PCA &pb, double *meanJK, double *evalJK, double *ratio, double *evecJK)
{
  if (meanJK) std::copy(pb.meanJK.begin(), pb.meanJK.end(), meanJK);
  if (evalJK) std::copy(pb.evalJK.begin(), pb.evalJK.end(), evalJK);
  if (ratio) std::copy(pb.ratio.begin(), pb.ratio.end(), ratio);
  if (evecJK) std::copy(pb.evecJK.begin(), pb.evecJK.end(), evecJK);
}

int exphost_cylinder_get_pca(int index, int *valid, double *meanJK, double *evalJK, double *ratio, double *evecJK,
                             double *raw_cos, double *raw_sin)
{
  return guarded([&] {
    auto *s = dynamic_cast<Cylinder *>(component_at(index)->force.get());
    if (!s) throw GenericError("component force is not cylinder", __FILE__, __LINE__);
    if (valid) *valid = s->pb.valid ? 1 : 0;
    if (s->pb.valid) copy_pca(s->pb, meanJK, evalJK, ratio, evecJK);
    if (raw_cos) std::copy(s->raw_cos.begin(), s->raw_cos.end(), raw_cos);
    if (raw_sin) std::copy(s->raw_sin.begin(), s->raw_sin.end(), raw_sin);
  });
}

int exphost_cyl_pca_hall(int mmax, int norder, int sampT, int subsamp, double nbodstot, int tk_type, double snr,
                         double hexp, const unsigned *numbT, const double *massT, const double *covV,
                         const double *covM, double *meanJK, double *evalJK, double *ratio, double *evecJK,
                         double *cosv, double *sinv)
{
  return guarded([&] {
    if (mmax < 0 || norder < 1 || sampT < 1 || tk_type < 0 || tk_type > 2)
      throw GenericError("exphost_cyl_pca_hall: bad argument", __FILE__, __LINE__);
    const size_t M1 = mmax + 1, N = norder, T = sampT;
    std::vector<unsigned> nb(numbT, numbT + T);
    std::vector<double> ms(massT, massT + T), v(covV, covV + T * M1 * N), m(covM, covM + T * M1 * N * N);
    Cylinder::PCA pb;
    if (!Cylinder::pca_hall(mmax, norder, (unsigned)sampT, subsamp != 0, nbodstot, nb, ms, v, m, pb))
      throw GenericError("exphost_cyl_pca_hall: no mass in the partitions", __FILE__, __LINE__);
    copy_pca(pb, meanJK, evalJK, ratio, evecJK);
    if (cosv && sinv) {
      std::vector<double> c(cosv, cosv + M1 * N), s(sinv, sinv + M1 * N);
      Cylinder::set_trimmed(pb, (Cylinder::TKType)tk_type, snr, hexp, c, s);
      std::copy(c.begin(), c.end(), cosv);
      std::copy(s.begin(), s.end(), sinv);
    }
  });
}

int exphost_pca_hall(int Lmax, int nmax, int sampT, int used, int subsamp, int tk_type, double hexp,
                     double snr, double tksmooth, double tkcum, const double *massT, const double *coefT,
                     const double *coefM, double *b_Hall, double *s_Hall, double *weight, double *evalJK,
                     double *evec, double *expcoef)
{
  return guarded([&] {
    AxisymmetricBasis p;
    p.Lmax = Lmax; p.nmax = nmax; p.pcavar = true; p.subsamp = subsamp;
    p.tk_type = (AxisymmetricBasis::TKType)tk_type;
    p.hexp = hexp; p.snr = snr; p.tksmooth = tksmooth; p.tkcum = tkcum;
    p.defSampT = (unsigned)sampT;
    p.allocate(0);
    const size_t L = (size_t)p.ldim();
    double mu = 0.0;
    for (int T = 0; T < sampT; T++) { p.massT[T] = massT[T]; mu += massT[T]; }
    p.muse = mu;
    for (size_t k = 0; k < (size_t)sampT * L * nmax; k++) p.expcoefT[k] = std::complex<double>(coefT[2 * k], coefT[2 * k + 1]);
    if (coefM) std::copy(coefM, coefM + (size_t)sampT * L * nmax * nmax, p.expcoefM.begin());
    p.pca_hall(true, used, nullptr);
    if (b_Hall) std::copy(p.b_Hall.begin(), p.b_Hall.end(), b_Hall);
    if (s_Hall) std::copy(p.s_Hall.begin(), p.s_Hall.end(), s_Hall);
    if (weight) std::copy(p.weight.begin(), p.weight.end(), weight);
    if (evalJK) std::copy(p.evalJK.begin(), p.evalJK.end(), evalJK);
    if (evec) std::copy(p.evec.begin(), p.evec.end(), evec);
    if (expcoef) p.smooth(expcoef);
  });
}

int exphost_add_interaction(int from, int on)
{
  return guarded([&] { g_comp->add_interaction(component_at(from), component_at(on)); });
}

int exphost_begin_run(void) { return guarded([] { begin_run(); }); }

int exphost_do_steps(int nsteps)
{
  return guarded([&] { for (int i = 0; i < nsteps; i++) do_step(this_step); });
}

int exphost_do_steps_timed(int nsteps, double *ms)
{
  return guarded([&] {
    Component *c0 = component_at(0);
    for (auto &c : g_comp->owned) EXPHOST_CHECK(expb200_component_sync(c->b200));
    EXPHOST_CHECK(expb200_component_event_record(c0->b200, 6));
    for (int i = 0; i < nsteps; i++) do_step(this_step);
    for (auto &c : g_comp->owned) EXPHOST_CHECK(expb200_component_sync(c->b200));
    EXPHOST_CHECK(expb200_component_event_record(c0->b200, 7));
    EXPHOST_CHECK(expb200_component_sync(c0->b200));
    EXPHOST_CHECK(expb200_component_event_elapsed(c0->b200, 6, 7, ms));
  });
}

int exphost_get_particles(int index, double *pos, double *vel, double *acc, double *pot,
                          unsigned *level, float *dtreq)
{
  return guarded([&] {
    Component *c = component_at(index);
    c->CudaToParticles();
    const size_t n = c->Number();
    for (size_t i = 0; i < n; i++) {
      const Particle &p = c->particles[i];
      for (int k = 0; k < 3; k++) {
        if (pos) pos[3 * i + k] = p.pos[k];
        if (vel) vel[3 * i + k] = p.vel[k];
        if (acc) acc[3 * i + k] = p.acc[k];
      }
      if (pot) pot[i] = p.pot;
      if (level) level[i] = p.level;
      if (dtreq) dtreq[i] = p.dtreq;
    }
  });
}

int exphost_get_sphere_coefs(int index, int which, int M, double *coef)
{
  return guarded([&] {
    auto *s = dynamic_cast<SphericalBasis *>(component_at(index)->force.get());
    if (!s) throw GenericError("component force is not sphereSL", __FILE__, __LINE__);
    auto v = which == 3 ? s->get_coefs() : s->get_level_coefs(which, M);
    std::memcpy(coef, v.data(), v.size() * sizeof(double));
  });
}

int exphost_get_cylinder_coefs(int index, double *cosv, double *sinv, double *cylmass)
{
  return guarded([&] {
    auto *s = dynamic_cast<Cylinder *>(component_at(index)->force.get());
    if (!s) throw GenericError("component force is not cylinder", __FILE__, __LINE__);
    std::vector<double> a, b;
    s->get_coefs(a, b);
    std::memcpy(cosv, a.data(), a.size() * sizeof(double));
    std::memcpy(sinv, b.data(), b.size() * sizeof(double));
    if (cylmass) *cylmass = s->get_cylmass();
  });
}

int exphost_level_report(int index, uint64_t *examined, uint64_t *switched, uint64_t *offlo,
                         uint64_t *offhi, uint64_t *counts)
{
  return guarded([&] {
    Component *c = component_at(index);
    const LevelReport &r = level_report[c];
    if (examined) *examined = r.examined;
    if (switched) *switched = r.switched;
    if (offlo) *offlo = r.offlo;
    if (offhi) *offhi = r.offhi;
    if (counts) {
      auto v = c->level_counts();
      for (size_t i = 0; i < v.size(); i++) counts[i] = v[i];
    }
  });
}

int exphost_sort_counters(int index, uint64_t counters[4])
{
  return guarded([&] {
    Component *c = component_at(index);
    EXPHOST_CHECK(expb200_component_sort_counters(c->b200, counters));
  });
}

int exphost_multistep_tables(int ms, int *mintvl_, int *mfirst_, int *dL, int *dN)
{
  return guarded([&] {
    if (ms < 0 || ms > 20) throw GenericError("multistep out of range", __FILE__, __LINE__);
    const unsigned keep = multistep;
    multistep = (unsigned)ms;
    initialize_multistep();
    for (int i = 0; i <= ms; i++) mintvl_[i] = mintvl[i];
    for (int i = 0; i <= Mstep; i++) mfirst_[i] = mfirst[i];
    for (int l = 0; l <= ms; l++)
      for (int n = 0; n < Mstep; n++) { dL[l * Mstep + n] = dstepL[l][n]; dN[l * Mstep + n] = dstepN[l][n]; }
    multistep = keep;
  });
}

// Serialises the component's coefficients with the force's dump_coefs (kind 0: NewCoefs
// stream, 1: old-style stream [sphere only], 2: the CoefStruct complex array as (re, im)
// doubles).  Returns the number of bytes (buf may be NULL to query the size).
long exphost_dump_coefs(int index, int kind, char *buf, long cap)
{
  long nbytes = -1;
  int rc = guarded([&] {
    PotAccel *f = component_at(index)->force.get();
    std::ostringstream out;
    if (auto *s = dynamic_cast<SphericalBasis *>(f)) {
      if (kind == 2) {
        SphStruct cs = s->coef_struct();
        out.write(reinterpret_cast<const char *>(cs.coefs.data()), cs.coefs.size() * sizeof(std::complex<double>));
      } else s->dump_coefs(out, kind == 0);
    } else if (auto *cy = dynamic_cast<Cylinder *>(f)) {
      if (kind == 2) {
        CylStruct cs = cy->coef_struct();
        out.write(reinterpret_cast<const char *>(cs.coefs.data()), cs.coefs.size() * sizeof(std::complex<double>));
      } else cy->dump_coefs(out);
    } else throw GenericError("component has no basis force", __FILE__, __LINE__);
    const std::string str = out.str();
    nbytes = (long)str.size();
    if (buf) {
      if (cap < nbytes) throw GenericError("buffer too small", __FILE__, __LINE__);
      std::memcpy(buf, str.data(), str.size());
    }
  });
  return rc ? rc : nbytes;
}

int exphost_get_com(int index, double *out, double *lev)
{
  return guarded([&] {
    Component *c = component_at(index);
    if (out) {
      out[0] = c->mtot;
      for (int k = 0; k < 3; k++) { out[1 + k] = c->com[k]; out[4 + k] = c->cov[k]; out[7 + k] = c->coa[k]; }
    }
    if (lev)
      for (size_t mm = 0; mm < c->com_mas.size(); mm++) {
        lev[10 * mm] = c->com_mas[mm];
        for (int k = 0; k < 3; k++) {
          lev[10 * mm + 1 + k] = c->com_lev[3 * mm + k];
          lev[10 * mm + 4 + k] = c->cov_lev[3 * mm + k];
          lev[10 * mm + 7 + k] = c->coa_lev[3 * mm + k];
        }
      }
  });
}

int exphost_state(double *tnow_, int *this_step_, double *pc, double *vc)
{
  if (tnow_) *tnow_ = tnow;
  if (this_step_) *this_step_ = this_step;
  // pc, vc hold multistep + 1 entries (the caller sizes them from the multistep it configured);
  // the CHK_STEP sums exist only for multistep > 0 (step.cc)
  for (size_t i = 0; i <= multistep; i++) {
    if (pc) pc[i] = i < pos_check.size() ? pos_check[i] : 0.0;
    if (vc) vc[i] = i < vel_check.size() ? vel_check[i] : 0.0;
  }
  return 0;
}

} // extern "C"
