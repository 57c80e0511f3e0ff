// oracle/ref_path_wrap.cc -- TEST INFRASTRUCTURE ONLY.
//
// C entry points over the reference's OWN hot-path member functions.  The bodies are not
// in this file: oracle/ref_extract.py cuts them from /root/reference into
// oracle/_ref/gen/*.inc at build time (git-ignored), and they are compiled here unmodified
// against the shim declarations of ref_shim/pathshim.H.  The result,
// oracle/_ref/libexpref_path.so, is what tests/test_oracle_ref.py pins the restatement
// (oracle/exp_oracle.cc, cyl_oracle.cc) against, bit for bit where the operation order is
// the same.  Array layouts are those of oracle/exp_oracle.h.
#include "ref_shim/pathshim.H"
#include <array>
#include <cstring>

// ---- globals the bodies name ------------------------------------------------------------
int nthrds = 1, myid = 0, VERBOSE = 0;
pthread_mutex_t io_lock = PTHREAD_MUTEX_INITIALIZER;
unsigned multistep = 0;
int this_step = 1, mdrft = 0, mstep = 0, Mstep = 1, shiftlevl = 0;
double dtime = 0.0, dynfracS = 1.0, dynfracD = 1.0e32, dynfracV = 0.01, dynfracA = 0.03, dynfracP = 0.05;
std::vector<int> mfirst;
std::vector<std::vector<int>> dstepL, dstepN;
double tnow = 0.0;
ComponentContainer *comp = nullptr;
bool EmpCylSL::enforce_limits = false;          /* exputil/EmpCylSL.cc:56 */

double factrl(int n) { return n <= 1 ? 1.0 : n * factrl(n - 1); }

// statics of src/multistep.cc:38-47 named by adjust_multistep_level_thread
static std::map<Component *, std::vector<unsigned>> offlo1, offhi1;
static std::vector<double> mindt1, maxdt1, adjtm1, adjtm2;
static std::vector<unsigned> numsw, numtt;
static const int mdtDim = 6;
static std::vector<std::vector<std::vector<unsigned>>> tmdt;

// ---- the reference's bodies ----------------------------------------------------------------
#include "_ref/gen/basis_mineps.inc"
#include "_ref/gen/basis_recursions.inc"

#include "_ref/gen/slgrid_r_to_xi.inc"
#include "_ref/gen/slgrid_xi_to_r.inc"
#include "_ref/gen/slgrid_d_xi_to_r.inc"
#include "_ref/gen/slgrid_get_pot_mat.inc"
#include "_ref/gen/slgrid_get_dens_mat.inc"
#include "_ref/gen/slgrid_get_force_mat.inc"
void SLGridSph::init_grid()
{
  xi.resize(numr);
#include "_ref/gen/slgrid_init_grid.inc"
  for (int i = 0; i < numr; i++) xi[i] = xmin + dxi * i;      /* exputil/SLGridMP2.cc:1375-1376 */
}

#include "_ref/gen/sphere_get_dpotl.inc"
#include "_ref/gen/sphere_get_potl.inc"
#include "_ref/gen/sph_coef_thread.inc"
#include "_ref/gen/sph_accel_thread.inc"
#include "_ref/gen/sph_get_pot_coefs_safe.inc"
void SphericalBasis::setup(int L, int N, int nthr)
{
  Lmax = L; nmax = N; nthrds = nthr;
#include "_ref/gen/sph_factorial.inc"
  const int nrow = (L + 1) * (L + 1), ldim = (L + 1) * (L + 2) / 2;
  use.assign(nthr, 0);
  muse1.assign(nthr, 0.0);
  legs.assign(nthr, Eigen::MatrixXd(L + 1, L + 1));
  dlegs = legs;
  potd.assign(nthr, Eigen::MatrixXd(L + 1, N));
  dpot = potd;
  cosm.assign(nthr, Eigen::VectorXd(L + 1));
  sinm = cosm;
  expcoef0.resize(nthr);
  for (auto &v : expcoef0) {
    v.resize(nrow);
    for (auto &p : v) p = std::make_shared<Eigen::VectorXd>(N);
  }
  expcoef.resize(nrow);
  for (auto &p : expcoef) p = std::make_shared<Eigen::VectorXd>(N);
  (void)ldim;
}

#include "_ref/gen/sph_multistep_update.inc"
#include "_ref/gen/sph_compute_multistep_coefficients.inc"

#include "_ref/gen/cyl_r_to_xi.inc"
#include "_ref/gen/cyl_z_to_y.inc"
#include "_ref/gen/cyl_get_pot.inc"
#include "_ref/gen/cyl_accumulate.inc"
#include "_ref/gen/cyl_accumulated_eval.inc"

#include "_ref/gen/incpos_thread.inc"
#include "_ref/gen/incvel_thread.inc"
#include "_ref/gen/multistep_level_thread.inc"

// ---- C entry points ----------------------------------------------------------------------
namespace {

struct SphHandle {
  SLGridSph grid;
  Sphere sph;
  Component comp, other;
};

void fill_component(Component &c, long n, const double *pos, const double *vel, const double *acc,
                    const double *pot, const double *potext, const double *mass, const float *scale,
                    const unsigned *level, const float *dtreq, const double *center, double rtrunc)
{
  c.store.resize(n);
  for (long i = 0; i < n; i++) {
    Particle &p = c.store[i];
    for (int k = 0; k < 3; k++) {
      p.pos[k] = pos[3 * i + k];
      p.vel[k] = vel ? vel[3 * i + k] : 0.0;
      p.acc[k] = acc ? acc[3 * i + k] : 0.0;
    }
    p.pot = pot ? pot[i] : 0.0;
    p.potext = potext ? potext[i] : 0.0;
    p.mass = mass ? mass[i] : 0.0;
    p.scale = scale ? scale[i] : -1.0f;
    p.level = level ? level[i] : 0u;
    p.dtreq = dtreq ? dtreq[i] : 0.0f;
    p.indx = (unsigned long)i;
  }
  // "center" is com0 + center (what Pos(.., Local|Centered) subtracts); carried in center[]
  for (int k = 0; k < 3; k++) { c.center[k] = center ? center[k] : 0.0; c.com0[k] = 0.0; }
  c.rtrunc = rtrunc;
}

} // namespace

extern "C" {

void *ref_sph_create(int lmax, int nmax, int numr, int cmap, double rmin, double rmax, double rmap,
                     const double *ev, const double *ef, const double *p0, const double *d0)
{
  SphHandle *h = new SphHandle;
  SLGridSph &g = h->grid;
  g.lmax = lmax; g.nmax = nmax; g.numr = numr; g.cmap = cmap;
  g.rmin = rmin; g.rmax = rmax; g.rmap = rmap;
  g.init_grid();
  g.p0.resize(numr); g.d0.resize(numr);
  for (int i = 0; i < numr; i++) { g.p0[i] = p0[i]; g.d0[i] = d0 ? d0[i] : 0.0; }
  g.table = std::shared_ptr<TableSph[]>(new TableSph[lmax + 1]);
  for (int l = 0; l <= lmax; l++) {
    TableSph &t = g.table[l];
    t.l = l;
    t.ev.resize(nmax);
    t.ef.resize(nmax, numr);
    for (int n = 0; n < nmax; n++) {
      t.ev[n] = ev[(size_t)l * nmax + n];
      for (int i = 0; i < numr; i++) t.ef(n, i) = ef[((size_t)l * nmax + n) * numr + i];
    }
  }
  h->sph.ortho = &h->grid;
  return h;
}
void ref_sph_destroy(void *hh) { delete static_cast<SphHandle *>(hh); }

void ref_slgrid_consts(void *hh, double *xmin, double *xmax, double *dxi)
{
  SLGridSph &g = static_cast<SphHandle *>(hh)->grid;
  *xmin = g.xmin; *xmax = g.xmax; *dxi = g.dxi;
}
double ref_r_to_xi(void *hh, double r) { return static_cast<SphHandle *>(hh)->grid.r_to_xi(r); }
double ref_xi_to_r(void *hh, double x) { return static_cast<SphHandle *>(hh)->grid.xi_to_r(x); }
double ref_d_xi_to_r(void *hh, double x) { return static_cast<SphHandle *>(hh)->grid.d_xi_to_r(x); }

static void copy_mat(const Eigen::MatrixXd &m, double *out)
{
  for (int i = 0; i < m.rows(); i++)
    for (int j = 0; j < m.cols(); j++) out[(size_t)i * m.cols() + j] = m(i, j);
}
void ref_slgrid_get_pot(void *hh, double r, double *mat)
{
  Eigen::MatrixXd m; static_cast<SphHandle *>(hh)->grid.get_pot(m, r); copy_mat(m, mat);
}
void ref_slgrid_get_force(void *hh, double r, double *mat)
{
  Eigen::MatrixXd m; static_cast<SphHandle *>(hh)->grid.get_force(m, r); copy_mat(m, mat);
}
void ref_slgrid_get_dens(void *hh, double r, double *mat)
{
  Eigen::MatrixXd m; static_cast<SphHandle *>(hh)->grid.get_dens(m, r); copy_mat(m, mat);
}
double ref_factorial(int lmax, int l, int m)
{
  Sphere s; s.setup(lmax, 1, 1); return s.factorial(l, m);
}

static void sph_config(SphHandle *h, double scale, double rmin, double rmax, const int *flags, int nthreads)
{
  Sphere &s = h->sph;
  s.setup(h->grid.lmax, h->grid.nmax, nthreads);
  s.scale = scale; s.rmin = rmin; s.rmax = rmax;
  s.NO_L0 = flags[0]; s.NO_L1 = flags[1]; s.EVEN_L = flags[2]; s.EVEN_M = flags[3]; s.M0_only = flags[4];
  s.component = &h->comp;
  s.cC = &h->comp;
}

/* SphericalBasis::determine_coefficients_thread for every thread id, then the thread sum of
 * src/SphericalBasis.cc:855-858.  list = levlist[mlevel].  With sampT > 0 the pcavar/pcaeof
 * covariance branches run too (one thread, as the reference serialises them with a mutex):
 * massT[sampT], countT[sampT], coefT[sampT][ldim][nmax][2], coefM[sampT][ldim][nmax][nmax][2],
 * tvar[ldim][nmax][nmax] (pcaeof; may be NULL). */
long ref_sph_coefficients(void *hh, double scale, double rmin, double rmax, const int *flags,
                          long n, const double *pos, const double *mass, const double *center,
                          double rtrunc, const int *list, long nlist, double adb, int nthreads,
                          double *coef, int sampT, double *muse, double *massT, unsigned *countT,
                          double *coefT, double *coefM, double *tvar)
{
  SphHandle *h = static_cast<SphHandle *>(hh);
  sph_config(h, scale, rmin, rmax, flags, nthreads);
  Sphere &s = h->sph;
  fill_component(h->comp, n, pos, nullptr, nullptr, nullptr, nullptr, mass, nullptr, nullptr, nullptr,
                 center, rtrunc);
  h->comp.levlist.assign(1, std::vector<int>(list, list + nlist));
  h->comp.adiabatic_value = adb;
  s.mlevel = 0;
  const int L = s.Lmax, N = s.nmax, ldim = (L + 1) * (L + 2) / 2, nrow = (L + 1) * (L + 1);
  s.compute = sampT > 0;
  s.pcavar = sampT > 0;
  s.pcaeof = sampT > 0 && tvar != nullptr;
  s.sampT = sampT > 0 ? sampT : 1;
  if (s.compute) {
    s.countT1.assign(sampT, 0);
    s.massT1.assign(sampT, 0.0);
    s.expcoefT1.assign(sampT, std::vector<SphericalBasis::VectorCP>(ldim));
    s.expcoefM1.assign(sampT, std::vector<SphericalBasis::MatrixCP>(ldim));
    for (int T = 0; T < sampT; T++)
      for (int k = 0; k < ldim; k++) {
        s.expcoefT1[T][k] = std::make_shared<Eigen::VectorXcd>(N);
        s.expcoefM1[T][k] = std::make_shared<Eigen::MatrixXcd>(N, N);
      }
    s.tvar.resize(ldim);
    for (auto &p : s.tvar) p = std::make_shared<Eigen::MatrixXd>(N, N);
  }
  for (int id = 0; id < nthreads; id++) {
    int arg = id;
    s.determine_coefficients_thread(&arg);
  }
  long use1 = 0;
  for (int i = 0; i < nthreads; i++) use1 += s.use[i];
  for (int i = 1; i < nthreads; i++)
    for (int l = 0; l < nrow; l++)
      for (int k = 0; k < N; k++) (*s.expcoef0[0][l])[k] += (*s.expcoef0[i][l])[k];
  for (int l = 0; l < nrow; l++)
    for (int k = 0; k < N; k++) coef[(size_t)l * N + k] = (*s.expcoef0[0][l])[k];
  if (s.compute) {
    if (muse) { *muse = 0.0; for (int i = 0; i < nthreads; i++) *muse += s.muse1[i]; }
    for (int T = 0; T < sampT; T++) {
      if (massT) massT[T] = s.massT1[T];
      if (countT) countT[T] = s.countT1[T];
      for (int k = 0; k < ldim; k++)
        for (int a = 0; a < N; a++) {
          if (coefT) {
            coefT[(((size_t)T * ldim + k) * N + a) * 2 + 0] = (*s.expcoefT1[T][k])[a].real();
            coefT[(((size_t)T * ldim + k) * N + a) * 2 + 1] = (*s.expcoefT1[T][k])[a].imag();
          }
          if (coefM)
            for (int b = 0; b < N; b++) {
              coefM[((((size_t)T * ldim + k) * N + a) * N + b) * 2 + 0] = (*s.expcoefM1[T][k])(a, b).real();
              coefM[((((size_t)T * ldim + k) * N + a) * N + b) * 2 + 1] = (*s.expcoefM1[T][k])(a, b).imag();
            }
        }
    }
    if (tvar)
      for (int k = 0; k < ldim; k++)
        for (int a = 0; a < N; a++)
          for (int b = 0; b < N; b++) tvar[((size_t)k * N + a) * N + b] = (*s.tvar[k])(a, b);
  }
  return use1;
}

/* SphericalBasis::determine_acceleration_and_potential_thread for every thread id over the
 * particles of `list` (levlist[lev >= mlevel] concatenated by the caller), adding into
 * acc/pot.  own_center != NULL: the particles belong to another component (use_external):
 * positions are converted with the owning component's centre, freeze() uses `center`. */
void ref_sph_accel(void *hh, double scale, double rmin, double rmax, const int *flags, long n,
                   const double *pos, const double *center, double rtrunc, const int *list, long nlist,
                   const double *coef, int nthreads, const double *own_center, double *acc, double *pot)
{
  SphHandle *h = static_cast<SphHandle *>(hh);
  sph_config(h, scale, rmin, rmax, flags, nthreads);
  Sphere &s = h->sph;
  fill_component(h->comp, n, pos, nullptr, acc, pot, nullptr, nullptr, nullptr, nullptr, nullptr, center, rtrunc);
  h->comp.levlist.assign(1, std::vector<int>(list, list + nlist));
  s.mlevel = 0;
  multistep = 0;
  s.use_external = 0;
  if (own_center) {
    s.use_external = 1;
    h->other.store.clear();
    for (int k = 0; k < 3; k++) { h->other.center[k] = own_center[k]; h->other.com0[k] = 0.0; }
    s.component = &h->other;
  }
  const int N = s.nmax, nrow = (s.Lmax + 1) * (s.Lmax + 1);
  for (int l = 0; l < nrow; l++)
    for (int k = 0; k < N; k++) (*s.expcoef[l])[k] = coef[(size_t)l * N + k];
  for (int id = 0; id < nthreads; id++) {
    int arg = id;
    s.determine_acceleration_and_potential_thread(&arg);
  }
  for (long i = 0; i < n; i++) {
    for (int k = 0; k < 3; k++) acc[3 * i + k] = h->comp.store[i].acc[k];
    pot[i] = h->comp.store[i].pot;
  }
}

/* ---- one KDK step with the reference's own thread functions on `nthreads` host threads ------
 * (the CPU baseline of bench.py): do_step's single-level sequence (src/step.cc:272-322):
 * incr_velocity(dt/2), incr_position(dt), determine_coefficients (threads + thread sum),
 * zero acc/pot, determine_acceleration_and_potential (threads), incr_velocity(dt/2), forked and
 * joined per phase like exp_thread_fork (src/PotAccel.cc:97-130).  The particle store is the flat
 * Component of ref_shim/pathshim.H (an index instead of the reference's hash map). */
void ref_sph_load(void *hh, double scale, double rmin, double rmax, const int *flags, long n,
                  const double *pos, const double *vel, const double *mass, int nthreads)
{
  SphHandle *h = static_cast<SphHandle *>(hh);
  sph_config(h, scale, rmin, rmax, flags, nthreads);
  fill_component(h->comp, n, pos, vel, nullptr, nullptr, nullptr, mass, nullptr, nullptr, nullptr, nullptr, 1e20);
  std::vector<int> all(n);
  for (long i = 0; i < n; i++) all[i] = (int)i;
  h->comp.levlist.assign(1, all);
  h->comp.pmap.clear();
  for (long i = 0; i < n; i++) h->comp.pmap[(unsigned long)i] = &h->comp.store[i];
  h->sph.mlevel = 0;
  h->sph.compute = false;
  multistep = 0;
}

} // extern "C"
#include <thread>
namespace {
template <class F> void fork_join(int nthreads, F &&f)
{
  std::vector<std::thread> th;
  for (int id = 1; id < nthreads; id++) th.emplace_back([&, id] { f(id); });
  f(0);
  for (auto &t : th) t.join();
}
} // namespace
extern "C" {

long ref_sph_step(void *hh, double dt, int with_kicks)
{
  SphHandle *h = static_cast<SphHandle *>(hh);
  Sphere &s = h->sph;
  const int nth = nthrds;
  ComponentContainer cc;
  cc.components.push_back(&h->comp);
  comp = &cc;
  if (with_kicks) {
    fork_join(nth, [&](int id) { thrd_pass_posvel a{0.5 * dt, -1, id}; incr_velocity_thread(&a); });
    fork_join(nth, [&](int id) { thrd_pass_posvel a{dt, -1, id}; incr_position_thread(&a); });
  }
  const int N = s.nmax, nrow = (s.Lmax + 1) * (s.Lmax + 1);
  for (auto &v : s.expcoef0) for (auto &p : v) p->setZero();
  std::fill(s.use.begin(), s.use.end(), 0);
  fork_join(nth, [&](int id) { int arg = id; s.determine_coefficients_thread(&arg); });
  long use1 = 0;
  for (int i = 0; i < nth; i++) use1 += s.use[i];
  for (int i = 1; i < nth; i++)
    for (int l = 0; l < nrow; l++)
      for (int k = 0; k < N; k++) (*s.expcoef0[0][l])[k] += (*s.expcoef0[i][l])[k];
  for (int l = 0; l < nrow; l++)
    for (int k = 0; k < N; k++) (*s.expcoef[l])[k] = (*s.expcoef0[0][l])[k];
  for (auto &p : h->comp.store) { p.acc[0] = p.acc[1] = p.acc[2] = 0.0; p.pot = 0.0; }
  s.use_external = 0;
  fork_join(nth, [&](int id) { int arg = id; s.determine_acceleration_and_potential_thread(&arg); });
  if (with_kicks)
    fork_join(nth, [&](int id) { thrd_pass_posvel a{0.5 * dt, -1, id}; incr_velocity_thread(&a); });
  comp = nullptr;
  return use1;
}

void ref_sph_read(void *hh, double *pos, double *vel, double *acc, double *pot, double *coef)
{
  SphHandle *h = static_cast<SphHandle *>(hh);
  const long n = (long)h->comp.store.size();
  for (long i = 0; i < n; i++) {
    const Particle &p = h->comp.store[i];
    for (int k = 0; k < 3; k++) {
      if (pos) pos[3 * i + k] = p.pos[k];
      if (vel) vel[3 * i + k] = p.vel[k];
      if (acc) acc[3 * i + k] = p.acc[k];
    }
    if (pot) pot[i] = p.pot;
  }
  if (coef) {
    const int N = h->sph.nmax, nrow = (h->sph.Lmax + 1) * (h->sph.Lmax + 1);
    for (int l = 0; l < nrow; l++)
      for (int k = 0; k < N; k++) coef[(size_t)l * N + k] = (*h->sph.expcoef[l])[k];
  }
}

/* ---- EmpCylSL ------------------------------------------------------------------------------ */
/* CylEXP (src/CylEXP.H): EmpCylSL + the multistep level tables; only the level-switch update is
 * compiled here (src/CylEXP.cc:159-188) */
class CylEXP : public EmpCylSL {
public:
  std::vector<std::vector<Eigen::MatrixXd>> differC1, differS1;     /* [thread][level](MMAX+1, rank3) */
  MstepArray cosL, sinL;
  void multistep_update(int from, int to, double r, double z, double phi, double mass, int id);
  void compute_multistep_coefficients(unsigned mlevel);
};
#include "_ref/gen/cylexp_multistep_update.inc"
#include "_ref/gen/cylexp_compute_multistep_coefficients.inc"

struct CylHandle { CylEXP e; };

void *ref_cyl_create(int MMAX, int NORDER, int NUMX, int NUMY, int CMAPR, int CMAPZ, double ASCALE,
                     double HSCALE, double Rtable, double XMIN, double dX, double YMIN, double dY,
                     const int *flags /* EVEN_M MMIN MLIM NMIN NLIM */, const double *potC,
                     const double *potS, const double *rforceC, const double *rforceS,
                     const double *zforceC, const double *zforceS)
{
  CylHandle *h = new CylHandle;
  EmpCylSL &e = h->e;
  e.MMAX = MMAX; e.NORDER = NORDER; e.rank3 = NORDER; e.NUMX = NUMX; e.NUMY = NUMY;
  e.CMAPR = CMAPR; e.CMAPZ = CMAPZ; e.ASCALE = ASCALE; e.HSCALE = HSCALE; e.Rtable = Rtable;
  e.XMIN = XMIN; e.dX = dX; e.YMIN = YMIN; e.dY = dY;
  e.EVEN_M = flags[0]; e.MMIN = flags[1]; e.MLIM = flags[2]; e.NMIN = flags[3]; e.NLIM = flags[4];
  auto load = [&](std::vector<std::vector<Eigen::MatrixXd>> &dst, const double *src) {
    dst.assign(MMAX + 1, std::vector<Eigen::MatrixXd>(NORDER, Eigen::MatrixXd(NUMX + 1, NUMY + 1)));
    if (!src) return;
    for (int m = 0; m <= MMAX; m++)
      for (int n = 0; n < NORDER; n++)
        for (int ix = 0; ix <= NUMX; ix++)
          for (int iy = 0; iy <= NUMY; iy++)
            dst[m][n](ix, iy) = src[(((size_t)m * NORDER + n) * (NUMX + 1) + ix) * (NUMY + 1) + iy];
  };
  load(e.potC, potC); load(e.potS, potS); load(e.rforceC, rforceC); load(e.rforceS, rforceS);
  load(e.zforceC, zforceC); load(e.zforceS, zforceS);
  e.accum_cos.assign(MMAX + 1, Eigen::VectorXd(NORDER));
  e.accum_sin.assign(MMAX + 1, Eigen::VectorXd(NORDER));
  return h;
}
void ref_cyl_destroy(void *hh) { delete static_cast<CylHandle *>(hh); }
double ref_cyl_r_to_xi(void *hh, double r) { return static_cast<CylHandle *>(hh)->e.r_to_xi(r); }
double ref_cyl_z_to_y(void *hh, double z) { return static_cast<CylHandle *>(hh)->e.z_to_y(z); }

void ref_cyl_get_pot(void *hh, double r, double z, double *Vc, double *Vs)
{
  EmpCylSL &e = static_cast<CylHandle *>(hh)->e;
  Eigen::MatrixXd vc, vs;
  e.get_pot(vc, vs, r, z);
  for (int m = 0; m <= e.MMAX; m++)
    for (int n = 0; n < e.NORDER; n++) {
      Vc[(size_t)m * e.NORDER + n] = vc(m, n);
      Vs[(size_t)m * e.NORDER + n] = m ? vs(m, n) : 0.0;
    }
}

/* EmpCylSL::accumulate for np particles already in cylinder coordinates (r, z, phi, mass, seq),
 * one thread.  cosN/sinN [(MMAX+1)][NORDER]; with sampT > 0 the covar branch:
 * numbT[sampT], massT[sampT], VC[sampT][MMAX+1][NORDER][2], MV[sampT][MMAX+1][NORDER][NORDER][2].
 * Returns howmany; *cylmass1 as the reference increments it (inside the m loop). */
long ref_cyl_accumulate(void *hh, long np, const double *r, const double *z, const double *phi,
                        const double *mass, const unsigned long *seq, double *cosN, double *sinN,
                        double *cylmass1, int sampT, unsigned *numbT, double *massT, double *VC, double *MV)
{
  EmpCylSL &e = static_cast<CylHandle *>(hh)->e;
  const int M = e.MMAX, N = e.NORDER;
  e.coefs_made.assign(1, false);
  e.howmany1.assign(1, std::vector<unsigned>(1, 0));
  e.vc.assign(1, Eigen::MatrixXd());
  e.vs.assign(1, Eigen::MatrixXd());
  e.cylmass1.assign(1, 0.0);
  e.cosN.assign(1, std::make_shared<EmpCylSL::VectorD2>(1, std::vector<Eigen::VectorXd>(M + 1, Eigen::VectorXd(N))));
  e.sinN.assign(1, std::make_shared<EmpCylSL::VectorD2>(1, std::vector<Eigen::VectorXd>(M + 1, Eigen::VectorXd(N))));
  e.covar = sampT > 0;
  e.sampT = sampT > 0 ? sampT : 0;
  if (e.covar) {
    e.numbT1.assign(1, std::vector<unsigned>(sampT, 0));
    e.massT1.assign(1, std::vector<double>(sampT, 0.0));
    e.VC.assign(1, std::vector<std::vector<Eigen::VectorXcd>>(sampT, std::vector<Eigen::VectorXcd>(M + 1, Eigen::VectorXcd(N))));
    e.MV.assign(1, std::vector<std::vector<Eigen::MatrixXcd>>(sampT, std::vector<Eigen::MatrixXcd>(M + 1, Eigen::MatrixXcd(N, N))));
  }
  for (long i = 0; i < np; i++) e.accumulate(r[i], z[i], phi[i], mass[i], seq ? seq[i] : (unsigned long)i, 0, 0, e.covar);
  for (int m = 0; m <= M; m++)
    for (int n = 0; n < N; n++) {
      cosN[(size_t)m * N + n] = e.cosN(0)[0][m][n];
      sinN[(size_t)m * N + n] = e.sinN(0)[0][m][n];
    }
  if (cylmass1) *cylmass1 = e.cylmass1[0];
  if (e.covar)
    for (int T = 0; T < sampT; T++) {
      if (numbT) numbT[T] = e.numbT1[0][T];
      if (massT) massT[T] = e.massT1[0][T];
      for (int m = 0; m <= M; m++)
        for (int a = 0; a < N; a++) {
          if (VC) {
            VC[(((size_t)T * (M + 1) + m) * N + a) * 2 + 0] = e.VC[0][T][m][a].real();
            VC[(((size_t)T * (M + 1) + m) * N + a) * 2 + 1] = e.VC[0][T][m][a].imag();
          }
          if (MV)
            for (int b = 0; b < N; b++) {
              MV[((((size_t)T * (M + 1) + m) * N + a) * N + b) * 2 + 0] = e.MV[0][T][m](a, b).real();
              MV[((((size_t)T * (M + 1) + m) * N + a) * N + b) * 2 + 1] = e.MV[0][T][m](a, b).imag();
            }
        }
    }
  return (long)e.howmany1[0][0];
}

/* the PCAVAR (which = 1) / PCAEOF (which = 2) branches of EmpCylSL::accumulate:
 * covV[sampT][MMAX+1][NORDER], covM[sampT][MMAX+1][NORDER][NORDER] (PCAVAR) or
 * tvar[MMAX+1][NORDER][NORDER] written to covM with sampT = 1 (PCAEOF) */
void ref_cyl_accumulate_pca(void *hh, long np, const double *r, const double *z, const double *phi,
                            const double *mass, const unsigned long *seq, int sampT, int which,
                            double *covV, double *covM)
{
  EmpCylSL &e = static_cast<CylHandle *>(hh)->e;
  const int M = e.MMAX, N = e.NORDER;
  e.coefs_made.assign(1, false);
  e.howmany1.assign(1, std::vector<unsigned>(1, 0));
  e.vc.assign(1, Eigen::MatrixXd());
  e.vs.assign(1, Eigen::MatrixXd());
  e.cylmass1.assign(1, 0.0);
  e.cosN.assign(1, std::make_shared<EmpCylSL::VectorD2>(1, std::vector<Eigen::VectorXd>(M + 1, Eigen::VectorXd(N))));
  e.sinN.assign(1, std::make_shared<EmpCylSL::VectorD2>(1, std::vector<Eigen::VectorXd>(M + 1, Eigen::VectorXd(N))));
  e.covar = false;
  e.PCAVAR = which == 1;
  e.PCAEOF = which == 2;
  e.sampT = sampT;
  e.numbT1.assign(1, std::vector<unsigned>(sampT, 0));
  e.massT1.assign(1, std::vector<double>(sampT, 0.0));
  e.covV.assign(1, std::vector<std::vector<Eigen::VectorXd>>(sampT, std::vector<Eigen::VectorXd>(M + 1, Eigen::VectorXd(N))));
  e.covM.assign(1, std::vector<std::vector<Eigen::MatrixXd>>(sampT, std::vector<Eigen::MatrixXd>(M + 1, Eigen::MatrixXd(N, N))));
  e.tvar.assign(1, std::vector<Eigen::MatrixXd>(M + 1, Eigen::MatrixXd(N, N)));
  for (long i = 0; i < np; i++) e.accumulate(r[i], z[i], phi[i], mass[i], seq ? seq[i] : (unsigned long)i, 0, 0, true);
  e.PCAVAR = e.PCAEOF = false;
  for (int T = 0; T < sampT; T++)
    for (int m = 0; m <= M; m++)
      for (int a = 0; a < N; a++) {
        if (covV) covV[((size_t)T * (M + 1) + m) * N + a] = which == 1 ? e.covV[0][T][m][a] : 0.0;
        if (covM)
          for (int b = 0; b < N; b++)
            covM[(((size_t)T * (M + 1) + m) * N + a) * N + b] =
                which == 1 ? e.covM[0][T][m](a, b) : (T == 0 ? e.tvar[0][m](a, b) : 0.0);
      }
}

void ref_cyl_accumulated_eval(void *hh, const double *accum_cos, const double *accum_sin, double r,
                              double z, double phi, double *p0, double *p, double *fr, double *fz, double *fp)
{
  EmpCylSL &e = static_cast<CylHandle *>(hh)->e;
  for (int m = 0; m <= e.MMAX; m++)
    for (int n = 0; n < e.NORDER; n++) {
      e.accum_cos[m][n] = accum_cos[(size_t)m * e.NORDER + n];
      e.accum_sin[m][n] = accum_sin[(size_t)m * e.NORDER + n];
    }
  *p0 = 0.0;
  e.accumulated_eval(r, z, phi, *p0, *p, *fr, *fz, *fp);
}

/* ---- leapfrog: incr_position_thread / incr_velocity_thread over one component ---------------- */
static void leapfrog(bool position, long n, double *pos, double *vel, const double *acc,
                     const unsigned *level, int nlev, double dt, int mlevel, int dim, int nthreads)
{
  Component c;
  fill_component(c, n, pos, vel, acc, nullptr, nullptr, nullptr, nullptr, level, nullptr, nullptr, 1e20);
  c.dim = dim;
  c.levlist.assign(nlev, std::vector<int>());
  for (long i = 0; i < n; i++) c.levlist[level ? level[i] : 0].push_back((int)i);
  for (long i = 0; i < n; i++) c.pmap[(unsigned long)i] = &c.store[i];
  ComponentContainer cc;
  cc.components.push_back(&c);
  comp = &cc;
  nthrds = nthreads;
  for (int id = 0; id < nthreads; id++) {
    thrd_pass_posvel a{dt, mlevel, id};
    if (position) incr_position_thread(&a);
    else incr_velocity_thread(&a);
  }
  comp = nullptr;
  for (long i = 0; i < n; i++)
    for (int k = 0; k < 3; k++) { pos[3 * i + k] = c.store[i].pos[k]; vel[3 * i + k] = c.store[i].vel[k]; }
}
void ref_incr_position(long n, double *pos, double *vel, const unsigned *level, int nlev, double dt,
                       int mlevel, int dim, int nthreads)
{
  leapfrog(true, n, pos, vel, nullptr, level, nlev, dt, mlevel, dim, nthreads);
}
void ref_incr_velocity(long n, double *pos, double *vel, const double *acc, const unsigned *level, int nlev,
                       double dt, int mlevel, int dim, int nthreads)
{
  leapfrog(false, n, pos, vel, acc, level, nlev, dt, mlevel, dim, nthreads);
}

/* ---- adjust_multistep_level_thread over the particles of every level >= minlev -------------- */
struct ref_level_params {
  double dtime, dynfracS, dynfracD, dynfracV, dynfracA, dynfracP;
  int multistep, shiftlevl, mfirst_mdrft, minlev, noswitch, first_call, dim, nthreads;
};
long ref_adjust_levels(long n, const double *vel, const double *acc, const double *pot, const double *potext,
                       const float *scale, unsigned *level, float *dtreq, const ref_level_params *P,
                       long *offlo, long *offhi)
{
  Component c;
  std::vector<double> zero(3 * n, 0.0);
  fill_component(c, n, zero.data(), vel, acc, pot, potext, nullptr, scale, level, dtreq, nullptr, 1e20);
  c.dim = P->dim;
  c.noswitch = P->noswitch;
  PotAccel force;
  c.force = &force;
  multistep = (unsigned)P->multistep;
  c.levlist.assign(P->multistep + 1, std::vector<int>());
  for (long i = 0; i < n; i++) c.levlist[level[i]].push_back((int)i);
  dtime = P->dtime; dynfracS = P->dynfracS; dynfracD = P->dynfracD; dynfracV = P->dynfracV;
  dynfracA = P->dynfracA; dynfracP = P->dynfracP; shiftlevl = P->shiftlevl;
  // firstCall = this_step==0 and mdrft==0 (src/multistep.cc:129); otherwise evaluate at a
  // substep whose lowest active level is mfirst_mdrft
  Mstep = 1 << P->multistep;
  mfirst.assign(Mstep + 1, 0);
  if (P->first_call) { this_step = 0; mdrft = 0; mstep = 0; mfirst[0] = P->mfirst_mdrft; }
  else { this_step = 1; mdrft = Mstep; mstep = 0; mfirst[Mstep] = P->mfirst_mdrft; }
  nthrds = P->nthreads;
  mindt1.assign(nthrds, 1.0e20); maxdt1.assign(nthrds, -1.0e20);
  adjtm1.assign(nthrds, 0.0); adjtm2.assign(nthrds, 0.0);
  numsw.assign(nthrds, 0); numtt.assign(nthrds, 0);
  offlo1[&c] = std::vector<unsigned>(nthrds, 0);
  offhi1[&c] = std::vector<unsigned>(nthrds, 0);
  tmdt.assign(nthrds, std::vector<std::vector<unsigned>>(P->multistep + 1, std::vector<unsigned>(mdtDim, 0)));
  for (int lev = P->minlev; lev <= P->multistep; lev++)
    for (int id = 0; id < nthrds; id++) {
      thrd_pass_sync a{lev, id, &c};
      adjust_multistep_level_thread(&a);
    }
  long sw = 0;
  *offlo = *offhi = 0;
  for (int id = 0; id < nthrds; id++) { sw += numsw[id]; *offlo += offlo1[&c][id]; *offhi += offhi1[&c][id]; }
  for (long i = 0; i < n; i++) { level[i] = c.store[i].level; dtreq[i] = c.store[i].dtreq; }
  offlo1.clear(); offhi1.clear();
  return sw;
}

} // extern "C"


// =============================================================================================
// Cylinder::determine_coefficients_thread / determine_acceleration_and_potential_thread
// (src/Cylinder.cc:748-896, 1266-1446), the reference's own bodies over the EmpCylSL of a
// ref_cyl_create handle and a flat Component: centre / rtrunc gate, body-frame rotation, the
// rcylmax gate, the erf blend to the monopole and the rotation of the force back.
// =============================================================================================
class Cylinder : public Basis {
public:
  CylEXP *ortho = nullptr;
  bool play_back = false, self_consistent = true;
  void multistep_update(int from, int to, Component *c, int i, int id) override;
  MixtureBasis *mix = nullptr;
  bool eof = false, compute = false;
  double rcylmax = 20.0, acyl = 1.0, cylmass = 0.0;
  std::vector<Eigen::Vector3d> pos, frc;
  std::vector<double> cylmass0;
  std::vector<int> offgrid;
  void *determine_coefficients_thread(void *arg);
  void *determine_acceleration_and_potential_thread(void *arg);
};

#include "_ref/gen/cylinder_multistep_update.inc"
#include "_ref/gen/cylinder_coef_thread.inc"
#include "_ref/gen/cylinder_accel_thread.inc"

namespace {
void cyl_component(Component &c, Orient &o, long n, const double *pos, const double *mass, const double *acc,
                   const double *pot, const unsigned *level, int nlev, const double *center, double rtrunc,
                   const double *body, const double *orig)
{
  fill_component(c, n, pos, nullptr, acc, pot, nullptr, mass, nullptr, level, nullptr, center, rtrunc);
  c.levlist.assign(nlev, std::vector<int>());
  for (long i = 0; i < n; i++) c.levlist[level ? level[i] : 0].push_back((int)i);
  for (long i = 0; i < n; i++) c.pmap[(unsigned long)i] = &c.store[i];
  if (body) {
    for (int k = 0; k < 9; k++) { o.body.m[k] = body[k]; o.orig.m[k] = orig[k]; }
    c.EJ = Orient::AXIS;
    c.orient = &o;
  }
}
} // namespace

extern "C" {

/* coefficient thread for levlist[mlevel] of one component, one thread: cosN/sinN [(MMAX+1)][NORDER],
 * *used, *cylmass0 as Cylinder::determine_coefficients sums them over threads (:1080-1086) */
void ref_cylinder_coef_thread(void *hh, long n, const double *pos, const double *mass, const unsigned *level,
                              int nlev, int mlevel, const double *center, double rtrunc, double rcylmax,
                              double acyl, double adb, const double *body, const double *orig, double *cosN,
                              double *sinN, int *used, double *cylmass0)
{
  CylEXP &e = static_cast<CylHandle *>(hh)->e;
  const int M = e.MMAX, N = e.NORDER;
  e.coefs_made.assign(nlev, false);
  e.howmany1.assign(nlev, std::vector<unsigned>(1, 0));
  e.vc.assign(1, Eigen::MatrixXd());
  e.vs.assign(1, Eigen::MatrixXd());
  e.cylmass1.assign(1, 0.0);
  e.cosN.clear(); e.sinN.clear();
  for (int L = 0; L < nlev; L++) {
    e.cosN.push_back(std::make_shared<EmpCylSL::VectorD2>(1, std::vector<Eigen::VectorXd>(M + 1, Eigen::VectorXd(N))));
    e.sinN.push_back(std::make_shared<EmpCylSL::VectorD2>(1, std::vector<Eigen::VectorXd>(M + 1, Eigen::VectorXd(N))));
  }
  e.covar = e.PCAVAR = e.PCAEOF = false;
  Component c;
  Orient o;
  cyl_component(c, o, n, pos, mass, nullptr, nullptr, level, nlev, center, rtrunc, body, orig);
  c.adiabatic_value = adb;
  Cylinder cy;
  cy.ortho = &e; cy.component = &c; cy.cC = &c; cy.mlevel = (unsigned)mlevel;
  cy.rcylmax = rcylmax; cy.acyl = acyl;
  cy.pos.resize(1); cy.frc.resize(1); cy.use.assign(1, 0); cy.cylmass0.assign(1, 0.0);
  multistep = (unsigned)(nlev - 1);
  nthrds = 1;
  int id = 0;
  cy.determine_coefficients_thread(&id);
  for (int m = 0; m <= M; m++)
    for (int k = 0; k < N; k++) {
      cosN[(size_t)m * N + k] = e.cosN(mlevel)[0][m][k];
      sinN[(size_t)m * N + k] = e.sinN(mlevel)[0][m][k];
    }
  *used = cy.use[0];
  *cylmass0 = cy.cylmass0[0];
}

/* Cylinder::multistep_update + CylEXP::multistep_update (src/Cylinder.cc:1752-1773, src/CylEXP.cc:159-188)
 * for a list of switching particles: differC/differS [nlev][MMAX+1][NORDER] */
void ref_cylinder_multistep_update(void *hh, long n, const double *pos, const double *mass, const double *center,
                                   double rtrunc, double adb, int nlev, long nsw, const int *idx, const int *from,
                                   const int *to, double *differC, double *differS)
{
  CylEXP &e = static_cast<CylHandle *>(hh)->e;
  const int M = e.MMAX, N = e.NORDER;
  e.vc.assign(1, Eigen::MatrixXd());
  e.vs.assign(1, Eigen::MatrixXd());
  e.differC1.assign(1, std::vector<Eigen::MatrixXd>(nlev, Eigen::MatrixXd(M + 1, N)));
  e.differS1.assign(1, std::vector<Eigen::MatrixXd>(nlev, Eigen::MatrixXd(M + 1, N)));
  Component c;
  Orient o;
  cyl_component(c, o, n, pos, mass, nullptr, nullptr, nullptr, 1, center, rtrunc, nullptr, nullptr);
  c.adiabatic_value = adb;
  Cylinder cy;
  cy.ortho = &e; cy.component = &c; cy.cC = &c;
  for (long k = 0; k < nsw; k++) cy.multistep_update(from[k], to[k], &c, idx[k], 0);
  for (int L = 0; L < nlev; L++)
    for (int m = 0; m <= M; m++)
      for (int k = 0; k < N; k++) {
        differC[((size_t)L * (M + 1) + m) * N + k] = e.differC1[0][L](m, k);
        differS[((size_t)L * (M + 1) + m) * N + k] = e.differS1[0][L](m, k);
      }
}

/* CylEXP::compute_multistep_coefficients (src/CylEXP.cc:192-282) with the tables of
 * initialize_multistep: cosL/cosN/sinL/sinN [nlev][MMAX+1][NORDER] -> accum_cos, accum_sin */
void ref_cylexp_compute_multistep_coefficients(void *hh, int nlev, int Mstep_, int mdrft_, const int *mfirst_,
                                               const int *dL, const int *dN, const double *cosL,
                                               const double *cosN, const double *sinL, const double *sinN,
                                               double *accum_cos, double *accum_sin)
{
  CylEXP &e = static_cast<CylHandle *>(hh)->e;
  const int M1 = e.MMAX + 1, N = e.NORDER;
  multistep = (unsigned)(nlev - 1);
  Mstep = Mstep_; mdrft = mdrft_; mstep = mdrft_ > 0 ? mdrft_ - 1 : 0;
  mfirst.assign(mfirst_, mfirst_ + Mstep_ + 1);
  dstepL.assign(nlev, std::vector<int>(Mstep_ + 1, 0));
  dstepN.assign(nlev, std::vector<int>(Mstep_ + 1, 0));
  for (int M = 0; M < nlev; M++)
    for (int k = 0; k < Mstep_; k++) { dstepL[M][k] = dL[M * Mstep_ + k]; dstepN[M][k] = dN[M * Mstep_ + k]; }
  auto fill = [&](EmpCylSL::MstepArray &dst, const double *src) {
    dst.clear();
    for (int M = 0; M < nlev; M++) {
      dst.push_back(std::make_shared<EmpCylSL::VectorD2>(1, std::vector<Eigen::VectorXd>(M1, Eigen::VectorXd(N))));
      for (int m = 0; m < M1; m++)
        for (int k = 0; k < N; k++) dst(M)[0][m][k] = src[((size_t)M * M1 + m) * N + k];
    }
  };
  fill(e.cosL, cosL); fill(e.cosN, cosN); fill(e.sinL, sinL); fill(e.sinN, sinN);
  for (int m = 0; m < M1; m++) { e.accum_cos[m].setZero(); e.accum_sin[m].setZero(); }
  e.compute_multistep_coefficients(0);
  for (int m = 0; m < M1; m++)
    for (int k = 0; k < N; k++) {
      accum_cos[(size_t)m * N + k] = e.accum_cos[m][k];
      accum_sin[(size_t)m * N + k] = e.accum_sin[m][k];
    }
}

/* force thread: levels >= mlevel of the target component (pos/level/center/rtrunc) with the
 * coefficients accum_cos/accum_sin and cylmass installed; owner_center != NULL is the use_external
 * case (origin = the owning component's centre, orientation = the owner's).  acc[n][3], pot[n] are
 * added to. */
void ref_cylinder_accel_thread(void *hh, long n, const double *pos, const unsigned *level, int nlev, int mlevel,
                               const double *center, double rtrunc, const double *owner_center,
                               const double *accum_cos, const double *accum_sin, double cylmass,
                               const double *body, const double *orig, double *acc, double *pot)
{
  CylEXP &e = static_cast<CylHandle *>(hh)->e;
  for (int m = 0; m <= e.MMAX; m++)
    for (int k = 0; k < e.NORDER; k++) {
      e.accum_cos[m][k] = accum_cos[(size_t)m * e.NORDER + k];
      e.accum_sin[m][k] = accum_sin[(size_t)m * e.NORDER + k];
    }
  Component target, owner;
  Orient o, o2;
  cyl_component(target, o2, n, pos, nullptr, acc, pot, level, nlev, center, rtrunc, nullptr, nullptr);
  Cylinder cy;
  cy.ortho = &e; cy.cC = &target; cy.mlevel = (unsigned)mlevel; cy.cylmass = cylmass;
  if (owner_center) {
    cyl_component(owner, o, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nlev, owner_center, 1.0e20, body, orig);
    cy.component = &owner;
    cy.use_external = 1;
  } else {
    if (body) { for (int k = 0; k < 9; k++) { o.body.m[k] = body[k]; o.orig.m[k] = orig[k]; } target.EJ = Orient::AXIS; target.orient = &o; }
    cy.component = &target;
  }
  cy.pos.resize(1); cy.frc.resize(1); cy.offgrid.assign(1, 0);
  multistep = (unsigned)(nlev - 1);
  nthrds = 1;
  int id = 0;
  cy.determine_acceleration_and_potential_thread(&id);
  for (long i = 0; i < n; i++) {
    for (int k = 0; k < 3; k++) acc[3 * i + k] = target.store[i].acc[k];
    pot[i] = target.store[i].pot;
  }
}

} // extern "C"

// SphericalBasis::compute_multistep_coefficients (src/SphericalBasis.cc:1231-1474): the coefficient
// table in use at drift sub-step mdrft_ = levels below mfirst[mdrft_] interpolated between their
// last (L) and next (N) tables, the active levels' N tables added.  coefL, coefN [nlev][nrow][nmax];
// mfirst_ [Mstep+1], dL, dN [nlev][Mstep] are initialize_multistep's tables.
extern "C" {
void ref_sph_compute_multistep_coefficients(void *hh, int nlev, int Mstep_, int mdrft_, const int *mfirst_,
                                            const int *dL, const int *dN, const double *coefL,
                                            const double *coefN, double *out)
{
  SphHandle *h = static_cast<SphHandle *>(hh);
  const int flags[5] = {0, 0, 0, 0, 0};
  sph_config(h, 1.0, 0.0, 1.0e300, flags, 1);
  Sphere &s = h->sph;
  const int nrow = (s.Lmax + 1) * (s.Lmax + 1), N = s.nmax;
  multistep = (unsigned)(nlev - 1);
  Mstep = Mstep_; mdrft = mdrft_; mstep = mdrft_ > 0 ? mdrft_ - 1 : 0;
  mfirst.assign(mfirst_, mfirst_ + Mstep_ + 1);
  dstepL.assign(nlev, std::vector<int>(Mstep_ + 1, 0));
  dstepN.assign(nlev, std::vector<int>(Mstep_ + 1, 0));
  for (int M = 0; M < nlev; M++)
    for (int k = 0; k < Mstep_; k++) { dstepL[M][k] = dL[M * Mstep_ + k]; dstepN[M][k] = dN[M * Mstep_ + k]; }
  s.expcoefN.assign(nlev, std::vector<SphericalBasis::VectorP>(nrow));
  s.expcoefL.assign(nlev, std::vector<SphericalBasis::VectorP>(nrow));
  for (int M = 0; M < nlev; M++)
    for (int r = 0; r < nrow; r++) {
      s.expcoefN[M][r] = std::make_shared<Eigen::VectorXd>(N);
      s.expcoefL[M][r] = std::make_shared<Eigen::VectorXd>(N);
      for (int c = 0; c < N; c++) {
        (*s.expcoefN[M][r])[c] = coefN[((size_t)M * nrow + r) * N + c];
        (*s.expcoefL[M][r])[c] = coefL[((size_t)M * nrow + r) * N + c];
      }
    }
  s.compute_multistep_coefficients();
  for (int r = 0; r < nrow; r++)
    for (int c = 0; c < N; c++) out[(size_t)r * N + c] = (*s.expcoef[r])[c];
}
} // extern "C"

// SphericalBasis::multistep_update (src/SphericalBasis.cc:1156-1228) for a list of switching
// particles: differ[nlev][(Lmax+1)^2][nmax] accumulates -contribution at the old level and
// +contribution at the new one
extern "C" {
void ref_sph_multistep_update(void *hh, double scale, double rmin, double rmax, const int *flags, long n,
                              const double *pos, const double *mass, const double *center, double rtrunc,
                              double adb, int nlev, long nsw, const int *idx, const int *from, const int *to,
                              double *differ)
{
  SphHandle *h = static_cast<SphHandle *>(hh);
  sph_config(h, scale, rmin, rmax, flags, 1);
  Sphere &s = h->sph;
  fill_component(h->comp, n, pos, nullptr, nullptr, nullptr, nullptr, mass, nullptr, nullptr, nullptr, center, rtrunc);
  h->comp.adiabatic_value = adb;
  const int nrow = (s.Lmax + 1) * (s.Lmax + 1), N = s.nmax;
  s.differ1.assign(1, std::vector<Eigen::MatrixXd>(nlev, Eigen::MatrixXd(nrow, N)));
  for (long k = 0; k < nsw; k++) s.multistep_update(from[k], to[k], &h->comp, idx[k], 0);
  for (int L = 0; L < nlev; L++)
    for (int r = 0; r < nrow; r++)
      for (int c = 0; c < N; c++) differ[((size_t)L * nrow + r) * N + c] = s.differ1[0][L](r, c);
}
} // extern "C"

// =============================================================================================
// fix_positions_thread (src/Component.cc:3266-3352): per-level sums of mass, mass*pos, mass*vel,
// mass*acc over the non-frozen particles of the levels >= mlevel
// =============================================================================================
#include "_ref/gen/fix_positions_thread.inc"

extern "C" {
/* out[nlev][10] = {mtot, com[3], cov[3], coa[3]} per level (rows below mlevel stay zero) */
void ref_fix_positions(long n, const double *pos, const double *vel, const double *acc, const double *mass,
                       const unsigned *level, int nlev, int mlevel, int dim, const double *center, double rtrunc,
                       double *out)
{
  Component c;
  fill_component(c, n, pos, vel, acc, nullptr, nullptr, mass, nullptr, level, nullptr, center, rtrunc);
  c.dim = dim;
  c.levlist.assign(nlev, std::vector<int>());
  for (long i = 0; i < n; i++) c.levlist[level ? level[i] : 0].push_back((int)i);
  multistep = (unsigned)(nlev - 1);
  nthrds = 1;
  thrd_pass_posn d;
  d.id = 0; d.c = &c; d.tidal = -1; d.consp = false; d.com_system = false; d.mlevel = (unsigned)mlevel;
  d.com.assign(3 * nlev, 0.0); d.cov.assign(3 * nlev, 0.0); d.coa.assign(3 * nlev, 0.0); d.mtot.assign(nlev, 0.0);
  fix_positions_thread(&d);
  for (int mm = 0; mm < nlev; mm++) {
    out[10 * mm] = d.mtot[mm];
    for (int k = 0; k < 3; k++) {
      out[10 * mm + 1 + k] = d.com[3 * mm + k];
      out[10 * mm + 4 + k] = d.cov[3 * mm + k];
      out[10 * mm + 7 + k] = d.coa[3 * mm + k];
    }
  }
}
} // extern "C"

// =============================================================================================
// pyEXP twin: BasisClasses::Spherical::accumulate / sph_eval / cyl_eval / crt_eval / computeAccel
// and BiorthBasis::legendre_R, the reference's own bodies (expui/BiorthBasis.cc, cut by
// ref_extract.py) over the SLGridSph of a ref_sph_create handle.  accumulate is compiled around its
// pcavar block (pyEXP's own subsample covariance: not on the path), so pcavar stays false here.
// =============================================================================================
static inline int omp_get_thread_num() { return 0; }

namespace BasisClasses {

class BiorthBasis {
public:
  void legendre_R(int lmax, double x, Eigen::MatrixXd &p);
  void legendre_R(int lmax, double x, Eigen::MatrixXd &p, Eigen::MatrixXd &dp);
};

class Spherical : public BiorthBasis {
public:
  SLGridSph *sl = nullptr;
  int lmax = 0, nmax = 0, N1 = 0, N2 = 1 << 30, used = 0;
  unsigned sampT = 1;
  double scale = 1.0, rmin = 0.0, rmax = 1.0e300, totalMass = 0.0, G = 1.0;
  bool NO_L0 = false, NO_L1 = false, EVEN_L = false, EVEN_M = false, M0_only = false, pcavar = false;
  std::vector<Eigen::MatrixXd> potd, dpot, dend, legs, dlegs;
  Eigen::MatrixXd factorial, expcoef;
  Eigen::VectorXi sampleCounts;
  Eigen::VectorXd sampleMasses;

  void get_pot(Eigen::MatrixXd &tab, double r) { sl->get_pot(tab, r); }       // SphericalSL, BiorthBasis.H:522-531
  void get_dens(Eigen::MatrixXd &tab, double r) { sl->get_dens(tab, r); }
  void get_force(Eigen::MatrixXd &tab, double r) { sl->get_force(tab, r); }

  void initialize()
  {
    potd.resize(1); dpot.resize(1); dend.resize(1); legs.resize(1); dlegs.resize(1);
    potd[0].resize(lmax + 1, nmax); dpot[0].resize(lmax + 1, nmax); dend[0].resize(lmax + 1, nmax);
    legs[0].resize(lmax + 1, lmax + 1); dlegs[0].resize(lmax + 1, lmax + 1);
    expcoef.resize((lmax + 1) * (lmax + 1), nmax);
    expcoef.setZero();
#include "_ref/gen/biorth_factorial.inc"
  }
  void accumulate(double x, double y, double z, double mass, unsigned long indx);
  std::vector<double> sph_eval(double r, double costh, double phi);
  std::vector<double> cyl_eval(double R, double z, double phi);
  std::vector<double> crt_eval(double x, double y, double z);
  void computeAccel(double x, double y, double z, Eigen::Ref<Eigen::Vector3d> acc);
};

#include "_ref/gen/biorth_legendre.inc"
#include "_ref/gen/biorth_accumulate_head.inc"
#include "_ref/gen/biorth_accumulate_tail.inc"
#include "_ref/gen/biorth_sph_eval.inc"
#include "_ref/gen/biorth_compute_accel.inc"
#include "_ref/gen/biorth_cyl_eval.inc"
#include "_ref/gen/biorth_crt_eval.inc"

} // namespace BasisClasses

extern "C" {

// flags: bit 0 NO_L0, 1 NO_L1, 2 EVEN_L, 3 EVEN_M, 4 M0_only
void *ref_pyexp_create(void *sph_handle, double scale, double rmin, double rmax, int flags)
{
  SphHandle *h = static_cast<SphHandle *>(sph_handle);
  auto *b = new BasisClasses::Spherical;
  b->sl = &h->grid;
  b->lmax = h->grid.lmax; b->nmax = h->grid.nmax;
  b->scale = scale; b->rmin = rmin; b->rmax = rmax;
  b->NO_L0 = flags & 1; b->NO_L1 = flags & 2; b->EVEN_L = flags & 4; b->EVEN_M = flags & 8; b->M0_only = flags & 16;
  b->initialize();
  return b;
}
void ref_pyexp_destroy(void *bb) { delete static_cast<BasisClasses::Spherical *>(bb); }

// Spherical::accumulate over n particles (reset first): expcoef [(lmax+1)^2][nmax], returns used
int ref_pyexp_accumulate(void *bb, size_t n, const double *mass, const double *pos, double *expcoef)
{
  auto *b = static_cast<BasisClasses::Spherical *>(bb);
  b->expcoef.setZero();
  b->used = 0; b->totalMass = 0.0;
  for (size_t i = 0; i < n; i++) b->accumulate(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], mass[i], i);
  copy_mat(b->expcoef, expcoef);
  return b->used;
}
void ref_pyexp_set_coefs(void *bb, const double *expcoef)
{
  auto *b = static_cast<BasisClasses::Spherical *>(bb);
  for (int i = 0; i < b->expcoef.rows(); i++)
    for (int j = 0; j < b->expcoef.cols(); j++) b->expcoef(i, j) = expcoef[(size_t)i * b->expcoef.cols() + j];
}
// which: 0 sph_eval(r, costh, phi), 1 cyl_eval(R, z, phi), 2 crt_eval(x, y, z); out[9] per point
void ref_pyexp_eval(void *bb, int which, size_t n, const double *pts, double *out)
{
  auto *b = static_cast<BasisClasses::Spherical *>(bb);
  for (size_t i = 0; i < n; i++) {
    const double a = pts[3 * i], c = pts[3 * i + 1], d = pts[3 * i + 2];
    std::vector<double> v = which == 0 ? b->sph_eval(a, c, d) : which == 1 ? b->cyl_eval(a, c, d) : b->crt_eval(a, c, d);
    for (int k = 0; k < 9; k++) out[9 * i + k] = v[k];
  }
}
void ref_pyexp_accel(void *bb, size_t n, const double *pos, double *acc)
{
  auto *b = static_cast<BasisClasses::Spherical *>(bb);
  for (size_t i = 0; i < n; i++) {
    Eigen::Vector3d a;
    b->computeAccel(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], a);
    for (int k = 0; k < 3; k++) acc[3 * i + k] = a.d[k];
  }
}

} // extern "C"
