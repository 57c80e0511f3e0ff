/*
 * exp_oracle.cc -- CPU oracle for EXP's basis-function-expansion force loop
 * (sphereSL part, leapfrog, multistep bookkeeping).
 *
 * TEST INFRASTRUCTURE ONLY (see exp_oracle.h).  PARITY PINNED: tests/test_oracle_ref.py holds
 * the recursions, the SLGridSph maps and table interpolation, the coefficient and force
 * threads, the leapfrog and the level assignment bit for bit to the reference's own member
 * functions compiled in place (oracle/ref_extract.py -> oracle/_ref/libexpref_path.so).
 *
 * Written from the behaviour of the reference sources; every function names
 * the reference file:line it restates.  Arithmetic is kept in the reference's
 * evaluation order (left-to-right products, the same loop nests) and the file
 * is compiled with -ffp-contract=off so that no FMA contraction differs from a
 * portable x86-64 build of the reference.
 */
#include "exp_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

namespace {

// src/expand.H:130
constexpr double DSMALL = 1.0e-16;
// src/Basis.cc:7
constexpr double MINEPS = 3.0 * std::numeric_limits<double>::epsilon();
// exputil/SLGridMP2.cc:10
constexpr double XOFFSET = 1.0e-8;

inline int LM(int lmax, int l, int m) { return l * (lmax + 1) + m; }

// exputil/biorth_grid.cc:183-189 (plain recursive n!)
double factrl(int n) { return n <= 1 ? 1.0 : n * factrl(n - 1); }

template <class F>
void fork_join(int nthreads, F &&body)
{
  // Mirrors PotAccel::exp_thread_fork (src/PotAccel.cc:97-130): nthrds workers,
  // joined before returning.
  if (nthreads <= 1) { body(0); return; }
  std::vector<std::thread> pool;
  for (int id = 0; id < nthreads; id++) pool.emplace_back([&body, id] { body(id); });
  for (auto &t : pool) t.join();
}

struct ParticleView {
  const orc_particles *S;
  // Component::Pos(i, j, Local|Centered)  (src/Component.H:747-757)
  double pos(int i, int k) const { return S->pos[3 * (long)i + k] - S->center[k]; }
  // Component::freeze  (src/Component.cc:4194-4202)
  bool freeze(int i) const {
    double r2 = 0.0;
    for (int k = 0; k < 3; k++) {
      double d = S->pos[3 * (long)i + k] - S->center[k];
      r2 += d * d;
    }
    return r2 > S->rtrunc * S->rtrunc;
  }
};

} // namespace

extern "C" {

// ---------------------------------------------------------------------------
// a1: Basis::legendre_R  (src/Basis.cc:14-52)
// ---------------------------------------------------------------------------
void orc_legendre_R(int lmax, double x, double *p)
{
  double pll = 1.0;
  p[LM(lmax, 0, 0)] = 1.0;
  if (lmax > 0) {
    double somx2 = std::sqrt((1.0 - x) * (1.0 + x));
    double fact = 1.0;
    for (int m = 1; m <= lmax; m++) {
      pll *= -fact * somx2;
      p[LM(lmax, m, m)] = pll;
      fact += 2.0;
    }
  }
  for (int m = 0; m < lmax; m++) {
    double pl2 = p[LM(lmax, m, m)];
    double pl1 = x * (2 * m + 1) * pl2;
    p[LM(lmax, m + 1, m)] = pl1;
    for (int l = m + 2; l <= lmax; l++) {
      double v = (x * (2 * l - 1) * pl1 - (l + m - 1) * pl2) / (l - m);
      p[LM(lmax, l, m)] = v;
      pl2 = pl1;
      pl1 = v;
    }
  }
}

// ---------------------------------------------------------------------------
// a2: Basis::dlegendre_R  (src/Basis.cc:54-93)
// ---------------------------------------------------------------------------
void orc_dlegendre_R(int lmax, double x, double *p, double *dp)
{
  orc_legendre_R(lmax, x, p);

  if (1.0 - std::fabs(x) < MINEPS) {
    if (x > 0) x = 1.0 - MINEPS;
    else       x = -(1.0 - MINEPS);
  }
  double somx2 = 1.0 / (x * x - 1.0);
  dp[LM(lmax, 0, 0)] = 0.0;
  for (int l = 1; l <= lmax; l++) {
    for (int m = 0; m < l; m++)
      dp[LM(lmax, l, m)] =
          somx2 * (x * l * p[LM(lmax, l, m)] - (l + m) * p[LM(lmax, l - 1, m)]);
    dp[LM(lmax, l, l)] = somx2 * x * l * p[LM(lmax, l, l)];
  }
}

// ---------------------------------------------------------------------------
// a3: Basis::sinecosine_R  (src/Basis.cc:95-111)
// ---------------------------------------------------------------------------
void orc_sinecosine_R(int mmax, double phi, double *c, double *s)
{
  c[0] = 1.0;
  s[0] = 0.0;
  if (mmax > 0) {
    c[1] = std::cos(phi);
    s[1] = std::sin(phi);
    for (int m = 2; m <= mmax; m++) {
      c[m] = 2.0 * c[1] * c[m - 1] - c[m - 2];
      s[m] = 2.0 * c[1] * s[m - 1] - s[m - 2];
    }
  }
}

// ---------------------------------------------------------------------------
// a4: normalisation table entry  (src/SphericalBasis.cc:330-335)
// ---------------------------------------------------------------------------
double orc_factorial(int l, int m)
{
  double f = std::sqrt((2.0 * l + 1.0) / (4.0 * M_PI) * factrl(l - m) / factrl(l + m));
  if (m) f *= M_SQRT2;
  return f;
}

// ---------------------------------------------------------------------------
// a5: SLGridSph grid + coordinate maps
//     (exputil/SLGridMP2.cc:711-765 maps, :1355-1382 init_table)
// ---------------------------------------------------------------------------
double orc_r_to_xi(const orc_slgrid *g, double r)
{
  if (g->cmap == 1) return (r / g->rmap - 1.0) / (r / g->rmap + 1.0);
  if (g->cmap == 2) return std::log(r);
  return r;
}

double orc_xi_to_r(const orc_slgrid *g, double xi)
{
  if (g->cmap == 1) return (1.0 + xi) / (1.0 - xi) * g->rmap;
  if (g->cmap == 2) return std::exp(xi);
  return xi;
}

double orc_d_xi_to_r(const orc_slgrid *g, double xi)
{
  if (g->cmap == 1) return 0.5 * (1.0 - xi) * (1.0 - xi) / g->rmap;
  if (g->cmap == 2) return std::exp(-xi);
  return 1.0;
}

void orc_slgrid_init(orc_slgrid *g)
{
  if (g->cmap == 1) {
    g->xmin = (g->rmin / g->rmap - 1.0) / (g->rmin / g->rmap + 1.0);
    g->xmax = (g->rmax / g->rmap - 1.0) / (g->rmax / g->rmap + 1.0);
  } else if (g->cmap == 2) {
    g->xmin = std::log(g->rmin);
    g->xmax = std::log(g->rmax);
  } else {
    g->xmin = g->rmin;
    g->xmax = g->rmax;
  }
  g->dxi = (g->xmax - g->xmin) / (g->numr - 1);
}

namespace {
inline double grid_xi(const orc_slgrid *g, int i) { return g->xmin + g->dxi * i; }
inline double EF(const orc_slgrid *g, int l, int n, int i)
{
  return g->ef[((long)l * g->nmax + n) * g->numr + i];
}
inline double EV(const orc_slgrid *g, int l, int n) { return g->ev[l * g->nmax + n]; }
} // namespace

// ---------------------------------------------------------------------------
// a6: SLGridSph::get_pot(Matrix&, x, which=1)  (exputil/SLGridMP2.cc:872-910)
//     USE_TABLE is defined (exputil/SLGridMP2.cc:8).
// ---------------------------------------------------------------------------
void orc_slgrid_get_pot(const orc_slgrid *g, double r, double *mat)
{
  double x = orc_r_to_xi(g, r);
  int indx = (int)((x - g->xmin) / g->dxi);
  if (indx < 0) indx = 0;
  if (indx > g->numr - 2) indx = g->numr - 2;

  double x1 = (grid_xi(g, indx + 1) - x) / g->dxi;
  double x2 = (x - grid_xi(g, indx)) / g->dxi;

  for (int l = 0; l <= g->lmax; l++)
    for (int n = 0; n < g->nmax; n++)
      mat[l * g->nmax + n] =
          (x1 * EF(g, l, n, indx) + x2 * EF(g, l, n, indx + 1)) /
          std::sqrt(EV(g, l, n)) * (x1 * g->p0[indx] + x2 * g->p0[indx + 1]);
}

// SLGridSph::get_dens(Matrix&, x, which=1)  (exputil/SLGridMP2.cc:913-951)
void orc_slgrid_get_dens(const orc_slgrid *g, double r, double *mat)
{
  double x = orc_r_to_xi(g, r);
  int indx = (int)((x - g->xmin) / g->dxi);
  if (indx < 0) indx = 0;
  if (indx > g->numr - 2) indx = g->numr - 2;

  double x1 = (grid_xi(g, indx + 1) - x) / g->dxi;
  double x2 = (x - grid_xi(g, indx)) / g->dxi;

  for (int l = 0; l <= g->lmax; l++)
    for (int n = 0; n < g->nmax; n++)
      mat[l * g->nmax + n] =
          (x1 * EF(g, l, n, indx) + x2 * EF(g, l, n, indx + 1)) *
          std::sqrt(EV(g, l, n)) * (x1 * g->d0[indx] + x2 * g->d0[indx + 1]);
}

// ---------------------------------------------------------------------------
// a7: SLGridSph::get_force(Matrix&, x, which=1)  (exputil/SLGridMP2.cc:954-989)
// ---------------------------------------------------------------------------
void orc_slgrid_get_force(const orc_slgrid *g, double r, double *mat)
{
  double x = orc_r_to_xi(g, r);
  int indx = (int)((x - g->xmin) / g->dxi);
  if (indx < 1) indx = 1;
  if (indx > g->numr - 2) indx = g->numr - 2;

  double p = (x - grid_xi(g, indx)) / g->dxi;
  double fac = orc_d_xi_to_r(g, x) / g->dxi;

  for (int l = 0; l <= g->lmax; l++)
    for (int n = 0; n < g->nmax; n++)
      mat[l * g->nmax + n] =
          fac * ((p - 0.5) * EF(g, l, n, indx - 1) * g->p0[indx - 1]
                 - 2.0 * p * EF(g, l, n, indx) * g->p0[indx]
                 + (p + 0.5) * EF(g, l, n, indx + 1) * g->p0[indx + 1]) /
          std::sqrt(EV(g, l, n));
}

// ---------------------------------------------------------------------------
// a8/a10: SphericalBasis::determine_coefficients_thread
//         (src/SphericalBasis.cc:429-599) for one thread slice.
// ---------------------------------------------------------------------------
namespace {

struct CovOut {
  int sampT = 0;
  double *muse = nullptr, *massT = nullptr, *coefT = nullptr, *coefM = nullptr;
  unsigned *countT = nullptr;
};

long sph_coef_slice(const orc_slgrid *g, const orc_sph_params *P,
                    const orc_particles *S, const int *list, long nbeg, long nend,
                    double adb, double *coef, const CovOut *cov)
{
  const int Lmax = g->lmax, nmax = g->nmax;
  const double fac0 = -4.0 * M_PI;
  const int ldim = (Lmax + 1) * (Lmax + 2) / 2;
  ParticleView V{S};

  std::vector<double> legs((Lmax + 1) * (Lmax + 1)), cosm(Lmax + 1), sinm(Lmax + 1);
  std::vector<double> potd((Lmax + 1) * nmax), wk(nmax), factorial((Lmax + 1) * (Lmax + 1));
  for (int l = 0; l <= Lmax; l++)
    for (int m = 0; m <= l; m++) factorial[LM(Lmax, l, m)] = orc_factorial(l, m);

  long use = 0;
  unsigned whch = 0;

  for (long i = nbeg; i < nend; i++) {
    int indx = list[i];
    if (V.freeze(indx)) continue;

    double mass = S->mass[indx] * adb;
    double xx = V.pos(indx, 0), yy = V.pos(indx, 1), zz = V.pos(indx, 2);
    double r2 = (xx * xx + yy * yy + zz * zz);
    double r = std::sqrt(r2) + DSMALL;

    if (r >= P->rmin and r <= P->rmax) {
      use++;
      double costh = zz / r;
      double phi = std::atan2(yy, xx);
      double rs = r / P->scale;

      orc_legendre_R(Lmax, costh, legs.data());
      orc_sinecosine_R(Lmax, phi, cosm.data(), sinm.data());
      orc_slgrid_get_pot(g, rs, potd.data());       // Sphere::get_potl, src/Sphere.cc:147-150

      if (cov) {
        *cov->muse += mass;
        whch = (unsigned)indx % cov->sampT;
        cov->countT[whch] += 1;
        cov->massT[whch] += mass;
      }

      for (int l = 0, loffset = 0, iC = 0; l <= Lmax; loffset += (2 * l + 1), l++) {
        for (int m = 0, moffset = 0; m <= l; m++) {
          double facL = factorial[LM(Lmax, l, m)] * legs[LM(Lmax, l, m)];
          if (m == 0) {
            double *row = coef + (long)(loffset + moffset) * nmax;
            for (int n = 0; n < nmax; n++) {
              // sqnorm(l,n) == 1 for Sphere (src/AxisymmetricBasis.cc:110-112)
              wk[n] = potd[l * nmax + n] * facL * mass * fac0 / 1.0;
              row[n] += wk[n];
            }
            if (cov) {
              double *T = cov->coefT + (((long)whch * ldim + iC) * nmax) * 2;
              double *M = cov->coefM + ((long)whch * ldim + iC) * nmax * nmax;
              for (int n = 0; n < nmax; n++) {
                T[2 * n] += wk[n];
                for (int o = 0; o < nmax; o++) M[n * nmax + o] += wk[n] * wk[o] / mass;
              }
            }
            iC++;
            moffset++;
          } else {
            if (not P->M0_only) {
              double fac1 = facL * cosm[m];
              double fac2 = facL * sinm[m];
              double *rowc = coef + (long)(loffset + moffset) * nmax;
              double *rows = coef + (long)(loffset + moffset + 1) * nmax;
              for (int n = 0; n < nmax; n++) {
                wk[n] = potd[l * nmax + n] * mass * fac0 / 1.0;
                rowc[n] += wk[n] * fac1;
                rows[n] += wk[n] * fac2;
              }
              if (cov) {
                double *T = cov->coefT + (((long)whch * ldim + iC) * nmax) * 2;
                double *M = cov->coefM + ((long)whch * ldim + iC) * nmax * nmax;
                for (int n = 0; n < nmax; n++) {
                  T[2 * n]     += wk[n] * fac1;
                  T[2 * n + 1] += wk[n] * fac2;
                  for (int o = 0; o < nmax; o++)
                    M[n * nmax + o] += wk[n] * wk[o] * facL * facL / mass;
                }
              }
            }
            iC++;
            moffset += 2;
          }
        }
      }
    }
  }
  return use;
}

} // namespace

long orc_sph_coefficients(const orc_slgrid *g, const orc_sph_params *P,
                          const orc_particles *S, const int *list, long nlist,
                          double adb, int nthreads, double *coef)
{
  const long ncoef = (long)(g->lmax + 1) * (g->lmax + 1) * g->nmax;
  if (nthreads < 1) nthreads = 1;
  // expcoef0[id] per thread, zeroed (src/SphericalBasis.cc:794-796)
  std::vector<std::vector<double>> part(nthreads, std::vector<double>(ncoef, 0.0));
  std::vector<long> use(nthreads, 0);

  fork_join(nthreads, [&](int id) {
    long nbeg = nlist * id / nthreads;          // src/SphericalBasis.cc:437-440
    long nend = nlist * (id + 1) / nthreads;
    use[id] = sph_coef_slice(g, P, S, list, nbeg, nend, adb, part[id].data(), nullptr);
  });

  // thread sum in id order (src/SphericalBasis.cc:855-858)
  long use1 = 0;
  for (int id = 0; id < nthreads; id++) use1 += use[id];
  for (int id = 1; id < nthreads; id++)
    for (long k = 0; k < ncoef; k++) part[0][k] += part[id][k];
  std::memcpy(coef, part[0].data(), ncoef * sizeof(double));
  return use1;
}

long orc_sph_coefficients_cov(const orc_slgrid *g, const orc_sph_params *P,
                              const orc_particles *S, const int *list, long nlist,
                              double adb, int sampT, double *coef, double *muse,
                              double *massT, unsigned *countT,
                              double *coefT, double *coefM)
{
  // The reference serialises the covariance updates with a global mutex
  // (src/SphericalBasis.cc:505-533); the oracle runs them on one thread.
  const long ncoef = (long)(g->lmax + 1) * (g->lmax + 1) * g->nmax;
  std::fill(coef, coef + ncoef, 0.0);
  CovOut cov;
  cov.sampT = sampT; cov.muse = muse; cov.massT = massT; cov.countT = countT;
  cov.coefT = coefT; cov.coefM = coefM;
  return sph_coef_slice(g, P, S, list, 0, nlist, adb, coef, &cov);
}

// ---------------------------------------------------------------------------
// a11: SphericalBasis::determine_acceleration_and_potential_thread
//      (src/SphericalBasis.cc:1476-1660), get_pot_coefs_safe (:1797-1813),
//      Sphere::get_dpotl (src/Sphere.cc:140-145).
// ---------------------------------------------------------------------------
namespace {

inline void pot_coefs(int nmax, const double *coefrow, const double *potd_l,
                      const double *dpot_l, double &p, double &dp)
{
  double pp = 0.0, dpp = 0.0;
  for (int i = 0; i < nmax; i++) {
    pp  += potd_l[i] * coefrow[i];
    dpp += dpot_l[i] * coefrow[i];
  }
  p = pp;
  dp = dpp;
}

void sph_accel_slice(const orc_slgrid *g, const orc_sph_params *P,
                     const orc_particles *S, const int *list, long nbeg, long nend,
                     const double *coef)
{
  const int Lmax = g->lmax, nmax = g->nmax;
  const double scale = P->scale, rmax = P->rmax;
  ParticleView V{S};

  std::vector<double> legs((Lmax + 1) * (Lmax + 1)), dlegs((Lmax + 1) * (Lmax + 1));
  std::vector<double> cosm(Lmax + 1), sinm(Lmax + 1);
  std::vector<double> potd((Lmax + 1) * nmax), dpot((Lmax + 1) * nmax);
  std::vector<double> factorial((Lmax + 1) * (Lmax + 1));
  for (int l = 0; l <= Lmax; l++)
    for (int m = 0; m <= l; m++) factorial[LM(Lmax, l, m)] = orc_factorial(l, m);

  const double mfactor = 1.0;   // no MixtureBasis on this path

  for (long i = nbeg; i < nend; i++) {
    int indx = list[i];
    if (V.freeze(indx)) continue;

    double xx = V.pos(indx, 0), yy = V.pos(indx, 1), zz = V.pos(indx, 2);

    double r = std::sqrt(xx * xx + yy * yy + zz * zz) + DSMALL;
    double costh = zz / r;
    double rs = r / scale;
    double phi = std::atan2(yy, xx);
    double r0 = 0.0;

    orc_dlegendre_R(Lmax, costh, legs.data(), dlegs.data());
    orc_sinecosine_R(Lmax, phi, cosm.data(), sinm.data());

    int ioff = 0;
    if (r > rmax) {
      ioff = 1;
      r0 = r;
      r = rmax;
      rs = r / scale;
    }

    double potl = 0.0, potr = 0.0, pott = 0.0, potp = 0.0;
    double p, dp, pc, dpc, ps, dps;

    orc_slgrid_get_pot(g, rs, potd.data());
    orc_slgrid_get_force(g, rs, dpot.data());

    if (!P->NO_L0) {
      pot_coefs(nmax, coef, potd.data(), dpot.data(), p, dp);
      if (ioff) {
        p *= rmax / r0;
        dp = -p / r0;
      }
      double facL = mfactor * factorial[LM(Lmax, 0, 0)];
      potl = facL * p;
      potr = facL * dp;
    }

    for (int l = 1, loffset = 1; l <= Lmax; loffset += (2 * l + 1), l++) {
      if (P->NO_L1 && l == 1) continue;
      if (P->EVEN_L && (l / 2) * 2 != l) continue;

      for (int m = 0, moffset = 0; m <= l; m++) {
        double facL = factorial[LM(Lmax, l, m)] * legs[LM(Lmax, l, m)] * mfactor;
        double facD = factorial[LM(Lmax, l, m)] * dlegs[LM(Lmax, l, m)] * mfactor;

        // NB: the reference `continue`s here *before* advancing moffset
        // (src/SphericalBasis.cc:1590-1594), so with EVEN_M/M0_only the later
        // m terms read shifted rows.  The oracle keeps that behaviour.
        if (P->EVEN_M && (m / 2) * 2 != m) continue;
        if (P->M0_only and m != 0) continue;

        const double *pl = potd.data() + l * nmax;
        const double *dl = dpot.data() + l * nmax;

        if (m == 0) {
          pot_coefs(nmax, coef + (long)(loffset + moffset) * nmax, pl, dl, p, dp);
          if (ioff) {
            p *= std::pow(rmax / r0, (double)(l + 1));
            dp = -p / r0 * (l + 1);
          }
          potl += facL * p;
          potr += facL * dp;
          pott += facD * p;
          moffset++;
        } else {
          pot_coefs(nmax, coef + (long)(loffset + moffset) * nmax, pl, dl, pc, dpc);
          pot_coefs(nmax, coef + (long)(loffset + moffset + 1) * nmax, pl, dl, ps, dps);
          if (ioff) {
            double facp = std::pow(rmax / r0, (double)(l + 1));
            double facdp = -1.0 / r0 * (l + 1);
            pc *= facp;
            ps *= facp;
            dpc = pc * facdp;
            dps = ps * facdp;
          }
          potl += facL * (pc * cosm[m] + ps * sinm[m]);
          potr += facL * (dpc * cosm[m] + dps * sinm[m]);
          pott += facD * (pc * cosm[m] + ps * sinm[m]);
          potp += facL * (-pc * sinm[m] + ps * cosm[m]) * m;
          moffset += 2;
        }
      }
    }

    double fac = xx * xx + yy * yy;

    potr /= scale * scale;
    potl /= scale;
    pott /= scale;
    potp /= scale;

    // AddAcc / AddPot (src/Component.H:914-921); no pseudo-acceleration here.
    double *a = S->acc + 3 * (long)indx;
    a[0] += -(potr * xx / r - pott * xx * zz / (r * r * r));
    a[1] += -(potr * yy / r - pott * yy * zz / (r * r * r));
    a[2] += -(potr * zz / r + pott * fac / (r * r * r));
    if (fac > DSMALL) {
      a[0] +=  potp * yy / fac;
      a[1] += -potp * xx / fac;
    }
    S->pot[indx] += potl;
  }
}

} // namespace

void orc_sph_accel(const orc_slgrid *g, const orc_sph_params *P,
                   const orc_particles *S, const int *list, long nlist,
                   const double *coef, int nthreads)
{
  if (nthreads < 1) nthreads = 1;
  fork_join(nthreads, [&](int id) {
    long nbeg = nlist * id / nthreads;          // src/SphericalBasis.cc:1509-1510
    long nend = nlist * (id + 1) / nthreads;
    sph_accel_slice(g, P, S, list, nbeg, nend, coef);
  });
}

// ---------------------------------------------------------------------------
// a13: SphericalBasis::multistep_update  (src/SphericalBasis.cc:1156-1228)
// ---------------------------------------------------------------------------
void orc_sph_multistep_update(const orc_slgrid *g, const orc_sph_params *P,
                              const orc_particles *S, int indx, double adb,
                              int from, int to, double *differ)
{
  const int Lmax = g->lmax, nmax = g->nmax;
  const long nrow = (long)(Lmax + 1) * (Lmax + 1);
  ParticleView V{S};
  if (V.freeze(indx)) return;

  double mass = S->mass[indx] * adb;
  double xx = V.pos(indx, 0), yy = V.pos(indx, 1), zz = V.pos(indx, 2);
  double r2 = (xx * xx + yy * yy + zz * zz);
  double r = std::sqrt(r2) + DSMALL;

  if (r < P->rmax) {
    double costh = zz / r;
    double phi = std::atan2(yy, xx);
    double rs = r / P->scale;
    double fac0 = -4.0 * M_PI;

    std::vector<double> legs((Lmax + 1) * (Lmax + 1)), cosm(Lmax + 1), sinm(Lmax + 1);
    std::vector<double> potd((Lmax + 1) * nmax);
    orc_legendre_R(Lmax, costh, legs.data());
    orc_sinecosine_R(Lmax, phi, cosm.data(), sinm.data());
    orc_slgrid_get_pot(g, rs, potd.data());

    double *F = differ + (long)from * nrow * nmax;
    double *T = differ + (long)to * nrow * nmax;

    for (int l = 0, loffset = 0; l <= Lmax; loffset += (2 * l + 1), l++) {
      for (int m = 0, moffset = 0; m <= l; m++) {
        double facL = orc_factorial(l, m) * legs[LM(Lmax, l, m)];
        if (m == 0) {
          for (int n = 0; n < nmax; n++) {
            double val = potd[l * nmax + n] * facL * mass * fac0 / 1.0;
            F[(loffset + moffset) * nmax + n] -= val;
            T[(loffset + moffset) * nmax + n] += val;
          }
          moffset++;
        } else {
          double fac1 = facL * cosm[m];
          double fac2 = facL * sinm[m];
          for (int n = 0; n < nmax; n++) {
            double val1 = potd[l * nmax + n] * fac1 * mass * fac0 / 1.0;
            double val2 = potd[l * nmax + n] * fac2 * mass * fac0 / 1.0;
            F[(loffset + moffset) * nmax + n] -= val1;
            F[(loffset + moffset + 1) * nmax + n] -= val2;
            T[(loffset + moffset) * nmax + n] += val1;
            T[(loffset + moffset + 1) * nmax + n] += val2;
          }
          moffset += 2;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// a12: SphericalBasis::compute_multistep_coefficients
//      (src/SphericalBasis.cc:1231-1333); the cylinder analogue
//      (src/CylEXP.cc:192-282) has the same weights.
// ---------------------------------------------------------------------------
void orc_multistep_interp(int multistep, int mdrft, const int *mfirst,
                          const int *dstepL, const int *dstepN,
                          long ncoef, const double *coefL, const double *coefN,
                          double *out)
{
  const int Mstep = 1 << multistep;
  std::fill(out, out + ncoef, 0.0);

  for (int M = 0; M < mfirst[mdrft]; M++) {
    double numer = static_cast<double>(mdrft - dstepL[M * Mstep + mdrft]);
    double denom = static_cast<double>(dstepN[M * Mstep + mdrft] - dstepL[M * Mstep + mdrft]);
    double b = numer / denom;
    double a = 1.0 - b;
    for (long k = 0; k < ncoef; k++)
      out[k] += a * coefL[M * ncoef + k] + b * coefN[M * ncoef + k];
  }
  for (int M = mfirst[mdrft]; M <= multistep; M++)
    for (long k = 0; k < ncoef; k++) out[k] += coefN[M * ncoef + k];
}

// ---------------------------------------------------------------------------
// a21: leapfrog  (src/incpos.cc:63-66, src/incvel.cc:68-69)
// ---------------------------------------------------------------------------
void orc_incr_position(orc_particles *S, const int *list, long nlist, double dt, int dim)
{
  for (long q = 0; q < nlist; q++) {
    long i = list ? list[q] : q;
    for (int k = 0; k < dim; k++) S->pos[3 * i + k] += S->vel[3 * i + k] * dt;
  }
}

void orc_incr_velocity(orc_particles *S, const int *list, long nlist, double dt, int dim)
{
  for (long q = 0; q < nlist; q++) {
    long i = list ? list[q] : q;
    for (int k = 0; k < dim; k++) S->vel[3 * i + k] += S->acc[3 * i + k] * dt;
  }
}

// ---------------------------------------------------------------------------
// f1: Component::fix_positions_cpu / fix_positions_thread  (src/Component.cc:3280-3590)
// without the consp/tidal escape branch (:3314-3334) and the EJ/c0/CF centre overrides
// (:3570-3588), which are host-side policy.
//   lev_sums[M][10] = {com_mas[M], com_lev[3M..], cov_lev[3M..], coa_lev[3M..]} is the
//   state the reference keeps between calls: rows M >= mlevel are zeroed and re-summed
//   (:3364-3368, :3310-3349), rows below keep their old values, then all levels are
//   added (:3484-3497) and divided by the total mass (:3540-3544).
//   out[10] = {mtot, com[3], cov[3], coa[3]}.
// ---------------------------------------------------------------------------
void orc_fix_positions(const orc_particles *S, int mlevel, int multistep, int dim,
                       double *lev_sums, double *out)
{
  for (int mm = mlevel; mm <= multistep; mm++)
    for (int r = 0; r < 10; r++) lev_sums[10 * mm + r] = 0.0;
  for (int mm = mlevel; mm <= multistep; mm++) {
    for (long i = 0; i < S->n; i++) {                   // levlist[mm] in index order
      if ((int)S->level[i] != mm) continue;
      double r2 = 0.0;                                   // Component::freeze, :4194-4202
      for (int k = 0; k < 3; k++) {
        double d = S->pos[3 * i + k] - S->center[k];
        r2 += d * d;
      }
      if (r2 > S->rtrunc * S->rtrunc) continue;
      double *L = lev_sums + 10 * mm;
      L[0] += S->mass[i];
      for (int k = 0; k < dim; k++) {
        L[1 + k] += S->mass[i] * S->pos[3 * i + k];
        L[4 + k] += S->mass[i] * S->vel[3 * i + k];
        L[7 + k] += S->mass[i] * S->acc[3 * i + k];
      }
    }
  }
  for (int r = 0; r < 10; r++) out[r] = 0.0;
  for (int mm = 0; mm <= multistep; mm++)
    for (int r = 0; r < 10; r++) out[r] += lev_sums[10 * mm + r];
  if (out[0] > 0.0)
    for (int r = 1; r < 10; r++) out[r] /= out[0];
}

// ---------------------------------------------------------------------------
// a22: adjust_multistep_level_thread  (src/multistep.cc:52-236)
// ---------------------------------------------------------------------------
long orc_adjust_levels(orc_particles *S, const int *list, long nlist,
                       const orc_level_params *P, int *which,
                       long *offlo_out, long *offhi_out)
{
  const double eps = 1.0e-10;
  long offlo = 0, offhi = 0, numsw = 0;

  for (long i = 0; i < nlist; i++) {
    long n = list ? list[i] : i;

    double dtr = 0.0, vtot = 0.0, atot = 0.0;
    for (int k = 0; k < P->dim; k++) {
      dtr  += S->vel[3 * n + k] * S->acc[3 * n + k];
      vtot += S->vel[3 * n + k] * S->vel[3 * n + k];
      atot += S->acc[3 * n + k] * S->acc[3 * n + k];
    }
    double ptot = std::fabs(S->pot[n] + S->potext[n]);

    double dsr = S->scale ? S->scale[n] : -1.0;
    double dts;
    if (dsr > 0) dts = P->dynfracS * dsr / std::fabs(std::sqrt(vtot) + eps);
    else         dts = 1.0 / eps;

    double dtd = P->dynfracD * 1.0 / std::sqrt(vtot + eps);
    double dtv = P->dynfracV * std::sqrt(vtot / (atot + eps));
    double dta = P->dynfracA * ptot / (std::fabs(dtr) + eps);
    double dtA = P->dynfracP * std::sqrt(ptot / (atot + eps));

    // std::map<double,int> dseq: the smallest key wins; on equal keys the
    // later insertion overwrites the index (src/multistep.cc:119-130).
    double best = dtd; int ibest = 0;
    auto offer = [&](double v, int id) {
      if (v < best) { best = v; ibest = id; }
      else if (v == best) { ibest = id; }
    };
    offer(dtv, 1);
    offer(dts, 2);
    if (dta > 0.0) offer(dta, 3);
    if (dtA > 0.0) offer(dtA, 4);

    double dt = std::max<double>(eps, best);

    if (P->noswitch) {
      if (P->reset_dtreq) S->dtreq[n] = std::numeric_limits<double>::max();
      if (dt < S->dtreq[n]) S->dtreq[n] = dt;
    } else {
      S->dtreq[n] = dt;                     // double -> float (include/Particle.H:58)
    }

    if (P->apply) {
      unsigned plev = S->level[n];
      unsigned nlev = plev;

      if (S->dtreq[n] > P->dtime) {
        nlev = 0;
        offhi++;
      } else {
        nlev = (int)std::floor(std::log(P->dtime / S->dtreq[n]) / std::log(2.0));
      }

      if (P->shiftlevl) {
        if (nlev > plev) {
          if (nlev - plev > (unsigned)P->shiftlevl) nlev = plev + P->shiftlevl;
        } else if (plev > nlev) {
          if (plev - nlev > (unsigned)P->shiftlevl) nlev = plev - P->shiftlevl;
        }
      }

      if (nlev > (unsigned)P->multistep) {
        nlev = P->multistep;
        offlo++;
      }

      nlev = std::max<int>(nlev, P->mfirst_mdrft);

      if (plev != nlev) {
        S->level[n] = nlev;
        numsw++;
      }
    }
    if (which) which[n] = ibest;
  }

  if (offlo_out) *offlo_out = offlo;
  if (offhi_out) *offhi_out = offhi;
  return numsw;
}

// ---------------------------------------------------------------------------
// a23: initialize_multistep  (src/multistep.cc:630-679)
// ---------------------------------------------------------------------------
void orc_multistep_tables(int multistep, int *mintvl, int *mfirst,
                          int *dstepL, int *dstepN)
{
  const int Mstep = 1 << multistep;

  mintvl[0] = Mstep;
  for (int n = 1; n <= multistep; n++) mintvl[n] = mintvl[n - 1] / 2;

  for (int ms = 0; ms <= Mstep; ms++) {
    for (int mlevel = 0; mlevel <= multistep; mlevel++) {
      bool active = (ms == 0) || ((ms % (1 << (multistep - mlevel))) == 0);
      if (active) { mfirst[ms] = mlevel; break; }
    }
  }

  for (int ms = 0; ms <= multistep; ms++) {
    int rev = multistep - ms;
    int d = 1 << ms;
    for (int n = 0; n < Mstep; n++) {
      dstepL[rev * Mstep + n] = (n / d) * d;
      dstepN[rev * Mstep + n] = dstepL[rev * Mstep + n] + d;
    }
  }
}

} // extern "C"
