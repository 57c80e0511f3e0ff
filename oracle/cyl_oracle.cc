/*
 * cyl_oracle.cc -- CPU oracle for the cylinder force (Cylinder + EmpCylSL tables).
 *
 * TEST INFRASTRUCTURE ONLY (see exp_oracle.h).  PARITY PINNED: tests/test_oracle_ref.py holds
 * these functions bit for bit to the reference's own EmpCylSL::r_to_xi / z_to_y / get_pot /
 * accumulate / accumulated_eval and Cylinder::determine_coefficients_thread /
 * determine_acceleration_and_potential_thread, cut from exputil/EmpCylSL.cc and src/Cylinder.cc and compiled in place
 * (oracle/ref_extract.py -> oracle/_ref/libexpref_path.so; the covar branch to 1e-14).
 *
 * Restates, in the reference's evaluation order:
 *   EmpCylSL::r_to_xi / z_to_y           exputil/EmpCylSL.cc:6446-6463, 7109-7117
 *   EmpCylSL::setup_table (grid consts)  exputil/EmpCylSL.cc:2123-2137
 *   EmpCylSL::get_pot                    exputil/EmpCylSL.cc:5557-5632
 *   EmpCylSL::accumulate (+covar)        exputil/EmpCylSL.cc:4049-4146
 *   EmpCylSL::accumulated_eval           exputil/EmpCylSL.cc:5256-5410
 *   Cylinder::determine_coefficients_thread              src/Cylinder.cc:823-890
 *   Cylinder::determine_acceleration_and_potential_thread src/Cylinder.cc:1266-1446
 */
#include "exp_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

namespace {

constexpr double DSMALL = 1.0e-16;     // src/expand.H:130

template <class F>
void fork_join(int nthreads, F &&body)
{
  if (nthreads <= 1) { body(0); return; }
  std::vector<std::thread> pool;
  for (int id = 0; id < nthreads; id++) pool.emplace_back([&body, id] { body(id); });
  for (auto &t : pool) t.join();
}

inline double T(const double *tab, const orc_cylgrid *g, int m, int n, int ix, int iy)
{
  return tab[(((long)m * g->NORDER + n) * (g->NUMX + 1) + ix) * (g->NUMY + 1) + iy];
}

struct Cell { int ix, iy; double c00, c10, c01, c11; };

// the index / weight arithmetic shared by get_pot and accumulated_eval
// (enforce_limits is false: exputil/EmpCylSL.cc:56)
inline Cell locate(const orc_cylgrid *g, double r, double z)
{
  double X = (orc_cyl_r_to_xi(g, r) - g->XMIN) / g->dX;
  double Y = (orc_cyl_z_to_y(g, z) - g->YMIN) / g->dY;
  int ix = (int)X, iy = (int)Y;
  if (ix < 0) ix = 0;
  if (iy < 0) iy = 0;
  if (ix >= g->NUMX) ix = g->NUMX - 1;
  if (iy >= g->NUMY) iy = g->NUMY - 1;
  double delx0 = (double)ix + 1.0 - X, dely0 = (double)iy + 1.0 - Y;
  double delx1 = X - (double)ix, dely1 = Y - (double)iy;
  return Cell{ix, iy, delx0 * dely0, delx1 * dely0, delx0 * dely1, delx1 * dely1};
}

inline double bil(const double *tab, const orc_cylgrid *g, int m, int n, const Cell &c)
{
  return T(tab, g, m, n, c.ix, c.iy) * c.c00 + T(tab, g, m, n, c.ix + 1, c.iy) * c.c10 +
         T(tab, g, m, n, c.ix, c.iy + 1) * c.c01 + T(tab, g, m, n, c.ix + 1, c.iy + 1) * c.c11;
}

inline void rotate(const double *R, double *v)
{
  if (!R) return;
  double a = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
  double b = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
  double c = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
  v[0] = a; v[1] = b; v[2] = c;
}

struct CovOut {
  int sampT = 0;
  double *massT = nullptr, *VC = nullptr, *MV = nullptr;
  unsigned *numbT = nullptr;
};

long cyl_coef_slice(const orc_cylgrid *g, const orc_cyl_params *P, const orc_particles *S,
                    const int *list, long nbeg, long nend, double adb,
                    double *cosN, double *sinN, double *cylmass0, const CovOut *cov)
{
  const int MMAX = g->MMAX, rank3 = g->NORDER;
  const double Rmax2 = P->rcylmax * P->rcylmax * P->acyl * P->acyl;
  std::vector<double> vc((std::max(1, MMAX) + 1) * rank3), vs((std::max(1, MMAX) + 1) * rank3);
  long use = 0;

  for (long i = nbeg; i < nend; i++) {
    int indx = list[i];
    // Component::freeze (src/Component.cc:4194-4202)
    double pos[3], r2f = 0.0;
    for (int k = 0; k < 3; k++) {
      pos[k] = S->pos[3 * (long)indx + k] - S->center[k];
      r2f += pos[k] * pos[k];
    }
    if (r2f > S->rtrunc * S->rtrunc) continue;
    rotate(P->body, pos);

    double xx = pos[0], yy = pos[1], zz = pos[2];
    double r2 = xx * xx + yy * yy;
    double r = std::sqrt(r2);
    double R2 = r2 + zz * zz;

    if (R2 < Rmax2) {
      double mas = S->mass[indx] * adb;
      double phi = std::atan2(yy, xx);

      // ---- EmpCylSL::accumulate (exputil/EmpCylSL.cc:4049-4146)
      double rr = std::sqrt(r * r + zz * zz);
      if (!(rr / g->ASCALE > g->Rtable)) {
        const double norm = -4.0 * M_PI;
        unsigned whch = 0;
        if (cov) {
          whch = (unsigned long)indx % cov->sampT;
          cov->numbT[whch] += 1;
          cov->massT[whch] += mas;
        }
        orc_cyl_get_pot(g, r, zz, vc.data(), vs.data());
        for (int mm = 0; mm <= MMAX; mm++) {
          double mcos = std::cos(phi * mm);
          double msin = std::sin(phi * mm);
          for (int nn = 0; nn < rank3; nn++) {
            double hold = norm * mas * mcos * vc[mm * rank3 + nn];
            cosN[mm * rank3 + nn] += hold;
            if (mm > 0) {
              hold = norm * mas * msin * vs[mm * rank3 + nn];
              sinN[mm * rank3 + nn] += hold;
            }
          }
          if (cov) {
            // vec = (vC cos + vS sin) + i (vC sin - vS cos), vC = norm*Vc row, vS = norm*Vs row (0 for m=0)
            std::vector<double> re(rank3), im(rank3);
            for (int nn = 0; nn < rank3; nn++) {
              double vC = vc[mm * rank3 + nn] * norm;
              double vS = mm ? vs[mm * rank3 + nn] * norm : 0.0;
              re[nn] = vC * mcos + vS * msin;
              im[nn] = vC * msin - vS * mcos;
            }
            double *VCp = cov->VC + (((long)whch * (MMAX + 1) + mm) * rank3) * 2;
            double *MVp = cov->MV + (((long)whch * (MMAX + 1) + mm) * rank3 * rank3) * 2;
            for (int nn = 0; nn < rank3; nn++) {
              VCp[2 * nn] += mas * re[nn];
              VCp[2 * nn + 1] += mas * im[nn];
              for (int oo = 0; oo < rank3; oo++) {
                // vec * vec^H: (re_n + i im_n)(re_o - i im_o)
                MVp[2 * (nn * rank3 + oo)]     += mas * (re[nn] * re[oo] + im[nn] * im[oo]);
                MVp[2 * (nn * rank3 + oo) + 1] += mas * (im[nn] * re[oo] - re[nn] * im[oo]);
              }
            }
          }
        }
      }
      use++;
      *cylmass0 += mas;
    }
  }
  return use;
}

} // namespace

extern "C" {

double orc_cyl_r_to_xi(const orc_cylgrid *g, double r)
{
  if (g->CMAPR > 0) return (r / g->ASCALE - 1.0) / (r / g->ASCALE + 1.0);
  return r;
}

double orc_cyl_z_to_y(const orc_cylgrid *g, double z)
{
  if (g->CMAPZ == 1)
    return z / (std::fabs(z) + std::numeric_limits<double>::min()) * std::asinh(std::fabs(z / g->HSCALE));
  else if (g->CMAPZ == 2)
    return z / std::sqrt(z * z + g->HSCALE * g->HSCALE);
  return z;
}

// EmpCylSL::setup_table (exputil/EmpCylSL.cc:2130-2137)
void orc_cyl_init(orc_cylgrid *g)
{
  g->Rtable = M_SQRT1_2 * g->RMAX;
  g->XMIN = orc_cyl_r_to_xi(g, g->RMIN * g->ASCALE);
  g->XMAX = orc_cyl_r_to_xi(g, g->Rtable * g->ASCALE);
  g->dX = (g->XMAX - g->XMIN) / g->NUMX;
  g->YMIN = orc_cyl_z_to_y(g, -g->Rtable * g->ASCALE);
  g->YMAX = orc_cyl_z_to_y(g, g->Rtable * g->ASCALE);
  g->dY = (g->YMAX - g->YMIN) / g->NUMY;
}

void orc_cyl_get_pot(const orc_cylgrid *g, double r, double z, double *Vc, double *Vs)
{
  if (z / g->ASCALE > g->Rtable) z = g->Rtable * g->ASCALE;
  if (z / g->ASCALE < -g->Rtable) z = -g->Rtable * g->ASCALE;
  Cell c = locate(g, r, z);
  const int rank3 = g->NORDER;
  for (int mm = 0; mm <= std::min<int>(g->MLIM, g->MMAX); mm++) {
    if (g->EVEN_M && (mm / 2) * 2 != mm) continue;
    for (int n = 0; n < rank3; n++) {
      Vc[mm * rank3 + n] = 1.0 * bil(g->potC, g, mm, n, c);
      if (mm) Vs[mm * rank3 + n] = 1.0 * bil(g->potS, g, mm, n, c);
    }
  }
}

long orc_cyl_coefficients(const orc_cylgrid *g, const orc_cyl_params *P, const orc_particles *S,
                          const int *list, long nlist, double adb, int nthreads,
                          double *cosN, double *sinN, double *cylmass)
{
  const long nc = (long)(g->MMAX + 1) * g->NORDER;
  if (nthreads < 1) nthreads = 1;
  std::vector<std::vector<double>> pc(nthreads, std::vector<double>(nc, 0.0)), ps = pc;
  std::vector<long> use(nthreads, 0);
  std::vector<double> mass0(nthreads, 0.0);
  fork_join(nthreads, [&](int id) {
    long nbeg = nlist * id / nthreads, nend = nlist * (id + 1) / nthreads;   // src/Cylinder.cc:829-830
    use[id] = cyl_coef_slice(g, P, S, list, nbeg, nend, adb, pc[id].data(), ps[id].data(), &mass0[id], nullptr);
  });
  long use1 = 0;
  double m1 = 0.0;
  for (int id = 0; id < nthreads; id++) { use1 += use[id]; m1 += mass0[id]; }
  // thread sum (exputil/EmpCylSL.cc:4175-4190)
  for (int id = 1; id < nthreads; id++)
    for (long k = 0; k < nc; k++) { pc[0][k] += pc[id][k]; ps[0][k] += ps[id][k]; }
  std::memcpy(cosN, pc[0].data(), nc * sizeof(double));
  std::memcpy(sinN, ps[0].data(), nc * sizeof(double));
  *cylmass = m1;
  return use1;
}

long orc_cyl_coefficients_cov(const orc_cylgrid *g, const orc_cyl_params *P, const orc_particles *S,
                              const int *list, long nlist, double adb, int sampT,
                              double *cosN, double *sinN, double *cylmass,
                              double *massT, unsigned *numbT, double *VC, double *MV)
{
  const long nc = (long)(g->MMAX + 1) * g->NORDER;
  std::fill(cosN, cosN + nc, 0.0);
  std::fill(sinN, sinN + nc, 0.0);
  *cylmass = 0.0;
  CovOut cov;
  cov.sampT = sampT; cov.massT = massT; cov.numbT = numbT; cov.VC = VC; cov.MV = MV;
  return cyl_coef_slice(g, P, S, list, 0, nlist, adb, cosN, sinN, cylmass, &cov);
}

// Cylinder::multistep_update (src/Cylinder.cc:1751-1773) + CylEXP::multistep_update
// (src/CylEXP.cc:159-188) for one particle: its contribution moves from level `from` to
// level `to` in differC/differS [(multistep+1)][(MMAX+1)][NORDER].  Note: no orientation
// rotation and no rcylmax gate here, unlike the coefficient pass.
void orc_cyl_multistep_update(const orc_cylgrid *g, const orc_particles *S, int indx, double adb,
                              int from, int to, double *differC, double *differS)
{
  const int MMAX = g->MMAX, rank3 = g->NORDER;
  double pos[3], r2f = 0.0;
  for (int k = 0; k < 3; k++) {
    pos[k] = S->pos[3 * (long)indx + k] - S->center[k];
    r2f += pos[k] * pos[k];
  }
  if (r2f > S->rtrunc * S->rtrunc) return;              // c->freeze(i)
  double mass = S->mass[indx] * adb;
  double xx = pos[0], yy = pos[1], zz = pos[2];
  double r = std::sqrt(xx * xx + yy * yy);
  double phi = std::atan2(yy, xx);
  double rr = std::sqrt(r * r + zz * zz);
  if (rr / g->ASCALE > g->Rtable) return;
  const double norm = -4.0 * M_PI;
  std::vector<double> vc((std::max(1, MMAX) + 1) * rank3), vs((std::max(1, MMAX) + 1) * rank3);
  orc_cyl_get_pot(g, r, zz, vc.data(), vs.data());
  const long lev = (long)(MMAX + 1) * rank3;
  for (int mm = 0; mm <= MMAX; mm++) {
    double mcos = std::cos(phi * mm), msin = std::sin(phi * mm);
    for (int nn = 0; nn < rank3; nn++) {
      double hold = norm * mass * mcos * vc[mm * rank3 + nn];
      differC[from * lev + mm * rank3 + nn] -= hold;
      differC[to * lev + mm * rank3 + nn] += hold;
      if (mm > 0) {
        hold = norm * mass * msin * vs[mm * rank3 + nn];
        differS[from * lev + mm * rank3 + nn] -= hold;
        differS[to * lev + mm * rank3 + nn] += hold;
      }
    }
  }
}

void orc_cyl_accumulated_eval(const orc_cylgrid *g, const double *accum_cos, const double *accum_sin,
                              double r, double z, double phi,
                              double *p0, double *p_, double *fr_, double *fz_, double *fp_)
{
  double fr = 0.0, fz = 0.0, fp = 0.0, p = 0.0;
  *p_ = *fr_ = *fz_ = *fp_ = 0.0;
  double rr = std::sqrt(r * r + z * z);
  if (rr / g->ASCALE > g->Rtable) return;
  Cell c = locate(g, r, z);
  const int rank3 = g->NORDER;
  for (int mm = std::max<int>(0, g->MMIN); mm <= std::min<int>(g->MLIM, g->MMAX); mm++) {
    if (g->EVEN_M && (mm / 2) * 2 != mm) continue;
    double ccos = std::cos(phi * mm), ssin = std::sin(phi * mm);
    for (int n = std::max<int>(0, g->NMIN); n < std::min<int>(g->NLIM, rank3); n++) {
      double fac = accum_cos[mm * rank3 + n] * ccos;
      p  += fac * bil(g->potC, g, mm, n, c);
      fr += fac * bil(g->rforceC, g, mm, n, c);
      fz += fac * bil(g->zforceC, g, mm, n, c);
      fac = accum_cos[mm * rank3 + n] * ssin;
      fp += fac * mm * bil(g->potC, g, mm, n, c);
      if (mm) {
        fac = accum_sin[mm * rank3 + n] * ssin;
        p  += fac * bil(g->potS, g, mm, n, c);
        fr += fac * bil(g->rforceS, g, mm, n, c);
        fz += fac * bil(g->zforceS, g, mm, n, c);
        fac = -accum_sin[mm * rank3 + n] * ccos;
        fp += fac * mm * bil(g->potS, g, mm, n, c);
      }
    }
    if (mm == 0) *p0 = p;
  }
  *p_ = p; *fr_ = fr; *fz_ = fz; *fp_ = fp;
}

void orc_cyl_accel(const orc_cylgrid *g, const orc_cyl_params *P, const orc_particles *S,
                   const int *list, long nlist, const double *accum_cos, const double *accum_sin,
                   double cylmass, int nthreads)
{
  constexpr double ratmin = 0.75;
  constexpr double maxerf = 3.0;
  constexpr double midpt = ratmin + 0.5 * (1.0 - ratmin);
  constexpr double rsmth = 0.5 * (1.0 - ratmin) / maxerf;
  double R2 = g->ASCALE * g->Rtable;
  R2 = R2 * R2;
  if (nthreads < 1) nthreads = 1;

  fork_join(nthreads, [&](int id) {
    long nbeg = nlist * id / nthreads, nend = nlist * (id + 1) / nthreads;
    for (long q = nbeg; q < nend; q++) {
      int indx = list[q];
      double pos[3], frc[3] = {0.0, 0.0, 0.0};
      for (int k = 0; k < 3; k++) pos[k] = S->pos[3 * (long)indx + k] - S->center[k];
      rotate(P->body, pos);
      double xx = pos[0], yy = pos[1], zz = pos[2];
      double r2 = xx * xx + yy * yy;
      double r = std::sqrt(r2) + DSMALL;
      double phi = std::atan2(yy, xx);
      double pa = 0.0, frac, cfrac;
      const double mfactor = 1.0;

      double ratio = std::sqrt((r2 + zz * zz) / R2);
      if (ratio >= 1.0) {
        frac = 0.0;
        cfrac = 1.0;
      } else if (ratio > ratmin) {
        frac = 0.5 * (1.0 - std::erf((ratio - midpt) / rsmth));
        cfrac = 1.0 - frac;
      } else {
        cfrac = 0.0;
        frac = 1.0;
      }
      cfrac *= mfactor;
      frac *= mfactor;

      if (ratio < 1.0) {
        double p0, p, fr, fz, fp;
        orc_cyl_accumulated_eval(g, accum_cos, accum_sin, r, zz, phi, &p0, &p, &fr, &fz, &fp);
        frc[0] = (fr * xx / r - fp * yy / r2) * frac;
        frc[1] = (fr * yy / r + fp * xx / r2) * frac;
        frc[2] = fz * frac;
        pa = p * frac;
      }
      if (ratio > ratmin) {
        double r3 = r2 + zz * zz;
        double p = -cylmass / std::sqrt(r3);
        double fr = p / r3;
        frc[0] += xx * fr * cfrac;
        frc[1] += yy * fr * cfrac;
        frc[2] += zz * fr * cfrac;
        pa += p * cfrac;
      }
      S->pot[indx] += pa;
      rotate(P->orig, frc);
      for (int j = 0; j < 3; j++) S->acc[3 * (long)indx + j] += frc[j];
    }
  });
}

} // extern "C"
