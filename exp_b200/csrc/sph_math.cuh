// sph_math.cuh -- per-particle arithmetic of the sphereSL path, shared by the
// CUDA kernels (sphere.cu) and by the host emulation used in CPU-side tests
// (tests/emul/sph_emul.cu).  Everything here is __host__ __device__ and free of
// warp-level choreography.
//
// The factorised evaluation ("radial particle-mesh")
// -------------------------------------------------
// The reference evaluates, per particle, potd(l,n) = P0(xi) * (x1 E[l,n,i] + x2 E[l,n,i+1])
// with E = ef/sqrt(ev) (exputil/SLGridMP2.cc:872-910) and then loops over n for every
// (l,m) (src/SphericalBasis.cc:513-590, 1602-1634).  Because potd is linear in the table
// nodes, the particle sum and the n sum commute:
//
//   coef[row][n] = sum_p Y_p[row] w_p potd_p(l,n)
//                = sum_i E[l,n,i] * G[i][row],   G[i][row] = sum_p Y_p[row] w_p P0_p (x1_p [i_p=i] + x2_p [i_p=i-1])
//
//   sum_n coef[row][n] potd(l,n) = P0 (x1 H[i][row] + x2 H[i+1][row]),  H[i][row] = sum_n E[l,n,i] coef[row][n]
//   sum_n coef[row][n] dpot(l,n) = fac ((p-1/2) p0[j-1] H[j-1] - 2p p0[j] H[j] + (p+1/2) p0[j+1] H[j+1])
//
// so a particle costs O((L+1)^2) instead of O((L+1)^2 nmax) and the nmax-contractions
// become two tiny dense products per step.  The result differs from the reference only by
// floating-point re-association (~1e-15 relative).
#pragma once

#include "common.cuh"

#if defined(__CUDACC__)
#define HD __host__ __device__ __forceinline__
#else
#define HD inline
#endif

#if defined(__CUDA_ARCH__)
#define EXPB200_LDG(p) __ldg(p)
#else
#define EXPB200_LDG(p) (*(p))
#endif

namespace expb200 {

#if !defined(__CUDACC__)
using std::exp; using std::fabs; using std::log; using std::sqrt;
#endif

// Geometry + grid constants of one Sphere, passed to kernels by value.
struct SphGeom {
  int lmax, nmax, numr, cmap;
  int nrow;                 // (lmax+1)^2
  int nslot;                // (lmax+1)(lmax+2)/2
  double rmap, xmin, dxi;   // SLGridSph grid (exputil/SLGridMP2.cc:1355-1382)
  double scale, rmin, rmax; // SphericalBasis scale and expansion radii
  double cx, cy, cz, rtrunc2;
};

// slot of (l,m) in the m-major order used by the field table: m outer, l inner
HD int sph_slot(int lmax, int l, int m) { return m * (lmax + 1) - m * (m - 1) / 2 + (l - m); }
// row of (l,m) in the reference's coefficient layout (cos row; sin row is +1 for m>0)
HD int sph_row(int l, int m) { return l * l + (m ? 2 * m - 1 : 0); }

// a5: SLGridSph::r_to_xi / d_xi_to_r (exputil/SLGridMP2.cc:711-726, 748-765)
HD double sph_r_to_xi(const SphGeom &g, double r)
{
  if (g.cmap == 1) return (r / g.rmap - 1.0) / (r / g.rmap + 1.0);
  if (g.cmap == 2) return log(r);
  return r;
}
HD double sph_d_xi_to_r(const SphGeom &g, double xi)
{
  if (g.cmap == 1) return 0.5 * (1.0 - xi) * (1.0 - xi) / g.rmap;
  if (g.cmap == 2) return exp(-xi);
  return 1.0;
}

// Interpolation stencil of one radius: nodes j-1, j, j+1 (j = the reference's force
// index, clamped to [1,numr-2]); bw = weights of the potential (a6), aw = weights of
// the radial derivative (a7), both with p0 folded in.
struct SphStencil {
  int j;
  int cell;        // the reference's potential index (clamped to [0,numr-2])
  double x1, x2;   // linear weights on (cell, cell+1)
};

HD SphStencil sph_stencil(const SphGeom &g, double rs, double &xi_out)
{
  double xi = sph_r_to_xi(g, rs);
  int raw = (int)((xi - g.xmin) / g.dxi);
  int cell = raw < 0 ? 0 : raw;
  if (cell > g.numr - 2) cell = g.numr - 2;
  int j = raw < 1 ? 1 : raw;
  if (j > g.numr - 2) j = g.numr - 2;
  SphStencil s;
  s.j = j;
  s.cell = cell;
  s.x1 = ((g.xmin + g.dxi * (cell + 1)) - xi) / g.dxi;
  s.x2 = (xi - (g.xmin + g.dxi * cell)) / g.dxi;
  xi_out = xi;
  return s;
}

// Radial-cell key of a particle for the coefficient pass: its potential cell when it
// contributes (not frozen, rmin <= r <= rmax: src/SphericalBasis.cc:489,
// src/Component.cc:4194-4202), else the dead cell numr-1.
// mode 0: coefficient accumulation (rmin <= r <= rmax, src/SphericalBasis.cc:489);
// mode 1: multistep_update (r < rmax only, src/SphericalBasis.cc:1177).
HD bool sph_in_range(const SphGeom &g, double r, int mode)
{
  return mode == 0 ? (r >= g.rmin && r <= g.rmax) : (r < g.rmax);
}

HD uint32_t sph_cell_of(const SphGeom &g, const rec4 &pm, int mode = 0)
{
  double x = pm.x - g.cx, y = pm.y - g.cy, z = pm.z - g.cz;
  double r2 = x * x + y * y + z * z;
  if (r2 > g.rtrunc2) return (uint32_t)(g.numr - 1);
  double r = sqrt(r2) + EXPB200_DSMALL;
  if (!sph_in_range(g, r, mode)) return (uint32_t)(g.numr - 1);
  double xi;
  SphStencil s = sph_stencil(g, r / g.scale, xi);
  return (uint32_t)s.cell;
}

// Visits the associated Legendre functions in m-outer / l-inner order, calling
// f(l, m, P_l^m, P_{l-1}^m) with the recursion of Basis::legendre_R (src/Basis.cc:14-52);
// on_m(m) is called once at the head of every m column.  inv[k] = 1/k.
template <int L, class M, class F>
HD void sph_for_each_plm(double x, const double *inv, M &&on_m, F &&f)
{
  double somx2 = sqrt((1.0 - x) * (1.0 + x));
  double pmm = 1.0, fact = 1.0;
#pragma unroll
  for (int m = 0; m <= L; m++) {
    if (m > 0) {
      pmm *= -fact * somx2;
      fact += 2.0;
    }
    on_m(m);
    f(m, m, pmm, 0.0);
    if (m < L) {
      double pl2 = pmm;
      double pl1 = x * (2 * m + 1) * pl2;
      f(m + 1, m, pl1, pl2);
#pragma unroll
      for (int l = m + 2; l <= L; l++) {
        double pll = (x * (2 * l - 1) * pl1 - (l + m - 1) * pl2) * inv[l - m];
        f(l, m, pll, pl1);
        pl2 = pl1;
        pl1 = pll;
      }
    }
  }
}

// Azimuth: cos(phi), sin(phi) of phi = atan2(y,x) without the transcendental
// (src/SphericalBasis.cc:493 then src/Basis.cc:104-105).
HD void sph_azimuth(double x, double y, double &c1, double &s1, double &R2)
{
  R2 = x * x + y * y;
  if (R2 > 0.0) {
    double Ri = 1.0 / sqrt(R2);
    c1 = x * Ri;
    s1 = y * Ri;
  } else {
    c1 = 1.0;
    s1 = 0.0;
  }
}

// ---------------------------------------------------------------------------
// Coefficient pass, per particle (a8): fills yv[row] = K_lm P_lm {1,cos,sin}(m phi)
// and returns the two node weights w1 (node cell) and w2 (node cell+1), which carry
// mass * adb * (-4 pi) * P0(xi).  Returns false when the particle does not contribute.
// ---------------------------------------------------------------------------
template <int L, class Store>
HD bool sph_deposit_particle(const SphGeom &g, const double *K, const double *inv,
                             const double *p0, const rec4 &pm, double adb,
                             int &cell, double &w1, double &w2, Store &&store, int mode = 0)
{
  double x = pm.x - g.cx, y = pm.y - g.cy, z = pm.z - g.cz;
  double r2 = x * x + y * y + z * z;
  if (r2 > g.rtrunc2) return false;
  double r = sqrt(r2) + EXPB200_DSMALL;
  if (!sph_in_range(g, r, mode)) return false;

  double costh = z / r;
  double c1, s1, R2;
  sph_azimuth(x, y, c1, s1, R2);

  double xi;
  SphStencil st = sph_stencil(g, r / g.scale, xi);
  double P0 = st.x1 * EXPB200_LDG(p0 + st.cell) + st.x2 * EXPB200_LDG(p0 + st.cell + 1);
  double wgt = pm.w * adb * (-4.0 * M_PI) * P0;
  cell = st.cell;
  w1 = wgt * st.x1;
  w2 = wgt * st.x2;

  double cm = 1.0, sm = 0.0, cm1 = 0.0, sm1 = 0.0;   // cos/sin(m phi), (m-1) phi
  sph_for_each_plm<L>(costh, inv, [&](int m) {
    if (m == 0) return;       // advance the Chebyshev recursion (src/Basis.cc:107-110)
    double cn, sn;
    if (m == 1) { cn = c1; sn = s1; }
    else { cn = 2.0 * c1 * cm - cm1; sn = 2.0 * c1 * sm - sm1; }
    cm1 = cm; sm1 = sm; cm = cn; sm = sn;
  }, [&](int l, int m, double plm, double) {
    double f = K[l * (kMaxL + 1) + m] * plm;
    int row = sph_row(l, m);
    if (m == 0) store(row, f);
    else { store(row, f * cm); store(row + 1, f * sm); }
  });
  return true;
}

// ---------------------------------------------------------------------------
// Force pass, per particle (a11).  H is the field table [numr][nslot] of
// (cos,sin) pairs built from the coefficients (see sphere.cu: sph_field_kernel).
// Adds into (ax,ay,az,pot).  Returns false for frozen particles.
// ---------------------------------------------------------------------------
template <int L>
HD bool sph_force_particle(const SphGeom &g, const double *K, const double *inv,
                           const double *p0, const double2 *H, const rec4 &pm,
                           double &ax, double &ay, double &az, double &pot)
{
  double x = pm.x - g.cx, y = pm.y - g.cy, z = pm.z - g.cz;
  double r2 = x * x + y * y + z * z;
  if (r2 > g.rtrunc2) return false;

  double r = sqrt(r2) + EXPB200_DSMALL;
  double costh = z / r;
  double rs = r / g.scale;
  double c1, s1, R2;
  sph_azimuth(x, y, c1, s1, R2);

  bool ioff = r > g.rmax;
  double r0 = r;
  if (ioff) {                       // src/SphericalBasis.cc:1555-1560
    r = g.rmax;
    rs = r / g.scale;
  }

  double xi;
  SphStencil st = sph_stencil(g, rs, xi);
  const int j = st.j;
  double p0m = EXPB200_LDG(p0 + j - 1), p0c = EXPB200_LDG(p0 + j), p0p = EXPB200_LDG(p0 + j + 1);

  // potential weights on nodes j-1, j, j+1
  double b0, b1, b2;
  if (st.cell == j) {
    double P0 = st.x1 * p0c + st.x2 * p0p;
    b0 = 0.0; b1 = P0 * st.x1; b2 = P0 * st.x2;
  } else {                          // cell == j-1 (only at the inner edge)
    double P0 = st.x1 * p0m + st.x2 * p0c;
    b0 = P0 * st.x1; b1 = P0 * st.x2; b2 = 0.0;
  }
  // derivative weights (exputil/SLGridMP2.cc:975-986)
  double p = (xi - (g.xmin + g.dxi * j)) / g.dxi;
  double fac = sph_d_xi_to_r(g, xi) / g.dxi;
  double a0 = fac * (p - 0.5) * p0m, a1 = -2.0 * fac * p * p0c, a2 = fac * (p + 0.5) * p0p;

  // Legendre derivative clamp (src/Basis.cc:81-86)
  double xd = costh;
  if (1.0 - fabs(xd) < EXPB200_MINEPS) xd = xd > 0 ? 1.0 - EXPB200_MINEPS : -(1.0 - EXPB200_MINEPS);
  double somx2i = 1.0 / (xd * xd - 1.0);

  // exterior multipole continuation factors (src/SphericalBasis.cc:1605-1628)
  double ratio = ioff ? g.rmax / r0 : 1.0;
  double r0i = ioff ? 1.0 / r0 : 0.0;

  const double2 *Hm = H + (size_t)(j - 1) * g.nslot;
  const double2 *Hc = Hm + g.nslot;
  const double2 *Hp = Hc + g.nslot;

  double potl = 0.0, potr = 0.0, pott = 0.0, potp = 0.0;
  double cm = 1.0, sm = 0.0, cm1 = 0.0, sm1 = 0.0;
  double pw_m = 1.0;                // ratio^m; the column head needs ratio^(m+1)
  double pw = 1.0;
  int slot = 0;

  sph_for_each_plm<L>(costh, inv, [&](int m) {
    if (m > 0) {
      double cn, sn;
      if (m == 1) { cn = c1; sn = s1; }
      else { cn = 2.0 * c1 * cm - cm1; sn = 2.0 * c1 * sm - sm1; }
      cm1 = cm; sm1 = sm; cm = cn; sm = sn;
      pw_m *= ratio;
    }
    pw = pw_m;
  }, [&](int l, int m, double plm, double plm1) {
    pw *= ratio;                    // ratio^(l+1)
    double dplm = somx2i * (xd * l * plm - (l + m) * plm1);
    double k = K[l * (kMaxL + 1) + m];
    double facL = k * plm, facD = k * dplm;

    double2 v0 = EXPB200_LDG(Hm + slot), v1 = EXPB200_LDG(Hc + slot), v2 = EXPB200_LDG(Hp + slot);
    slot++;
    double pc = b0 * v0.x + b1 * v1.x + b2 * v2.x;
    double dpc = a0 * v0.x + a1 * v1.x + a2 * v2.x;
    if (ioff) {
      pc *= pw;
      dpc = -pc * r0i * (l + 1);
    }
    if (m == 0) {
      potl += facL * pc;
      potr += facL * dpc;
      pott += facD * pc;
    } else {
      double ps = b0 * v0.y + b1 * v1.y + b2 * v2.y;
      double dps = a0 * v0.y + a1 * v1.y + a2 * v2.y;
      if (ioff) {
        ps *= pw;
        dps = -ps * r0i * (l + 1);
      }
      double A = pc * cm + ps * sm;
      potl += facL * A;
      potr += facL * (dpc * cm + dps * sm);
      pott += facD * A;
      potp += facL * (-pc * sm + ps * cm) * m;
    }
  });

  // src/SphericalBasis.cc:1636-1653 (note: r is the clamped radius when r0 > rmax)
  potr /= g.scale * g.scale;
  potl /= g.scale;
  pott /= g.scale;
  potp /= g.scale;

  double r3 = r * r * r;
  ax += -(potr * x / r - pott * x * z / r3);
  ay += -(potr * y / r - pott * y * z / r3);
  az += -(potr * z / r + pott * R2 / r3);
  if (R2 > EXPB200_DSMALL) {
    ax += potp * y / R2;
    ay += -potp * x / R2;
  }
  pot += potl;
  return true;
}

} // namespace expb200
