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
//                = K_row sum_i E[l,n,i] * G[row][i],
//   G[row][i]    = sum_p P_p[row] w_p P0_p (x1_p [i_p=i] + x2_p [i_p=i-1])   (P = Y / K)
//
//   sum_n coef[row][n] potd(l,n) = P0 (x1 H[i][row] + x2 H[i+1][row]),  H[i][row] = sum_n E[l,n,i] coef[row][n]
//   sum_n coef[row][n] dpot(l,n) = fac ((q-1/2) p0[j-1] H[j-1] - 2q p0[j] H[j] + (q+1/2) p0[j+1] H[j+1])
//
// so a particle costs O((L+1)^2) instead of O((L+1)^2 nmax) and the nmax-contractions
// become two tiny dense products per step.  Within a radial cell both sums are
// polynomials in the cell coordinate (x2 = p, x1 = 1-p):
//
//   P0 (x1 H_i + x2 H_{i+1}) = c0 + c1 p + c2 p^2        (cell i, p in [0,1))
//   dpot-sum / fac           = d0 + d1 q                 (stencil centre j, q = p when j == i)
//
// and the field table stores (c0,c1,c2,d0,d1) per (cell, l, m) with the normalisation
// K_lm and the 1/scale factors folded in: the force costs 3 FMAs per term for the radial
// part.  The result differs from the reference only by floating-point re-association
// and by reciprocal-multiplies in place of divisions (~1e-13 relative; budget 1e-10/1e-8).
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

// Branch-free device forms of 1/sqrt(x), sqrt(x) and 1/x for the normal-range arguments of
// this path (radii, cylindrical radii, grid coordinates): the hardware seed (MUFU.RSQ64H /
// MUFU.RCP64H) and the same Newton polish the CUDA library routines use, without their
// denormal/inf/nan slow-path branch.  The branch costs a convergence barrier per call, which
// keeps the compiler from interleaving the set-up chains of the two particles a lane handles
// (the set-up is latency-bound).  Results are within 1 ulp of the IEEE operations (the host
// emulation uses those).
HD double fast_rsqrt(double x)
{
#if defined(__CUDA_ARCH__)
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x * y, y, 1.0);              // 1 - x y^2
  const double t = fma(e, 0.375, 0.5);               // y (1 + e/2 + 3 e^2/8)
  y = fma(t, y * e, y);
  const double e2 = fma(-x * y, y, 1.0);
  return fma(0.5 * y, e2, y);
#else
  return 1.0 / sqrt(x);
#endif
}
HD double fast_sqrt(double x)
{
#if defined(__CUDA_ARCH__)
  const double xs = x > 2.3e-308 ? x : 2.3e-308;     // sqrt(0) = 0 without a branch
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(xs));
  const double e = fma(-xs * y, y, 1.0);
  const double t = fma(e, 0.375, 0.5);
  y = fma(t, y * e, y);
  const double s = x * y;
  return fma(fma(-s, s, x), 0.5 * y, s);
#else
  return sqrt(x);
#endif
}
HD double fast_rcp(double x)
{
#if defined(__CUDA_ARCH__)
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, fma(e, e, e), y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
#else
  return 1.0 / x;
#endif
}

// Geometry + grid constants of one Sphere, passed to kernels by value.
struct SphGeom {
  int lmax, nmax, numr, cmap;
  int nrow;                 // (lmax+1)^2
  int nslot;                // (lmax+1)(lmax+2)/2
  double rmap, xmin, dxi;   // SLGridSph grid (exputil/SLGridMP2.cc:1355-1382)
  double inv_rmap, inv_dxi, inv_scale, inv_rmax;
  double scale, rmin, rmax; // SphericalBasis scale and expansion radii
  double cx, cy, cz, rtrunc2;
  // Force evaluated on ANOTHER component's particles (use_external): the expansion origin
  // (cx,cy,cz) is the owning component's centre, Component::freeze uses the particles' own
  // component's centre (fx,fy,fz) (src/SphericalBasis.cc:1519-1540).  ext = 0: both coincide.
  double fx, fy, fz;
  int ext;
  // covariance pass: 1 = the pcaeof variant (tvar, src/SphericalBasis.cc:536-545, 573-581), whose
  // m > 0 products carry no angular factor
  int cov_unit_m;
};

constexpr int kSphPoly = 5;   // double2 entries per (cell, slot): c0 c1 c2 d0 d1, each (cos, sin)

// slot of (l,m) in the m-major order used by the field table: m outer, l inner
HD constexpr int sph_slot(int lmax, int l, int m) { return m * (lmax + 1) - m * (m - 1) / 2 + (l - m); }
// row of (l,m) in the reference's coefficient layout (cos row; sin row is +1 for m>0)
HD constexpr int sph_row(int l, int m) { return l * l + (m ? 2 * m - 1 : 0); }

// a5: SLGridSph::r_to_xi / d_xi_to_r (exputil/SLGridMP2.cc:711-726, 748-765)
HD double sph_r_to_xi(const SphGeom &g, double r)
{
  if (g.cmap == 1) {
    double q = r * g.inv_rmap;
    return (q - 1.0) * fast_rcp(q + 1.0);
  }
  if (g.cmap == 2) return log(r);
  return r;
}
HD double sph_d_xi_to_r(const SphGeom &g, double xi)
{
  if (g.cmap == 1) return 0.5 * (1.0 - xi) * (1.0 - xi) * g.inv_rmap;
  if (g.cmap == 2) return exp(-xi);
  return 1.0;
}

// Position of a radius on the table: t = (xi - xmin)/dxi, so that the reference's
// indx = (int)t (exputil/SLGridMP2.cc:887, 969), x2 = t - indx, x1 = 1 - x2.
HD double sph_grid_coord(const SphGeom &g, double rs, double &xi)
{
  xi = sph_r_to_xi(g, rs);
  return (xi - g.xmin) * g.inv_dxi;
}
HD int sph_clamp_cell(const SphGeom &g, double t, int lo)
{
  int raw = (int)t;
  int c = raw < lo ? lo : raw;
  return c > g.numr - 2 ? g.numr - 2 : c;
}

// mode 0: coefficient accumulation (rmin <= r <= rmax, src/SphericalBasis.cc:489);
// mode 1: multistep_update (r < rmax only, src/SphericalBasis.cc:1177).
HD bool sph_in_range(const SphGeom &g, double r, int mode)
{
  return mode == 0 ? (r >= g.rmin && r <= g.rmax) : (r < g.rmax);
}

// Radial-cell key of a particle for the coefficient pass: its potential cell when it
// contributes (not frozen, in range: src/SphericalBasis.cc:489, src/Component.cc:4194-4202),
// else the dead cell numr-1.  The deposit takes the cell from this key.
HD uint32_t sph_cell_of(const SphGeom &g, const rec4 &pm, int mode = 0)
{
  double x = pm.x - g.cx, y = pm.y - g.cy, z = pm.z - g.cz;
  double r2 = x * x + y * y + z * z;
  if (r2 > g.rtrunc2) return (uint32_t)(g.numr - 1);
  double r = fast_sqrt(r2) + EXPB200_DSMALL;
  if (!sph_in_range(g, r, mode)) return (uint32_t)(g.numr - 1);
  double xi;
  double t = sph_grid_coord(g, r * g.inv_scale, xi);
  return (uint32_t)sph_clamp_cell(g, t, 0);
}

// Azimuth: cos(phi), sin(phi) of phi = atan2(y,x) without the transcendental
// (src/SphericalBasis.cc:493 then src/Basis.cc:104-105).  iR2 = 1/(x^2+y^2).
HD void sph_azimuth(double x, double y, double &c1, double &s1, double &R2, double &iR2)
{
  R2 = x * x + y * y;
  if (R2 > 0.0) {
    double Ri = fast_rsqrt(R2);
    c1 = x * Ri;
    s1 = y * Ri;
    iR2 = Ri * Ri;
  } else {
    c1 = 1.0;
    s1 = 0.0;
    iR2 = 0.0;
  }
}

// Visits the associated Legendre functions in m-outer / l-inner order, calling
// f(l, m, P_l^m, P_{l-1}^m) with the recursion of Basis::legendre_R (src/Basis.cc:14-52);
// on_m(m) is called once at the head of every m column and end_m(m) at its end.
// Only the columns M0 <= m <= M1 are expanded (on_m still runs for m < M0, so that a
// caller's azimuthal recursion advances).  inv[k] = 1/k.
template <int L, int M0, int M1, class M, class F, class E>
HD void sph_for_each_plm(double x, const double *inv, M &&on_m, F &&f, E &&end_m)
{
  double somx2 = fast_sqrt((1.0 - x) * (1.0 + x));
  double pmm = 1.0, fact = 1.0;
#pragma unroll
  for (int m = 0; m <= M1; m++) {
    if (m > 0) {
      pmm *= -fact * somx2;
      fact += 2.0;
    }
    on_m(m);
    if (m < M0) continue;
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
    end_m(m);
  }
}

// ---------------------------------------------------------------------------
// Coefficient pass, per particle (a8).  The particle's radial cell comes from its
// sort key.  Produces the two node weights w1 (node cell) and w2 (node cell+1), which
// carry mass * adb * (-4 pi) * P0(xi), and the angular arguments.
// ---------------------------------------------------------------------------
struct SphDep { double w1, w2, costh, c1, s1; };

HD void sph_deposit_setup(const SphGeom &g, const double *p0, const rec4 &pm, double adb,
                          int cell, SphDep &d)
{
  double x = pm.x - g.cx, y = pm.y - g.cy, z = pm.z - g.cz;
  double r2 = x * x + y * y + z * z;
  double r = fast_sqrt(r2) + EXPB200_DSMALL;
  d.costh = z * fast_rcp(r);
  double R2, iR2;
  sph_azimuth(x, y, d.c1, d.s1, R2, iR2);
  double xi;
  double t = sph_grid_coord(g, r * g.inv_scale, xi);
  double x2 = t - cell, x1 = 1.0 - x2;
  double P0 = x1 * EXPB200_LDG(p0 + cell) + x2 * EXPB200_LDG(p0 + cell + 1);
  double wgt = pm.w * adb * (-4.0 * M_PI) * P0;
  d.w1 = wgt * x1;
  d.w2 = wgt * x2;
}

// Visits yv[row] = P_lm {1,cos,sin}(m phi) (without the normalisation K_lm, which the
// contraction applies per row) for every coefficient row of the columns M0 <= m <= M1.
template <int L, int M0, int M1, class Store>
HD void sph_deposit_rows(const SphDep &d, const double *inv, Store &&store)
{
  double cm = 1.0, sm = 0.0, cm1 = 0.0, sm1 = 0.0;   // cos/sin(m phi), (m-1) phi
  sph_for_each_plm<L, M0, M1>(d.costh, inv, [&](int m) {
    if (m == 0) return;       // advance the Chebyshev recursion (src/Basis.cc:107-110)
    double cn, sn;
    if (m == 1) { cn = d.c1; sn = d.s1; }
    else { cn = 2.0 * d.c1 * cm - cm1; sn = 2.0 * d.c1 * sm - sm1; }
    cm1 = cm; sm1 = sm; cm = cn; sm = sn;
  }, [&](int l, int m, double plm, double) {
    if (m == 0) store(l, 0, 0, plm);
    else { store(l, m, 0, plm * cm); store(l, m, 1, plm * sm); }
  }, [](int) {});
}

// Compact index of (l, m, cos/sin) among the rows of the columns M0..M1 (the register
// accumulators of one deposit sweep), and their number.
template <int L, int M0>
HD constexpr int sph_compact(int l, int m, int cs)
{
  return M0 == 0 ? (m == 0 ? l : (L + 1) + 2 * (sph_slot(L, l, m) - (L + 1)) + cs)
                 : 2 * (sph_slot(L, l, m) - sph_slot(L, M0, M0)) + cs;
}
template <int L, int M0, int M1>
HD constexpr int sph_compact_count()
{
  return sph_compact<L, M0>(L, M1, M1 > 0 ? 1 : 0) + 1;
}

// ---------------------------------------------------------------------------
// Field table entry of (row i, one slot) from the node values of H at i-1, i, i+1
// (K_lm already folded in): out[0..2] = c0,c1,c2 of cell i, out[3..4] = d0,d1 of the
// stencil centred on i; cos in .x, sin in .y.  1/scale (potential) and 1/scale^2
// (radial derivative) are folded in (src/SphericalBasis.cc:1640-1643).
// ---------------------------------------------------------------------------
HD void sph_poly_entry(const SphGeom &g, const double *p0, int i, double2 Hm, double2 Hc, double2 Hp,
                       double2 *out)
{
  const double is = g.inv_scale, is2 = g.inv_scale * g.inv_scale;
  double2 z2; z2.x = 0.0; z2.y = 0.0;
  for (int k = 0; k < kSphPoly; k++) out[k] = z2;
  if (i <= g.numr - 2) {
    double p0c = p0[i], dp0 = p0[i + 1] - p0[i];
    double dHx = Hp.x - Hc.x, dHy = Hp.y - Hc.y;
    out[0].x = p0c * Hc.x * is;                 out[0].y = p0c * Hc.y * is;
    out[1].x = (p0c * dHx + dp0 * Hc.x) * is;   out[1].y = (p0c * dHy + dp0 * Hc.y) * is;
    out[2].x = dp0 * dHx * is;                  out[2].y = dp0 * dHy * is;
  }
  if (i >= 1 && i <= g.numr - 2) {
    double wm = p0[i - 1], wc = p0[i], wp = p0[i + 1];
    double Wmx = wm * Hm.x, Wcx = wc * Hc.x, Wpx = wp * Hp.x;
    double Wmy = wm * Hm.y, Wcy = wc * Hc.y, Wpy = wp * Hp.y;
    out[3].x = 0.5 * (Wpx - Wmx) * is2;         out[3].y = 0.5 * (Wpy - Wmy) * is2;
    out[4].x = (Wmx - 2.0 * Wcx + Wpx) * is2;   out[4].y = (Wmy - 2.0 * Wcy + Wpy) * is2;
  }
}

// ---------------------------------------------------------------------------
// Force pass, per particle (a11).  T is the field table [numr][nslot][kSphPoly] of
// (cos,sin) pairs (see sph_poly_entry).  sph_force_setup does the per-particle geometry
// (returns false for frozen particles) and tells which table rows the particle reads
// (S.cell for c0..c2, S.j for d0,d1); sph_force_eval walks the (l,m) terms given pointers
// to those two rows (global memory, or a shared-memory copy) and adds into
// (ax,ay,az,pot).  OFF = true compiles the r > rmax multipole continuation
// (src/SphericalBasis.cc:1555-1560, 1605-1628) in; with OFF = false the caller
// guarantees r <= rmax.
// ---------------------------------------------------------------------------
struct SphForce {
  double x, y, z, ir, costh, c1, s1, R2, iR2;
  double ratio, r0i;          // r > rmax continuation
  double p, q, fac, xd, somx2i;
  int cell, j;
  bool ioff;
};

// Legendre derivative clamp (src/Basis.cc:81-86)
HD double sph_clamp_x(double x)
{
  if (1.0 - fabs(x) < EXPB200_MINEPS) return x > 0 ? 1.0 - EXPB200_MINEPS : -(1.0 - EXPB200_MINEPS);
  return x;
}

template <bool OFF>
HD bool sph_force_setup(const SphGeom &g, const rec4 &pm, SphForce &S)
{
  S.x = pm.x - g.cx; S.y = pm.y - g.cy; S.z = pm.z - g.cz;
  double r2 = S.x * S.x + S.y * S.y + S.z * S.z;
  if (g.ext) {
    const double a = pm.x - g.fx, b = pm.y - g.fy, c = pm.z - g.fz;
    if (a * a + b * b + c * c > g.rtrunc2) return false;
  } else if (r2 > g.rtrunc2) return false;

  double r = fast_sqrt(r2) + EXPB200_DSMALL;
  S.ir = fast_rcp(r);
  S.costh = S.z * S.ir;
  sph_azimuth(S.x, S.y, S.c1, S.s1, S.R2, S.iR2);

  S.ioff = OFF && r > g.rmax;
  S.ratio = 1.0; S.r0i = 0.0;
  if (OFF && S.ioff) {               // evaluate at rmax, continue outwards
    S.ratio = g.rmax * S.ir;
    S.r0i = S.ir * g.inv_scale;
    r = g.rmax;
    S.ir = g.inv_rmax;
  }

  double xi;
  const double t = sph_grid_coord(g, r * g.inv_scale, xi);
  S.cell = sph_clamp_cell(g, t, 0);
  S.j = sph_clamp_cell(g, t, 1);
  S.p = t - S.cell;
  S.q = t - S.j;
  S.fac = sph_d_xi_to_r(g, xi) * g.inv_dxi;

  S.xd = sph_clamp_x(S.costh);
  S.somx2i = fast_rcp(S.xd * S.xd - 1.0);
  return true;
}

// Cartesian projection of the four field sums of a particle (src/SphericalBasis.cc:1640-1653;
// note: r is the clamped radius when r0 > rmax): sums = {potl, potr, pott, potp} as the (l,m)
// loops leave them; o[0..3] += (ax, ay, az, pot).
HD void sph_force_project(const SphForce &S, bool ioff, const double *sums, double *o)
{
  const double tt = sums[2] * S.somx2i, rr = ioff ? sums[1] : sums[1] * S.fac;
  const double ir = S.ir, ir3 = ir * ir * ir;
  o[0] += -(rr * S.x * ir - tt * S.x * S.z * ir3);
  o[1] += -(rr * S.y * ir - tt * S.y * S.z * ir3);
  o[2] += -(rr * S.z * ir + tt * S.R2 * ir3);
  if (S.R2 > EXPB200_DSMALL) {
    o[0] += sums[3] * S.y * S.iR2;
    o[1] += -sums[3] * S.x * S.iR2;
  }
  o[3] += sums[0];
}

// The (l,m) loops of one particle: sums = {potl, potr, pott, potp} (potr without the
// d xi/d r factor, pott without 1/(x^2-1)).  Reads only costh, xd, c1, s1, p, q (and ratio,
// r0i, ioff with OFF) of S.
template <int L, bool OFF, class LD>
HD void sph_force_sums(const double *inv, const SphForce &S, const double2 *Tc, const double2 *Td,
                       LD &&ld, double *sums)
{
  const double p = S.p, q = S.q, xd = S.xd;
  const bool ioff = OFF && S.ioff;
  double potl = 0.0, potr = 0.0, pott = 0.0, potp = 0.0;
  double cm = 1.0, sm = 0.0, cm1 = 0.0, sm1 = 0.0;
  double Cc, Cs, Dc, Ds, Uc, Us;     // column sums over l of P pc, P ps, P dpc, P dps, P' pc, P' ps
  double pw_m = 1.0, pw = 1.0;       // ratio^m at the column head; ratio^(l+1) per term

  sph_for_each_plm<L, 0, L>(S.costh, inv, [&](int m) {
    if (m > 0) {
      double cn, sn;
      if (m == 1) { cn = S.c1; sn = S.s1; }
      else { cn = 2.0 * S.c1 * cm - cm1; sn = 2.0 * S.c1 * sm - sm1; }
      cm1 = cm; sm1 = sm; cm = cn; sm = sn;
      if (OFF) pw_m *= S.ratio;
    }
    if (OFF) pw = pw_m;
    Cc = Cs = Dc = Ds = Uc = Us = 0.0;
  }, [&](int l, int m, double plm, double plm1) {
    const int o = sph_slot(L, l, m) * kSphPoly;
    double2 c0 = ld(Tc + o), c1p = ld(Tc + o + 1), c2 = ld(Tc + o + 2);
    double2 d0 = ld(Td + o + 3), d1 = ld(Td + o + 4);
    double dpl = (xd * l) * plm - (l + m) * plm1;      // (x^2-1) dP/dx
    double pc = (c2.x * p + c1p.x) * p + c0.x;
    double dpc = d1.x * q + d0.x;
    if (OFF) {
      pw *= S.ratio;                // ratio^(l+1)
      if (ioff) {
        pc *= pw;
        dpc = -pc * S.r0i * (l + 1);
      }
    }
    Cc += plm * pc;
    Dc += plm * dpc;
    Uc += dpl * pc;
    if (m > 0) {
      double ps = (c2.y * p + c1p.y) * p + c0.y;
      double dps = d1.y * q + d0.y;
      if (OFF && ioff) {
        ps *= pw;
        dps = -ps * S.r0i * (l + 1);
      }
      Cs += plm * ps;
      Ds += plm * dps;
      Us += dpl * ps;
    }
  }, [&](int m) {
    if (m == 0) {
      potl += Cc;
      potr += Dc;
      pott += Uc;
    } else {
      potl += Cc * cm + Cs * sm;
      potr += Dc * cm + Ds * sm;
      pott += Uc * cm + Us * sm;
      potp += (Cs * cm - Cc * sm) * m;
    }
  });

  sums[0] = potl; sums[1] = potr; sums[2] = pott; sums[3] = potp;
}

template <int L, bool OFF, class LD>
HD void sph_force_eval(const double *inv, const SphForce &S, const double2 *Tc, const double2 *Td,
                       LD &&ld, double &ax, double &ay, double &az, double &pot)
{
  double sums[4], o[4] = {ax, ay, az, pot};
  sph_force_sums<L, OFF>(inv, S, Tc, Td, ld, sums);
  sph_force_project(S, OFF && S.ioff, sums, o);
  ax = o[0]; ay = o[1]; az = o[2]; pot = o[3];
}

// Two particles that read the same two table rows (same cell and stencil centre) evaluated
// together, so that every table entry is delivered to the lane once for both: the register
// file's input bandwidth, not the FP64 pipe, bounds the single-particle form (a table entry
// is 16 B per lane per term).  r <= rmax for both (no continuation).
template <int L, class LD>
HD void sph_force_sums2(const double *inv, const SphForce &A, const SphForce &B, const double2 *Tc,
                        const double2 *Td, LD &&ld, double *sa, double *sb)
{
  double potl[2] = {0.0, 0.0}, potr[2] = {0.0, 0.0}, pott[2] = {0.0, 0.0}, potp[2] = {0.0, 0.0};
  double cm[2] = {1.0, 1.0}, sm[2] = {0.0, 0.0}, cm1[2] = {0.0, 0.0}, sm1[2] = {0.0, 0.0};
  const double x[2] = {A.costh, B.costh}, xd[2] = {A.xd, B.xd};
  const double c1[2] = {A.c1, B.c1}, s1[2] = {A.s1, B.s1};
  const double p[2] = {A.p, B.p}, q[2] = {A.q, B.q};
  double somx2[2], pmm[2] = {1.0, 1.0};
#pragma unroll
  for (int k = 0; k < 2; k++) somx2[k] = fast_sqrt((1.0 - x[k]) * (1.0 + x[k]));
  double fact = 1.0;

#pragma unroll
  for (int m = 0; m <= L; m++) {
    if (m > 0) {
#pragma unroll
      for (int k = 0; k < 2; k++) {
        pmm[k] *= -fact * somx2[k];
        double cn, sn;
        if (m == 1) { cn = c1[k]; sn = s1[k]; }
        else { cn = 2.0 * c1[k] * cm[k] - cm1[k]; sn = 2.0 * c1[k] * sm[k] - sm1[k]; }
        cm1[k] = cm[k]; sm1[k] = sm[k]; cm[k] = cn; sm[k] = sn;
      }
      fact += 2.0;
    }
    double Cc[2] = {0.0, 0.0}, Cs[2] = {0.0, 0.0}, Dc[2] = {0.0, 0.0}, Ds[2] = {0.0, 0.0},
           Uc[2] = {0.0, 0.0}, Us[2] = {0.0, 0.0};
    double pl1[2], pl2[2] = {0.0, 0.0};
#pragma unroll
    for (int l = m; l <= L; l++) {
      double plm[2], plm1[2];
#pragma unroll
      for (int k = 0; k < 2; k++) {
        if (l == m) { plm[k] = pmm[k]; plm1[k] = 0.0; }
        else if (l == m + 1) { plm[k] = x[k] * (2 * m + 1) * pmm[k]; plm1[k] = pmm[k]; }
        else { plm[k] = (x[k] * (2 * l - 1) * pl1[k] - (l + m - 1) * pl2[k]) * inv[l - m]; plm1[k] = pl1[k]; }
      }
      const int o = sph_slot(L, l, m) * kSphPoly;
      const double2 c0 = ld(Tc + o), c1p = ld(Tc + o + 1), c2 = ld(Tc + o + 2);
      const double2 d0 = ld(Td + o + 3), d1 = ld(Td + o + 4);
#pragma unroll
      for (int k = 0; k < 2; k++) {
        const double dpl = (xd[k] * l) * plm[k] - (l + m) * plm1[k];
        const double pc = (c2.x * p[k] + c1p.x) * p[k] + c0.x;
        const double dpc = d1.x * q[k] + d0.x;
        Cc[k] += plm[k] * pc;
        Dc[k] += plm[k] * dpc;
        Uc[k] += dpl * pc;
        if (m > 0) {
          const double ps = (c2.y * p[k] + c1p.y) * p[k] + c0.y;
          const double dps = d1.y * q[k] + d0.y;
          Cs[k] += plm[k] * ps;
          Ds[k] += plm[k] * dps;
          Us[k] += dpl * ps;
        }
        pl2[k] = plm1[k];
        pl1[k] = plm[k];
      }
    }
#pragma unroll
    for (int k = 0; k < 2; k++) {
      if (m == 0) {
        potl[k] += Cc[k];
        potr[k] += Dc[k];
        pott[k] += Uc[k];
      } else {
        potl[k] += Cc[k] * cm[k] + Cs[k] * sm[k];
        potr[k] += Dc[k] * cm[k] + Ds[k] * sm[k];
        pott[k] += Uc[k] * cm[k] + Us[k] * sm[k];
        potp[k] += (Cs[k] * cm[k] - Cc[k] * sm[k]) * m;
      }
    }
  }

  sa[0] = potl[0]; sa[1] = potr[0]; sa[2] = pott[0]; sa[3] = potp[0];
  sb[0] = potl[1]; sb[1] = potr[1]; sb[2] = pott[1]; sb[3] = potp[1];
}

template <int L, class LD>
HD void sph_force_eval2(const double *inv, const SphForce &A, const SphForce &B, const double2 *Tc,
                        const double2 *Td, LD &&ld, double *oa, double *ob)
{
  double sa[4], sb[4];
  sph_force_sums2<L>(inv, A, B, Tc, Td, ld, sa, sb);
  sph_force_project(A, false, sa, oa);
  sph_force_project(B, false, sb, ob);
}

template <int L, bool OFF>
HD bool sph_force_particle(const SphGeom &g, const double *inv, const double2 *T, const rec4 &pm,
                           double &ax, double &ay, double &az, double &pot)
{
  SphForce S;
  if (!sph_force_setup<OFF>(g, pm, S)) return false;
  sph_force_eval<L, OFF>(inv, S, T + (size_t)S.cell * g.nslot * kSphPoly,
                         T + (size_t)S.j * g.nslot * kSphPoly,
                         [](const double2 *a) { return EXPB200_LDG(a); }, ax, ay, az, pot);
  return true;
}

} // namespace expb200
