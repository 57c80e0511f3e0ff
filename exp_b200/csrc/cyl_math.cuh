// cyl_math.cuh -- per-particle arithmetic of the cylinder path (Cylinder + EmpCylSL
// tables), shared by the CUDA kernels (cylinder.cu) and the host emulation in tests.
//
// Same factorisation as the sphere (sph_math.cuh), on the 2-D (x,y) table grid: the
// bilinear interpolant of EmpCylSL::get_pot / accumulated_eval
// (exputil/EmpCylSL.cc:5557-5632, 5256-5410) is linear in the four corner nodes, so
//   cosN[m][n] = sum_node potC[m][n][node] * G[cos m][node],
//   G[k][node] = sum_p (-4 pi mass_p) trig_k(phi_p) * (corner weight of p on node)
// and the force needs only  F[node][kind][k] = sum_n table_kind[m][n][node] * coef[m][n].
// A particle then deposits / reads 13 azimuthal terms on 4 nodes instead of 234 (702)
// table values on 4 nodes.
#pragma once

#include "common.cuh"

#if defined(__CUDACC__)
#define HDC __host__ __device__ __forceinline__
#else
#define HDC inline
#endif

#if defined(__CUDA_ARCH__)
#define EXPB200_LDG_C(p) __ldg(p)
#else
#define EXPB200_LDG_C(p) (*(p))
#endif

namespace expb200 {

#if !defined(__CUDACC__)
using std::asinh; using std::erf; using std::fabs; using std::sqrt;
#endif

constexpr int kCylMaxM = 16;

struct CylGeom {
  int mmax, norder, numx, numy, cmapr, cmapz;
  int nk;                    // 2*mmax+1 azimuthal terms: cos 0..mmax, sin 1..mmax
  int ncell;                 // numx*numy
  int nnode;                 // (numx+1)*(numy+1)
  double ascale, hscale, rtable, xmin, dx, ymin, dy;
  double rmax2;              // (rcylmax*acyl)^2, src/Cylinder.cc:752
  double r2grid;             // (ascale*rtable)^2, src/Cylinder.cc:1281-1282
  double cx, cy, cz, rtrunc2;
  int rotate;
  double body[9], orig[9];   // orient->transformBody()/transformOrig(), row-major
  // covariance pass: 0 = covar (complex vec vec^H), 1 = PCAVAR (moduli with the azimuthal
  // factors), 2 = PCAEOF (moduli of the table values), exputil/EmpCylSL.cc:4089-4141
  int cov_mode;
};

// EmpCylSL::r_to_xi (exputil/EmpCylSL.cc:6446-6463)
HDC double cyl_r_to_xi(const CylGeom &g, double r)
{
  if (g.cmapr > 0) return (r / g.ascale - 1.0) / (r / g.ascale + 1.0);
  return r;
}
// EmpCylSL::z_to_y (exputil/EmpCylSL.cc:7109-7117)
HDC double cyl_z_to_y(const CylGeom &g, double z)
{
  if (g.cmapz == 1) return z / (fabs(z) + 2.2250738585072014e-308) * asinh(fabs(z / g.hscale));
  if (g.cmapz == 2) return z / sqrt(z * z + g.hscale * g.hscale);
  return z;
}

struct CylCell { int ix, iy; double c00, c10, c01, c11; };

// index / weights of get_pot and accumulated_eval (enforce_limits == false)
HDC CylCell cyl_locate(const CylGeom &g, double r, double z)
{
  double X = (cyl_r_to_xi(g, r) - g.xmin) / g.dx;
  double Y = (cyl_z_to_y(g, z) - g.ymin) / g.dy;
  int ix = (int)X, iy = (int)Y;
  if (ix < 0) ix = 0;
  if (iy < 0) iy = 0;
  if (ix >= g.numx) ix = g.numx - 1;
  if (iy >= g.numy) iy = g.numy - 1;
  double delx0 = (double)ix + 1.0 - X, dely0 = (double)iy + 1.0 - Y;
  double delx1 = X - (double)ix, dely1 = Y - (double)iy;
  return CylCell{ix, iy, delx0 * dely0, delx1 * dely0, delx0 * dely1, delx1 * dely1};
}

HDC void cyl_rot(const double *R, double &x, double &y, double &z)
{
  double a = R[0] * x + R[1] * y + R[2] * z;
  double b = R[3] * x + R[4] * y + R[5] * z;
  double c = R[6] * x + R[7] * y + R[8] * z;
  x = a; y = b; z = c;
}

// Classification of a particle for the coefficient pass
// (src/Cylinder.cc:836-858, exputil/EmpCylSL.cc:4062-4063):
//   returns 0 = unused, 1 = used but off the table (counts + mass only), 2 = deposits.
// mode 1 = multistep_update (src/Cylinder.cc:1751-1773): no orientation, no rcylmax gate.
HDC int cyl_classify(const CylGeom &g, const rec4 &pm, double &x, double &y, double &z,
                     double &r, int mode = 0)
{
  x = pm.x - g.cx; y = pm.y - g.cy; z = pm.z - g.cz;
  if (x * x + y * y + z * z > g.rtrunc2) return 0;          // Component::freeze
  if (g.rotate && mode == 0) cyl_rot(g.body, x, y, z);
  double r2 = x * x + y * y;
  r = sqrt(r2);
  double R2 = r2 + z * z;
  if (mode == 0 && !(R2 < g.rmax2)) return 0;
  double rr = sqrt(r * r + z * z);
  if (rr / g.ascale > g.rtable) return 1;
  return 2;
}

// key within a level: cell index, ncell (used, off table) or ncell+1 (unused)
HDC uint32_t cyl_cell_key(const CylGeom &g, const rec4 &pm, int mode = 0)
{
  double x, y, z, r;
  int cls = cyl_classify(g, pm, x, y, z, r, mode);
  if (cls == 0) return (uint32_t)g.ncell + 1u;
  if (cls == 1) return (uint32_t)g.ncell;
  double zz = z;
  if (zz / g.ascale > g.rtable) zz = g.rtable * g.ascale;
  if (zz / g.ascale < -g.rtable) zz = -g.rtable * g.ascale;
  CylCell c = cyl_locate(g, r, zz);
  return (uint32_t)(c.ix * g.numy + c.iy);
}

// cos(m phi), sin(m phi) for m = 0..mmax from (x, y); phi = atan2(y, x)
template <class F>
HDC void cyl_trig(int mmax, double x, double y, double r, F &&f)
{
  double c1 = 1.0, s1 = 0.0;
  if (r > 0.0) { c1 = x / r; s1 = y / r; }
  double cm = 1.0, sm = 0.0, cm1 = 0.0, sm1 = 0.0;
  f(0, 1.0, 0.0);
  for (int m = 1; m <= mmax; m++) {
    double cn, sn;
    if (m == 1) { cn = c1; sn = s1; }
    else { cn = 2.0 * c1 * cm - cm1; sn = 2.0 * c1 * sm - sm1; }
    cm1 = cm; sm1 = sm; cm = cn; sm = sn;
    f(m, cm, sm);
  }
}

// Coefficient pass per particle: trig[k] (k = m for cos, mmax+m for sin) scaled by
// -4 pi * mass * adb, and the cell with its four corner weights.
template <class Store>
HDC bool cyl_deposit_particle(const CylGeom &g, const rec4 &pm, double adb, CylCell &cell,
                              Store &&store, int mode = 0)
{
  double x, y, z, r;
  if (cyl_classify(g, pm, x, y, z, r, mode) != 2) return false;
  double zz = z;
  if (zz / g.ascale > g.rtable) zz = g.rtable * g.ascale;       // EmpCylSL::get_pot clamp
  if (zz / g.ascale < -g.rtable) zz = -g.rtable * g.ascale;
  cell = cyl_locate(g, r, zz);
  const double w = -4.0 * M_PI * (pm.w * adb);
  cyl_trig(g.mmax, x, y, r, [&](int m, double c, double s) {
    store(m, w * c);
    if (m) store(g.mmax + m, w * s);
  });
  return true;
}

// Force pass per particle (src/Cylinder.cc:1354-1436 + exputil/EmpCylSL.cc:5256-5410).
// F is the field table [node][3][nkpad] (kind 0 pot, 1 rforce, 2 zforce).
HDC void cyl_force_particle(const CylGeom &g, const double *F, int nkpad, double cylmass,
                            const rec4 &pm, double &ax, double &ay, double &az, double &pot)
{
  const double ratmin = 0.75, maxerf = 3.0;
  const double midpt = ratmin + 0.5 * (1.0 - ratmin);
  const double rsmth = 0.5 * (1.0 - ratmin) / maxerf;

  double x = pm.x - g.cx, y = pm.y - g.cy, z = pm.z - g.cz;
  if (g.rotate) cyl_rot(g.body, x, y, z);
  double r2 = x * x + y * y;
  double r = sqrt(r2) + EXPB200_DSMALL;
  double ratio = sqrt((r2 + z * z) / g.r2grid);
  double frac, cfrac;
  if (ratio >= 1.0) { frac = 0.0; cfrac = 1.0; }
  else if (ratio > ratmin) { frac = 0.5 * (1.0 - erf((ratio - midpt) / rsmth)); cfrac = 1.0 - frac; }
  else { cfrac = 0.0; frac = 1.0; }

  double f0 = 0.0, f1 = 0.0, f2 = 0.0, pa = 0.0;
  if (ratio < 1.0) {
    double p = 0.0, fr = 0.0, fz = 0.0, fp = 0.0;
    double rr = sqrt(r * r + z * z);
    if (!(rr / g.ascale > g.rtable)) {
      CylCell c = cyl_locate(g, r, z);
      const int stride = 3 * nkpad;
      const double *n00 = F + (size_t)(c.ix * (g.numy + 1) + c.iy) * stride;
      const double *n01 = n00 + stride;                       // (ix, iy+1)
      const double *n10 = n00 + (size_t)(g.numy + 1) * stride; // (ix+1, iy)
      const double *n11 = n10 + stride;
      // phi = atan2(y, x) of the (unshifted) planar position
      double rp = sqrt(r2);
      cyl_trig(g.mmax, x, y, rp, [&](int m, double cm, double sm) {
        auto bil = [&](int kind, int k) {
          int o = kind * nkpad + k;
          return EXPB200_LDG_C(n00 + o) * c.c00 + EXPB200_LDG_C(n10 + o) * c.c10 +
                 EXPB200_LDG_C(n01 + o) * c.c01 + EXPB200_LDG_C(n11 + o) * c.c11;
        };
        double pc = bil(0, m), rc = bil(1, m), zc = bil(2, m);
        p += cm * pc;
        fr += cm * rc;
        fz += cm * zc;
        fp += sm * pc * m;
        if (m) {
          int k = g.mmax + m;
          double ps = bil(0, k), rs = bil(1, k), zs = bil(2, k);
          p += sm * ps;
          fr += sm * rs;
          fz += sm * zs;
          fp += -cm * ps * m;
        }
      });
    }
    f0 = (fr * x / r - fp * y / r2) * frac;
    f1 = (fr * y / r + fp * x / r2) * frac;
    f2 = fz * frac;
    pa = p * frac;
  }
  if (ratio > ratmin) {
    double r3 = r2 + z * z;
    double p = -cylmass / sqrt(r3);
    double fr = p / r3;
    f0 += x * fr * cfrac;
    f1 += y * fr * cfrac;
    f2 += z * fr * cfrac;
    pa += p * cfrac;
  }
  if (g.rotate) cyl_rot(g.orig, f0, f1, f2);
  ax += f0; ay += f1; az += f2; pot += pa;
}

} // namespace expb200
