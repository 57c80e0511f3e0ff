// cyl_emul.cu -- TEST INFRASTRUCTURE: the cylinder's __host__ __device__ per-particle
// arithmetic (exp_b200/csrc/cyl_math.cuh) run in plain host loops with the data flow of
// cylinder.cu (node sums, table contraction, field table, 4-node gather).  Never shipped.
#include "../../exp_b200/csrc/cyl_math.cuh"

#include <vector>

using namespace expb200;

static CylGeom make_geom(const expb200_cylinder_config *cfg, const double *center, double rtrunc)
{
  CylGeom g;
  g.mmax = cfg->mmax; g.norder = cfg->norder; g.numx = cfg->numx; g.numy = cfg->numy;
  g.cmapr = cfg->cmapr; g.cmapz = cfg->cmapz;
  g.nk = 2 * cfg->mmax + 1; g.ncell = cfg->numx * cfg->numy; g.nnode = (cfg->numx + 1) * (cfg->numy + 1);
  g.ascale = cfg->ascale; g.hscale = cfg->hscale;
  g.rtable = M_SQRT1_2 * cfg->rmax;
  g.cx = center[0]; g.cy = center[1]; g.cz = center[2]; g.rtrunc2 = rtrunc * rtrunc; g.rotate = 0;
  g.xmin = cyl_r_to_xi(g, cfg->rmin * cfg->ascale);
  g.dx = (cyl_r_to_xi(g, g.rtable * cfg->ascale) - g.xmin) / cfg->numx;
  g.ymin = cyl_z_to_y(g, -g.rtable * cfg->ascale);
  g.dy = (cyl_z_to_y(g, g.rtable * cfg->ascale) - g.ymin) / cfg->numy;
  g.rmax2 = cfg->rcylmax * cfg->rcylmax * cfg->acyl * cfg->acyl;
  g.r2grid = (cfg->ascale * g.rtable) * (cfg->ascale * g.rtable);
  for (int k = 0; k < 9; k++) g.body[k] = g.orig[k] = (k % 4 == 0) ? 1.0 : 0.0;
  return g;
}

extern "C" {

int emul_cyl_coefficients(const expb200_cylinder_config *cfg, const double *potC, const double *potS,
                          long n, const double *pos, const double *mass, const double *center,
                          double rtrunc, double adb, const double *body, double *cosN, double *sinN,
                          long *used, double *cylmass)
{
  CylGeom g = make_geom(cfg, center, rtrunc);
  if (body) { for (int k = 0; k < 9; k++) g.body[k] = body[k]; g.rotate = 1; }
  std::vector<double> G((size_t)g.nk * g.nnode, 0.0), trig(g.nk);
  long nu = 0; double ms = 0.0;
  for (long i = 0; i < n; i++) {
    rec4 pm{pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], mass[i]};
    uint32_t key = cyl_cell_key(g, pm);
    if (key <= (uint32_t)g.ncell) { nu++; ms += pm.w * adb; }
    CylCell c;
    if (!cyl_deposit_particle(g, pm, adb, c, [&](int k, double v) { trig[k] = v; })) continue;
    if ((uint32_t)(c.ix * g.numy + c.iy) != key) return -2;
    const int n00 = c.ix * (g.numy + 1) + c.iy;
    for (int k = 0; k < g.nk; k++) {
      double *Gk = G.data() + (size_t)k * g.nnode;
      Gk[n00] += trig[k] * c.c00;
      Gk[n00 + g.numy + 1] += trig[k] * c.c10;
      Gk[n00 + 1] += trig[k] * c.c01;
      Gk[n00 + g.numy + 2] += trig[k] * c.c11;
    }
  }
  for (int m = 0; m <= g.mmax; m++)
    for (int nn = 0; nn < g.norder; nn++) {
      double sc = 0.0, ss = 0.0;
      const double *tc = potC + ((size_t)m * g.norder + nn) * g.nnode;
      const double *ts = potS + ((size_t)m * g.norder + nn) * g.nnode;
      for (int i = 0; i < g.nnode; i++) {
        sc += tc[i] * G[(size_t)m * g.nnode + i];
        if (m) ss += ts[i] * G[(size_t)(g.mmax + m) * g.nnode + i];
      }
      // rows the reference's get_pot never fills (m > MLIM, odd m with EVEN_M) stay zero
      const bool off = m > cfg->MLIM || (cfg->EVEN_M && (m & 1));
      cosN[m * g.norder + nn] = off ? 0.0 : sc;
      sinN[m * g.norder + nn] = off ? 0.0 : ss;
    }
  *used = nu; *cylmass = ms;
  return 0;
}

int emul_cyl_accel(const expb200_cylinder_config *cfg, const double *const *tabs /* [3][2] */,
                   const double *cosN, const double *sinN, double cylmass, long n, const double *pos,
                   const double *mass, const double *center, double rtrunc, const double *body,
                   const double *orig, double *acc, double *pot)
{
  CylGeom g = make_geom(cfg, center, rtrunc);
  if (body) { for (int k = 0; k < 9; k++) { g.body[k] = body[k]; g.orig[k] = orig[k]; } g.rotate = 1; }
  const int nkpad = (g.nk + 1) & ~1;
  std::vector<double> F((size_t)g.nnode * 3 * nkpad, 0.0);
  const int mmin = cfg->MMIN > 0 ? cfg->MMIN : 0, mlim = cfg->MLIM < g.mmax ? cfg->MLIM : g.mmax;
  const int nmin = cfg->NMIN > 0 ? cfg->NMIN : 0, nlim = cfg->NLIM < g.norder ? cfg->NLIM : g.norder;
  for (int node = 0; node < g.nnode; node++)
    for (int kind = 0; kind < 3; kind++)
      for (int k = 0; k < g.nk; k++) {
        int cs = k > g.mmax, m = cs ? k - g.mmax : k;
        double v = 0.0;
        if (m >= mmin && m <= mlim && !(cfg->EVEN_M && (m & 1))) {
          const double *tab = tabs[kind * 2 + cs] + (size_t)m * g.norder * g.nnode + node;
          const double *c = (cs ? sinN : cosN) + (size_t)m * g.norder;
          for (int nn = nmin; nn < nlim; nn++) v += tab[(size_t)nn * g.nnode] * c[nn];
        }
        F[((size_t)node * 3 + kind) * nkpad + k] = v;
      }
  for (long i = 0; i < n; i++) {
    rec4 pm{pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], mass[i]};
    double ax = 0, ay = 0, az = 0, pp = 0;
    cyl_force_particle(g, F.data(), nkpad, cylmass, pm, ax, ay, az, pp);
    acc[3 * i] += ax; acc[3 * i + 1] += ay; acc[3 * i + 2] += az; pot[i] += pp;
  }
  return 0;
}
}
