// sph_host.h -- host-side setup shared by sphere.cu and the CPU emulation used in
// tests (tests/emul/sph_emul.cu): grid constants, normalisation constants and the
// (l,m) -> coefficient-row maps.
#pragma once
#include "sph_math.cuh"

#include <cmath>
#include <vector>

namespace expb200 {

inline double host_factrl(int n) { return n <= 1 ? 1.0 : n * host_factrl(n - 1); }  // exputil/biorth_grid.cc:183-189

inline SphGeom make_sph_geom(const expb200_sphere_config &cfg)
{
  SphGeom g;
  const int L = cfg.lmax, R = cfg.numr;
  g.lmax = L; g.nmax = cfg.nmax; g.numr = R; g.cmap = cfg.cmap;
  g.nrow = (L + 1) * (L + 1);
  g.nslot = (L + 1) * (L + 2) / 2;
  g.rmap = cfg.rmap;
  // SLGridSph::init_table (exputil/SLGridMP2.cc:1362-1375)
  if (cfg.cmap == 1) {
    g.xmin = (cfg.rmin_grid / cfg.rmap - 1.0) / (cfg.rmin_grid / cfg.rmap + 1.0);
    double xmax = (cfg.rmax_grid / cfg.rmap - 1.0) / (cfg.rmax_grid / cfg.rmap + 1.0);
    g.dxi = (xmax - g.xmin) / (R - 1);
  } else if (cfg.cmap == 2) {
    g.xmin = std::log(cfg.rmin_grid);
    g.dxi = (std::log(cfg.rmax_grid) - g.xmin) / (R - 1);
  } else {
    g.xmin = cfg.rmin_grid;
    g.dxi = (cfg.rmax_grid - cfg.rmin_grid) / (R - 1);
  }
  g.scale = cfg.scale; g.rmin = cfg.rmin; g.rmax = cfg.rmax;
  g.inv_rmap = 1.0 / g.rmap; g.inv_dxi = 1.0 / g.dxi; g.inv_scale = 1.0 / g.scale; g.inv_rmax = 1.0 / g.rmax;
  g.cx = g.cy = g.cz = 0.0; g.rtrunc2 = 1.0e300;
  g.fx = g.fy = g.fz = 0.0; g.ext = 0; g.cov_unit_m = 0;
  return g;
}

// factorial(l,m) of src/SphericalBasis.cc:330-335 and the reciprocal table
inline void make_sph_constants(std::vector<double> &K, std::vector<double> &inv)
{
  K.assign((kMaxL + 1) * (kMaxL + 1), 0.0);
  inv.assign(2 * kMaxL + 4, 0.0);
  for (int l = 0; l <= kMaxL; l++)
    for (int m = 0; m <= l; m++) {
      double f = std::sqrt((2.0 * l + 1.0) / (4.0 * M_PI) * host_factrl(l - m) / host_factrl(l + m));
      if (m) f *= M_SQRT2;
      K[l * (kMaxL + 1) + m] = f;
    }
  for (int k = 1; k < 2 * kMaxL + 4; k++) inv[k] = 1.0 / k;
}

struct SphMaps {
  std::vector<int> row_l, row_m;          // coefficient row -> (l, m)
  std::vector<int> slot_l, slot_m, rowc, rows;   // field slot -> l, m, cos row, sin row (-1: no term)
};

// The force loop's slot -> coefficient-row map, following the reference's loop literally
// (src/SphericalBasis.cc:1572-1634), including the fact that EVEN_M / M0_only `continue`
// before moffset advances.
inline SphMaps make_sph_maps(const expb200_sphere_config &cfg)
{
  const int L = cfg.lmax;
  const int nrow = (L + 1) * (L + 1), nslot = (L + 1) * (L + 2) / 2;
  SphMaps M;
  M.row_l.resize(nrow); M.row_m.resize(nrow);
  for (int l = 0; l <= L; l++)
    for (int m = 0; m <= l; m++) {
      int r = sph_row(l, m);
      M.row_l[r] = l; M.row_m[r] = m;
      if (m) { M.row_l[r + 1] = l; M.row_m[r + 1] = m; }
    }
  M.slot_l.resize(nslot); M.slot_m.resize(nslot); M.rowc.assign(nslot, -1); M.rows.assign(nslot, -1);
  for (int l = 0; l <= L; l++)
    for (int m = 0; m <= l; m++) { M.slot_l[sph_slot(L, l, m)] = l; M.slot_m[sph_slot(L, l, m)] = m; }
  if (!cfg.NO_L0) M.rowc[sph_slot(L, 0, 0)] = 0;
  for (int l = 1, loffset = 1; l <= L; loffset += (2 * l + 1), l++) {
    if (cfg.NO_L1 && l == 1) continue;
    if (cfg.EVEN_L && (l / 2) * 2 != l) continue;
    for (int m = 0, moffset = 0; m <= l; m++) {
      if (cfg.EVEN_M && (m / 2) * 2 != m) continue;
      if (cfg.M0_only && m != 0) continue;
      int slot = sph_slot(L, l, m);
      if (m == 0) { M.rowc[slot] = loffset + moffset; moffset++; }
      else { M.rowc[slot] = loffset + moffset; M.rows[slot] = loffset + moffset + 1; moffset += 2; }
    }
  }
  return M;
}

} // namespace expb200
