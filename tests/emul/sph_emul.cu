// sph_emul.cu -- TEST INFRASTRUCTURE: runs the *same* __host__ __device__ per-particle
// arithmetic as the CUDA kernels (exp_b200/csrc/sph_math.cuh) in plain host loops, so
// that the factorised evaluation can be checked against the oracle on a CPU-only box.
// It emulates the data flow of sphere.cu (stable sort by radial cell, ordered per-cell
// sums, contraction, field table, stencil force) but none of its warp choreography.
// Never shipped or loaded by exp_b200/.
#include "../../exp_b200/csrc/sph_host.h"

#include <algorithm>
#include <numeric>

using namespace expb200;

namespace {
template <int L>
void coef_impl(const SphGeom &g, const double *K, const double *inv, const double *p0,
               const double *Ep, long n, const double *pos, const double *mass, double adb,
               int M0_only, double *coef, long *used)
{
  const int nrow = g.nrow;
  std::vector<uint32_t> cell(n);
  for (long i = 0; i < n; i++)
    cell[i] = sph_cell_of(g, rec4{pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], mass[i]});
  std::vector<long> perm(n);
  std::iota(perm.begin(), perm.end(), 0L);
  std::stable_sort(perm.begin(), perm.end(), [&](long a, long b) { return cell[a] < cell[b]; });

  std::vector<double> G((size_t)nrow * g.numr, 0.0), yv(nrow);
  long nused = 0;
  for (long k = 0; k < n; k++) {
    long i = perm[k];
    if ((int)cell[i] == g.numr - 1) continue;
    const int c = (int)cell[i];
    rec4 pm{pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], mass[i]};
    SphDep d;
    sph_deposit_setup(g, p0, pm, adb, c, d);
    sph_deposit_rows<L, 0, L>(d, inv, [&](int l, int m, int cs, double v) { yv[sph_row(l, m) + cs] = v; });
    nused++;
    for (int r = 0; r < nrow; r++) {
      G[(size_t)r * g.numr + c] += d.w1 * yv[r];
      G[(size_t)r * g.numr + c + 1] += d.w2 * yv[r];
    }
  }
  SphMaps M;
  expb200_sphere_config cfg{};
  cfg.lmax = g.lmax;
  M = make_sph_maps(cfg);
  for (int row = 0; row < nrow; row++) {
    int l = M.row_l[row], m = M.row_m[row];
    for (int nn = 0; nn < g.nmax; nn++) {
      const double *E = Ep + ((size_t)l * g.nmax + nn) * g.numr;
      double s = 0.0;
      for (int i = 0; i < g.numr; i++) s += E[i] * G[(size_t)row * g.numr + i];
      coef[(size_t)row * g.nmax + nn] = (M0_only && m) ? 0.0 : K[l * (kMaxL + 1) + m] * s;
    }
  }
  *used = nused;
}

template <int L>
void accel_impl(const SphGeom &g, const double *K, const double *inv, const double *p0,
                const double2 *T, long n, const double *pos, const double *mass, double *acc,
                double *pot)
{
  for (long i = 0; i < n; i++) {
    rec4 pm{pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], mass[i]};
    double ax = 0, ay = 0, az = 0, pp = 0;
    if (!sph_force_particle<L, true>(g, inv, T, pm, ax, ay, az, pp)) continue;
    acc[3 * i] += ax; acc[3 * i + 1] += ay; acc[3 * i + 2] += az; pot[i] += pp;
  }
}
} // namespace

#define DISPATCH(lmax, F, ...)                                              \
  switch (lmax) {                                                           \
    case 0: F<0>(__VA_ARGS__); break; case 1: F<1>(__VA_ARGS__); break;     \
    case 2: F<2>(__VA_ARGS__); break; case 3: F<3>(__VA_ARGS__); break;     \
    case 4: F<4>(__VA_ARGS__); break; case 5: F<5>(__VA_ARGS__); break;     \
    case 6: F<6>(__VA_ARGS__); break; case 7: F<7>(__VA_ARGS__); break;     \
    case 8: F<8>(__VA_ARGS__); break; default: return -1;                   \
  }

extern "C" {

int emul_sph_coefficients(const expb200_sphere_config *cfg, const double *ev, const double *ef,
                          const double *p0, long n, const double *pos, const double *mass,
                          const double *center, double rtrunc, double adb, double *coef, long *used)
{
  SphGeom g = make_sph_geom(*cfg);
  g.cx = center[0]; g.cy = center[1]; g.cz = center[2]; g.rtrunc2 = rtrunc * rtrunc;
  std::vector<double> K, inv;
  make_sph_constants(K, inv);
  const int L = cfg->lmax, N = cfg->nmax, R = cfg->numr;
  std::vector<double> Ep((size_t)(L + 1) * N * R);
  for (int l = 0; l <= L; l++)
    for (int nn = 0; nn < N; nn++) {
      double q = std::sqrt(ev[l * N + nn]);
      for (int i = 0; i < R; i++) Ep[((size_t)l * N + nn) * R + i] = ef[((size_t)l * N + nn) * R + i] / q;
    }
  DISPATCH(L, coef_impl, g, K.data(), inv.data(), p0, Ep.data(), n, pos, mass, adb, cfg->M0_only, coef, used)
  return 0;
}

int emul_sph_accel(const expb200_sphere_config *cfg, const double *ev, const double *ef,
                   const double *p0, const double *coef, long n, const double *pos,
                   const double *mass, const double *center, double rtrunc, double *acc, double *pot)
{
  SphGeom g = make_sph_geom(*cfg);
  g.cx = center[0]; g.cy = center[1]; g.cz = center[2]; g.rtrunc2 = rtrunc * rtrunc;
  std::vector<double> K, inv;
  make_sph_constants(K, inv);
  SphMaps M = make_sph_maps(*cfg);
  const int L = cfg->lmax, N = cfg->nmax, R = cfg->numr;
  std::vector<double2> H((size_t)R * g.nslot), T((size_t)R * g.nslot * kSphPoly);
  for (int i = 0; i < R; i++)
    for (int slot = 0; slot < g.nslot; slot++) {
      int l = M.slot_l[slot], m = M.slot_m[slot], rc = M.rowc[slot], rs = M.rows[slot];
      double c = 0, s = 0;
      if (rc >= 0)
        for (int nn = 0; nn < N; nn++) {
          double e = ef[((size_t)l * N + nn) * R + i] / std::sqrt(ev[l * N + nn]);
          c += e * coef[(size_t)rc * N + nn];
          if (rs >= 0) s += e * coef[(size_t)rs * N + nn];
        }
      const double Kf = K[l * (kMaxL + 1) + m];
      H[(size_t)i * g.nslot + slot] = make_double2(Kf * c, Kf * s);
    }
  const double2 zero = make_double2(0.0, 0.0);
  for (int i = 0; i < R; i++)
    for (int slot = 0; slot < g.nslot; slot++) {
      double2 Hm = i > 0 ? H[(size_t)(i - 1) * g.nslot + slot] : zero;
      double2 Hp = i < R - 1 ? H[(size_t)(i + 1) * g.nslot + slot] : zero;
      sph_poly_entry(g, p0, i, Hm, H[(size_t)i * g.nslot + slot], Hp,
                     &T[((size_t)i * g.nslot + slot) * kSphPoly]);
    }
  DISPATCH(L, accel_impl, g, K.data(), inv.data(), p0, T.data(), n, pos, mass, acc, pot)
  return 0;
}
}
