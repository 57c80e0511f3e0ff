// cylinder.cu -- the cylinder force (Cylinder + EmpCylSL tables) on B200.
//
// Replaces Cylinder::determine_coefficients_cuda / determine_acceleration_cuda /
// DtoH_coefs / HtoD_coefs and EmpCylSL::initialize_cuda (src/cudaCylinder.cu,
// src/cudaEmpCylSL.cu).  Not a port: the reference gathers 4 corners x 234 (702) table
// values per particle from 3-D textures; here a particle touches 13 azimuthal terms on 4
// nodes and the table contractions are dense products per step (see cyl_math.cuh).
//
// Coefficient call: key (level, cell) -> stable sort -> deposit (one warp per chunk of
// sorted particles, fixed order) -> combine (cell partials -> node sums G[k][node]) ->
// contract (cosN/sinN[m][n] = <pot table, G>).
// Force call: field F[node][kind][k] = sum_n table * coef, then a 4-node gather per particle.
#include "common.cuh"
#include "sort.cuh"
#include "cyl_math.cuh"
#include "profile.cuh"

#include <cmath>
#include <vector>

namespace expb200 {
namespace {

constexpr int kCylDepositWarps = 8;

__global__ void __launch_bounds__(256)
cyl_key_kernel(CylGeom g, const rec4 *__restrict__ posm, const unsigned *__restrict__ level,
               size_t n, int use_levels, int maxlev, int mlevel, uint32_t *__restrict__ keys,
               double *__restrict__ block_mass)
{
  // keys for every particle + this CTA's (fixed-order) sum of the masses of the used
  // particles of level mlevel (cylmass0, src/Cylinder.cc:866-867)
  __shared__ double sh[8];
  double msum = 0.0;
  const uint32_t kpl = (uint32_t)g.ncell + 2u;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    rec4 pm = ld_rec(&posm[i]);
    uint32_t ck = cyl_cell_key(g, pm);
    uint32_t lev = 0;
    if (use_levels) {
      lev = level[i];
      if (lev > (uint32_t)maxlev) lev = maxlev;
    }
    keys[i] = lev * kpl + ck;
    if ((int)lev == mlevel && ck <= (uint32_t)g.ncell) msum += pm.w;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) msum += __shfl_down_sync(0xffffffffu, msum, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = msum;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < 8; k++) t += sh[k];
    block_mass[blockIdx.x] = t;
  }
}

__global__ void cyl_mass_kernel(const double *__restrict__ block_mass, int nblocks, double adb,
                                double *__restrict__ out)
{
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < nblocks; k++) t += block_mass[k];
    out[0] = t * adb;
  }
}

// keys for multistep_update: particles whose level changed get (side level, cell with the
// update's own gating), everything else the sentinel key nlev*(ncell+2).
__global__ void __launch_bounds__(256)
cyl_switch_key_kernel(CylGeom g, const rec4 *__restrict__ posm, const unsigned *__restrict__ level,
                      const unsigned *__restrict__ prev, size_t n, int use_prev, int nlev,
                      uint32_t *__restrict__ keys)
{
  const uint32_t kpl = (uint32_t)g.ncell + 2u;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    unsigned a = prev[i], b = level[i];
    uint32_t key = (uint32_t)nlev * kpl;
    if (a != b) {
      uint32_t lev = use_prev ? a : b;
      if (lev >= (uint32_t)nlev) lev = nlev - 1;
      key = lev * kpl + cyl_cell_key(g, ld_rec(&posm[i]), 1);
    }
    keys[i] = key;
  }
}

__global__ void cyl_axpy_kernel(double *__restrict__ dst, const double *__restrict__ src, double sign, size_t n)
{
  size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (k < n) dst[k] += sign * src[k];
}

// deposit: see sphere.cu::sph_deposit_kernel for the scheme.  Here a particle carries nk
// azimuthal terms and four corner weights; lane q owns accumulators (k, corner) = (q>>2, q&3).
__global__ void __launch_bounds__(kCylDepositWarps * 32)
cyl_deposit_kernel(CylGeom g, const rec4 *__restrict__ posm, const uint32_t *__restrict__ perm,
                   const uint32_t *__restrict__ seg, int keybase, int CH, double adb,
                   int stride, int mode, double *__restrict__ PA)
{
  // per warp: 32 rows of `stride` doubles (nk azimuthal terms + 4 corner weights, odd
  // stride => conflict-free 64-bit lanes) + 32 cell ids
  extern __shared__ double cyl_smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double *tile = cyl_smem + (size_t)w * (32 * stride + 16);
  int *cells = reinterpret_cast<int *>(tile + 32 * stride);
  const int nk = g.nk, nacc = 4 * nk;

  const uint32_t lo = seg[keybase], hi = seg[keybase + g.ncell];   // deposited particles only
  const size_t chunk = (size_t)blockIdx.x * kCylDepositWarps + w;
  const size_t beg = (size_t)lo + chunk * (size_t)CH;
  if (beg >= hi) return;
  const size_t end = beg + CH < hi ? beg + CH : hi;

  constexpr int APL = (4 * (2 * kCylMaxM + 1) + 31) / 32;   // accumulators per lane (max)
  double acc[APL];
#pragma unroll
  for (int j = 0; j < APL; j++) acc[j] = 0.0;
  int cur = -1;

  auto flush = [&](int c) {
    size_t slot = (chunk + (size_t)c) * nacc;
#pragma unroll
    for (int j = 0; j < APL; j++) {
      int q = lane + 32 * j;
      if (q < nacc) PA[slot + q] = acc[j];
      acc[j] = 0.0;
    }
  };

  double *my = tile + lane * stride;
  for (size_t k0 = beg; k0 < end; k0 += 32) {
    const size_t k = k0 + lane;
    int cellid = -1;
    bool used = false;
    if (k < end) {
      rec4 pm = ld_rec(&posm[perm[k]]);
      CylCell c;
      used = cyl_deposit_particle(g, pm, adb, c, [&](int kk, double v) { my[kk] = v; }, mode);
      if (used) {
        cellid = c.ix * g.numy + c.iy;
        my[nk] = c.c00; my[nk + 1] = c.c10; my[nk + 2] = c.c01; my[nk + 3] = c.c11;
      }
    }
    if (!used) {
      for (int kk = 0; kk < nk + 4; kk++) my[kk] = 0.0;
    }
    cells[lane] = cellid;
    __syncwarp();

    const int cnt = end - k0 < 32 ? (int)(end - k0) : 32;
    for (int jj = 0; jj < cnt; jj++) {
      const int c = cells[jj];
      if (c < 0) continue;
      if (c != cur) {
        if (cur >= 0) flush(cur);
        cur = c;
      }
      const double *row = tile + jj * stride;
#pragma unroll
      for (int j = 0; j < APL; j++) {
        int q = lane + 32 * j;
        if (q < nacc) acc[j] += row[q >> 2] * row[nk + (q & 3)];
      }
    }
    __syncwarp();
  }
  if (cur >= 0) flush(cur);
}

// combine: G[k][node] = sum over the (up to four) cells that have `node` as a corner of
// that cell's partials for that corner, chunks in increasing order.
__global__ void __launch_bounds__(128)
cyl_combine_kernel(CylGeom g, const uint32_t *__restrict__ seg, int keybase, int CH,
                   const double *__restrict__ PA, double *__restrict__ G)
{
  const int node = blockIdx.x * blockDim.x + threadIdx.x;
  if (node >= g.nnode) return;
  const int jx = node / (g.numy + 1), jy = node % (g.numy + 1);
  const int nk = g.nk, nacc = 4 * nk;
  const uint32_t lo = seg[keybase];
  for (int k = 0; k < nk; k++) {
    double sum = 0.0;
#pragma unroll
    for (int corner = 0; corner < 4; corner++) {
      const int cx = jx - (corner & 1), cy = jy - (corner >> 1);
      if (cx < 0 || cy < 0 || cx >= g.numx || cy >= g.numy) continue;
      const int c = cx * g.numy + cy;
      const uint32_t s0 = seg[keybase + c], s1 = seg[keybase + c + 1];
      if (s1 <= s0) continue;
      const size_t j0 = (s0 - lo) / CH, j1 = (s1 - 1 - lo) / CH;
      double part = 0.0;
      for (size_t j = j0; j <= j1; j++) part += PA[(j + c) * nacc + 4 * k + corner];
      sum += part;
    }
    G[(size_t)k * g.nnode + node] = sum;
  }
}

// contract: coef[cs][m][n] = sum_node pot{C,S}[m][n][node] * G[k][node]; one CTA per (cs,m,n)
__global__ void __launch_bounds__(256)
cyl_contract_kernel(CylGeom g, const double *__restrict__ potC, const double *__restrict__ potS,
                    const double *__restrict__ G, int mlim, int even_m, double *__restrict__ coef)
{
  __shared__ double sh[8];
  const int per = (g.mmax + 1) * g.norder;
  const int id = blockIdx.x;
  const int cs = id / per, mn = id % per, m = mn / g.norder;
  double t = 0.0;
  // rows EmpCylSL::get_pot never fills (m > MLIM, odd m with EVEN_M:
  // exputil/EmpCylSL.cc:5602-5605) stay zero
  const bool off = m > mlim || (even_m && (m & 1));
  if (!(cs == 1 && m == 0) && !off) {
    const double *tab = (cs ? potS : potC) + (size_t)mn * g.nnode;
    const double *Gk = G + (size_t)(cs ? g.mmax + m : m) * g.nnode;
    double s = 0.0;
    for (int i = threadIdx.x; i < g.nnode; i += 256) s += tab[i] * Gk[i];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) for (int k = 0; k < 8; k++) t += sh[k];
  }
  if (threadIdx.x == 0) coef[id] = t;
}

// field: F[node][kind][k] = sum_{n in [nmin,nlim)} table_kind{C,S}[m][n][node] * coef{cos,sin}[m][n]
// for the m the reference's accumulated_eval visits (MMIN/MLIM/EVEN_M), else 0.
struct CylTables { const double *t[3][2]; };
__global__ void __launch_bounds__(128)
cyl_field_kernel(CylGeom g, CylTables T, const double *__restrict__ coef, int mmin, int mlim,
                 int even_m, int nmin, int nlim, int nkpad, double *__restrict__ F)
{
  const int node = blockIdx.x * blockDim.x + threadIdx.x;
  const int kind = blockIdx.y / g.nk, k = blockIdx.y % g.nk;
  if (node >= g.nnode) return;
  const int cs = k > g.mmax, m = cs ? k - g.mmax : k;
  double v = 0.0;
  const bool on = m >= mmin && m <= mlim && !(even_m && (m & 1));
  if (on) {
    const double *tab = T.t[kind][cs] + (size_t)m * g.norder * g.nnode + node;
    const double *c = coef + (size_t)cs * (g.mmax + 1) * g.norder + (size_t)m * g.norder;
    for (int n = nmin; n < nlim; n++) v += tab[(size_t)n * g.nnode] * c[n];
  }
  F[((size_t)node * 3 + kind) * nkpad + k] = v;
}

__global__ void __launch_bounds__(128)
cyl_force_kernel(CylGeom g, const double *__restrict__ F, int nkpad, const double *__restrict__ cylmass,
                 const rec4 *__restrict__ posm, rec4 *__restrict__ accp,
                 const uint32_t *__restrict__ perm, const uint32_t *__restrict__ seg, int keybase, size_t n)
{
  const size_t lo = seg ? seg[keybase] : 0;
  const size_t k = lo + blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (k >= n) return;
  const size_t i = perm ? perm[k] : k;
  rec4 pm = ld_rec(&posm[i]);
  double ax = 0.0, ay = 0.0, az = 0.0, pot = 0.0;
  cyl_force_particle(g, F, nkpad, cylmass[0], pm, ax, ay, az, pot);
  rec4 a = ld_rec_rw(&accp[i]);
  a.x += ax; a.y += ay; a.z += az; a.w += pot;
  st_rec(&accp[i], a);
}

struct InterpArgs { int nlev, mfirst; double a[32], b[32]; };
__global__ void cyl_interp_kernel(const double *__restrict__ cL, const double *__restrict__ cN,
                                  InterpArgs A, size_t ncoef, double *__restrict__ out)
{
  size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (k >= ncoef) return;
  double s = 0.0;
  for (int M = 0; M < A.mfirst; M++) s += A.a[M] * cL[M * ncoef + k] + A.b[M] * cN[M * ncoef + k];
  for (int M = A.mfirst; M < A.nlev; M++) s += cN[M * ncoef + k];
  out[k] = s;
}

int cyl_chunk(size_t n)
{
  size_t q = n / ((size_t)kSMsB200 * 32 * 32);
  if (q < 1) q = 1;
  if (q > 32) q = 32;
  return (int)q * 32;
}

} // namespace
} // namespace expb200

using namespace expb200;

struct expb200_cylinder_s {
  int device = 0;
  expb200_cylinder_config cfg;
  CylGeom geom;
  int nlev = 1;
  int nkpad = 0;
  size_t ncoef = 0;                 // 2*(mmax+1)*norder
  double *tab[3][2] = {};           // [kind][cs] -> [(mmax+1)][norder][nnode]
  double *coef = nullptr;           // accumulated [cos | sin | mass]   (ncoef + 1 doubles)
  double *coef_use = nullptr;       // installed   [cos | sin | cylmass]
  double *coefN = nullptr, *coefL = nullptr;      // [nlev][ncoef+1]
  double *differ = nullptr, *tmp = nullptr;       // level-switch differences [nlev][ncoef+1], scratch
  double *G = nullptr;              // [nk][nnode]
  double *F = nullptr;              // [nnode][3][nkpad]
  double *PA = nullptr;
  size_t part_slots = 0;
  double *block_mass = nullptr;
  bool field_valid = false, have_accum = false;
  cudaStream_t last_stream = nullptr, own_stream = nullptr;
  cudaStream_t work_stream()
  {
    if (last_stream && expb200::stream_is_live(last_stream)) return last_stream;
    last_stream = nullptr;
    return own_stream;
  }
};

namespace {

CylGeom geom_for(const expb200_cylinder_s *s, const expb200_component_s *c)
{
  CylGeom g = s->geom;
  g.cx = c->center[0]; g.cy = c->center[1]; g.cz = c->center[2];
  g.rtrunc2 = c->rtrunc * c->rtrunc;
  return g;
}

int ensure_sorted(expb200_cylinder_s *s, expb200_component_s *c, int mlevel, double adb, bool want_mass)
{
  const uint32_t kpl = (uint32_t)s->geom.ncell + 2u;
  const uint32_t nkeys = (uint32_t)s->nlev * kpl;
  const bool valid = c->sort_kind == 3 && c->sort_owner == s && c->sort_pos_ver == c->pos_ver &&
                     c->sort_lev_ver == c->lev_ver;
  if (valid && !want_mass) return EXPB200_OK;
  if (c->seg_cap < (size_t)nkeys + 2) {
    if (c->seg) cudaFree(c->seg);
    c->seg = nullptr;
    CU_TRY(cudaMalloc(&c->seg, ((size_t)nkeys + 2) * sizeof(uint32_t)));
    c->seg_cap = (size_t)nkeys + 2;
  }
  CylGeom g = geom_for(s, c);
  const unsigned nb = grid_for(c->n ? c->n : 1, 256);
  {
    ProfRegion pr(K_KEY, c->stream, 2);
    cyl_key_kernel<<<nb, 256, 0, c->stream>>>(g, c->posm, c->level, c->n, s->nlev > 1, s->nlev - 1,
                                              s->nlev > 1 ? mlevel : 0, c->keys[valid ? (c->cur ^ 1) : 0],
                                              s->block_mass);
    cyl_mass_kernel<<<1, 32, 0, c->stream>>>(s->block_mass, nb, adb, s->coef + s->ncoef);
    CU_LAUNCH_CHECK();
  }
  if (valid) return EXPB200_OK;      // keys recomputed into the spare buffer only for the mass
  if (c->n) {
    ProfRegion pr(K_SORT, c->stream, radix_sort_launches(c->n, nkeys, false));
    int cur = 0;
    int rc = radix_sort_pairs(c->stream, c->keys, c->perm, c->hist, c->rank, c->n, nkeys, &cur);
    if (rc) return rc;
    c->cur = cur;
  }
  int rc;
  { ProfRegion pr(K_SEG, c->stream);
    rc = segment_offsets(c->stream, c->keys[c->cur], c->n, c->seg, nkeys); }
  if (rc) return rc;
  c->sort_kind = 3;
  c->sort_owner = s;
  c->keys_per_level = kpl;
  c->sort_pos_ver = c->pos_ver;
  c->sort_lev_ver = c->lev_ver;
  return EXPB200_OK;
}

int build_field(expb200_cylinder_s *s, cudaStream_t st)
{
  const CylGeom &g = s->geom;
  CylTables T;
  for (int k = 0; k < 3; k++) for (int cs = 0; cs < 2; cs++) T.t[k][cs] = s->tab[k][cs];
  const int mmin = s->cfg.MMIN > 0 ? s->cfg.MMIN : 0;
  const int mlim = s->cfg.MLIM < g.mmax ? s->cfg.MLIM : g.mmax;
  const int nmin = s->cfg.NMIN > 0 ? s->cfg.NMIN : 0;
  const int nlim = s->cfg.NLIM < g.norder ? s->cfg.NLIM : g.norder;
  dim3 grid(div_up(g.nnode, 128), 3 * g.nk);
  ProfRegion pr(K_CYL_FIELD, st);
  cyl_field_kernel<<<grid, 128, 0, st>>>(g, T, s->coef_use, mmin, mlim, s->cfg.EVEN_M, nmin, nlim,
                                         s->nkpad, s->F);
  CU_LAUNCH_CHECK();
  s->field_valid = true;
  return EXPB200_OK;
}

} // namespace

extern "C" {

int expb200_cylinder_create(int device, const expb200_cylinder_config *cfg,
                            const double *potC, const double *potS, const double *rforceC,
                            const double *rforceS, const double *zforceC, const double *zforceS,
                            expb200_cylinder *out)
{
  REQUIRE(cfg && potC && potS && rforceC && rforceS && zforceC && zforceS && out, "NULL argument");
  REQUIRE(cfg->mmax >= 0 && cfg->mmax <= kCylMaxM, "mmax=%d outside [0,%d]", cfg->mmax, kCylMaxM);
  REQUIRE(cfg->norder >= 1 && cfg->numx >= 1 && cfg->numy >= 1, "bad table shape");
  REQUIRE(cfg->multistep >= 0 && cfg->multistep < 30, "multistep=%d out of range", cfg->multistep);
  REQUIRE(cfg->ascale > 0 && cfg->hscale > 0, "scales must be positive");
  int ndev = 0;
  CU_TRY(cudaGetDeviceCount(&ndev));
  REQUIRE(device >= 0 && device < ndev, "device %d not present (%d visible)", device, ndev);
  DeviceGuard gd(device);

  auto *s = new expb200_cylinder_s;
  s->device = device;
  s->cfg = *cfg;
  s->nlev = cfg->multistep + 1;
  CylGeom &g = s->geom;
  g.mmax = cfg->mmax; g.norder = cfg->norder; g.numx = cfg->numx; g.numy = cfg->numy;
  g.cmapr = cfg->cmapr; g.cmapz = cfg->cmapz;
  g.nk = 2 * cfg->mmax + 1;
  g.ncell = cfg->numx * cfg->numy;
  g.nnode = (cfg->numx + 1) * (cfg->numy + 1);
  g.ascale = cfg->ascale; g.hscale = cfg->hscale;
  // EmpCylSL::setup_table (exputil/EmpCylSL.cc:2130-2137)
  g.rtable = M_SQRT1_2 * cfg->rmax;
  g.cx = g.cy = g.cz = 0.0; g.rtrunc2 = 1.0e300; g.rotate = 0;
  g.xmin = cyl_r_to_xi(g, cfg->rmin * cfg->ascale);
  double xmax = cyl_r_to_xi(g, g.rtable * cfg->ascale);
  g.dx = (xmax - g.xmin) / cfg->numx;
  g.ymin = cyl_z_to_y(g, -g.rtable * cfg->ascale);
  double ymax = cyl_z_to_y(g, g.rtable * cfg->ascale);
  g.dy = (ymax - g.ymin) / cfg->numy;
  g.rmax2 = cfg->rcylmax * cfg->rcylmax * cfg->acyl * cfg->acyl;
  g.r2grid = (cfg->ascale * g.rtable) * (cfg->ascale * g.rtable);
  for (int k = 0; k < 9; k++) g.body[k] = g.orig[k] = (k % 4 == 0) ? 1.0 : 0.0;
  s->nkpad = (g.nk + 1) & ~1;
  s->ncoef = 2 * (size_t)(g.mmax + 1) * g.norder;

  const size_t tbytes = (size_t)(g.mmax + 1) * g.norder * g.nnode * sizeof(double);
  const double *src[3][2] = {{potC, potS}, {rforceC, rforceS}, {zforceC, zforceS}};
  for (int k = 0; k < 3; k++)
    for (int cs = 0; cs < 2; cs++) {
      CU_TRY(cudaMalloc(&s->tab[k][cs], tbytes));
      CU_TRY(cudaMemcpy(s->tab[k][cs], src[k][cs], tbytes, cudaMemcpyHostToDevice));
    }
  const size_t cb = (s->ncoef + 1) * sizeof(double);
  CU_TRY(cudaMalloc(&s->coef, cb));
  CU_TRY(cudaMalloc(&s->coef_use, cb));
  CU_TRY(cudaMalloc(&s->coefN, s->nlev * cb));
  CU_TRY(cudaMalloc(&s->coefL, s->nlev * cb));
  CU_TRY(cudaMemset(s->coef, 0, cb));
  CU_TRY(cudaMemset(s->coef_use, 0, cb));
  CU_TRY(cudaMemset(s->coefN, 0, s->nlev * cb));
  CU_TRY(cudaMemset(s->coefL, 0, s->nlev * cb));
  CU_TRY(cudaMalloc(&s->differ, s->nlev * cb));
  CU_TRY(cudaMalloc(&s->tmp, cb));
  CU_TRY(cudaMemset(s->differ, 0, s->nlev * cb));
  CU_TRY(cudaMalloc(&s->G, (size_t)g.nk * g.nnode * sizeof(double)));
  CU_TRY(cudaMalloc(&s->F, (size_t)g.nnode * 3 * s->nkpad * sizeof(double)));
  CU_TRY(cudaMalloc(&s->block_mass, (size_t)kSMsB200 * 8 * sizeof(double)));
  CU_TRY(cudaStreamCreateWithFlags(&s->own_stream, cudaStreamNonBlocking));
  *out = s;
  return EXPB200_OK;
}

int expb200_cylinder_destroy(expb200_cylinder s)
{
  if (!s) return EXPB200_OK;
  DeviceGuard g(s->device);
  cudaDeviceSynchronize();
  for (int k = 0; k < 3; k++) for (int cs = 0; cs < 2; cs++) cudaFree(s->tab[k][cs]);
  cudaFree(s->coef); cudaFree(s->coef_use); cudaFree(s->coefN); cudaFree(s->coefL);
  cudaFree(s->differ); cudaFree(s->tmp);
  cudaFree(s->G); cudaFree(s->F); cudaFree(s->PA); cudaFree(s->block_mass);
  if (s->own_stream) cudaStreamDestroy(s->own_stream);
  delete s;
  return EXPB200_OK;
}

int expb200_cylinder_set_orientation(expb200_cylinder s, const double *body, const double *orig)
{
  REQUIRE(s, "cylinder is NULL");
  if (body && orig) {
    for (int k = 0; k < 9; k++) { s->geom.body[k] = body[k]; s->geom.orig[k] = orig[k]; }
    s->geom.rotate = 1;
  } else {
    s->geom.rotate = 0;
  }
  return EXPB200_OK;
}

int expb200_cylinder_determine_coefficients(expb200_cylinder s, expb200_component c, int mlevel,
                                            double adb, double *cos_host, double *sin_host,
                                            uint64_t *used, double *mass)
{
  REQUIRE(s && c, "NULL argument");
  REQUIRE(s->device == c->device, "device mismatch");
  REQUIRE(mlevel >= 0 && mlevel < s->nlev, "mlevel=%d outside [0,%d]", mlevel, s->nlev - 1);
  DeviceGuard gd(s->device);
  cudaStream_t st = c->stream;
  { cudaStream_t ws = s->work_stream(); if (ws != st) CU_TRY(cudaStreamSynchronize(ws)); }
  s->last_stream = st;
  int rc = ensure_sorted(s, c, mlevel, adb, true);
  if (rc) return rc;
  CylGeom g = geom_for(s, c);
  g.rotate = s->geom.rotate;
  const int CH = cyl_chunk(c->n);
  const size_t slots = (c->n + CH - 1) / CH + (size_t)g.ncell + 1;
  if (slots > s->part_slots) {
    if (s->PA) cudaFree(s->PA);
    s->PA = nullptr; s->part_slots = 0;
    CU_TRY(cudaMalloc(&s->PA, slots * 4 * g.nk * sizeof(double)));
    s->part_slots = slots;
  }
  const int keybase = (s->nlev > 1 ? mlevel : 0) * (g.ncell + 2);
  if (c->n) {
    ProfRegion pr(K_CYL_DEPOSIT, st);
    unsigned nchunks = div_up(c->n, CH);
    const int stride = (g.nk + 4) | 1;
    const size_t sm = (size_t)kCylDepositWarps * (32 * stride + 16) * sizeof(double);
    cyl_deposit_kernel<<<div_up(nchunks, kCylDepositWarps), kCylDepositWarps * 32, sm, st>>>(
        g, c->posm, c->perm[c->cur], c->seg, keybase, CH, adb, stride, 0, s->PA);
    CU_LAUNCH_CHECK();
  }
  { ProfRegion pr(K_COMBINE, st);
    cyl_combine_kernel<<<div_up(g.nnode, 128), 128, 0, st>>>(g, c->seg, keybase, CH, s->PA, s->G); }
  { ProfRegion pr(K_CONTRACT, st);
    cyl_contract_kernel<<<(unsigned)s->ncoef, 256, 0, st>>>(g, s->tab[0][0], s->tab[0][1], s->G,
                                                            s->cfg.MLIM, s->cfg.EVEN_M, s->coef); }
  CU_LAUNCH_CHECK();
  s->have_accum = true;
  const size_t half = s->ncoef / 2;
  if (cos_host) CU_TRY(cudaMemcpyAsync(cos_host, s->coef, half * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (sin_host) CU_TRY(cudaMemcpyAsync(sin_host, s->coef + half, half * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (mass) CU_TRY(cudaMemcpyAsync(mass, s->coef + s->ncoef, sizeof(double), cudaMemcpyDeviceToHost, st));
  uint32_t lohi[2] = {0, 0};
  if (used) {
    CU_TRY(cudaMemcpyAsync(&lohi[0], c->seg + keybase, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(&lohi[1], c->seg + keybase + g.ncell + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  }
  if (cos_host || sin_host || mass || used) CU_TRY(cudaStreamSynchronize(st));
  if (used) *used = lohi[1] - lohi[0];
  return EXPB200_OK;
}

int expb200_cylinder_coef_device(expb200_cylinder s, double **dptr, size_t *count, void **stream)
{
  REQUIRE(s && dptr, "NULL argument");
  *dptr = s->coef;
  if (count) *count = s->ncoef + 1;      // cos | sin | mass: one all-reduce covers them all
  if (stream) *stream = s->work_stream();
  return EXPB200_OK;
}

int expb200_cylinder_set_coefficients(expb200_cylinder s, const double *cos_host,
                                      const double *sin_host, double cylmass)
{
  REQUIRE(s, "cylinder is NULL");
  DeviceGuard gd(s->device);
  cudaStream_t st = s->work_stream();
  const size_t half = s->ncoef / 2;
  if (cos_host && sin_host) {
    CU_TRY(cudaStreamSynchronize(st));
    CU_TRY(cudaMemcpy(s->coef_use, cos_host, half * sizeof(double), cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(s->coef_use + half, sin_host, half * sizeof(double), cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(s->coef_use + s->ncoef, &cylmass, sizeof(double), cudaMemcpyHostToDevice));
  } else {
    REQUIRE(s->have_accum, "no coefficients accumulated on the device yet");
    CU_TRY(cudaMemcpyAsync(s->coef_use, s->coef, (s->ncoef + 1) * sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (cylmass >= 0.0) {
      CU_TRY(cudaStreamSynchronize(st));
      CU_TRY(cudaMemcpy(s->coef_use + s->ncoef, &cylmass, sizeof(double), cudaMemcpyHostToDevice));
    }
  }
  s->field_valid = false;
  return EXPB200_OK;
}

int expb200_cylinder_determine_acceleration(expb200_cylinder s, expb200_component c, int mlevel)
{
  REQUIRE(s && c, "NULL argument");
  REQUIRE(s->device == c->device, "device mismatch");
  REQUIRE(mlevel >= 0 && mlevel < s->nlev, "mlevel=%d outside [0,%d]", mlevel, s->nlev - 1);
  DeviceGuard gd(s->device);
  if (c->n == 0) return EXPB200_OK;
  cudaStream_t st = c->stream;
  { cudaStream_t ws = s->work_stream(); if (ws != st) CU_TRY(cudaStreamSynchronize(ws)); }
  s->last_stream = st;
  int rc = ensure_sorted(s, c, mlevel, 1.0, false);
  if (rc) return rc;
  if (!s->field_valid && (rc = build_field(s, st))) return rc;
  CylGeom g = geom_for(s, c);
  g.rotate = s->geom.rotate;
  const int keybase = (s->nlev > 1 ? mlevel : 0) * (g.ncell + 2);
  ProfRegion pr(K_CYL_FORCE, st);
  cyl_force_kernel<<<div_up(c->n, 128), 128, 0, st>>>(g, s->F, s->nkpad, s->coef_use + s->ncoef, c->posm,
                                                      c->accp, c->perm[c->cur], c->seg, keybase, c->n);
  CU_LAUNCH_CHECK();
  return EXPB200_OK;
}

// multistep tables, as for the sphere (src/CylEXP.cc:192-282, exputil/EmpCylSL.cc:2004-2014)
int expb200_cylinder_multistep_commit(expb200_cylinder s, int mlevel)
{
  REQUIRE(s && mlevel >= 0 && mlevel < s->nlev, "bad argument");
  REQUIRE(s->have_accum, "no coefficients accumulated yet");
  DeviceGuard gd(s->device);
  cudaStream_t st = s->work_stream();
  const size_t n1 = s->ncoef + 1;
  double *N = s->coefN + (size_t)mlevel * n1, *L = s->coefL + (size_t)mlevel * n1;
  CU_TRY(cudaMemcpyAsync(L, N, n1 * sizeof(double), cudaMemcpyDeviceToDevice, st));
  CU_TRY(cudaMemcpyAsync(N, s->coef, n1 * sizeof(double), cudaMemcpyDeviceToDevice, st));
  return EXPB200_OK;
}

int expb200_cylinder_compute_multistep_coefficients(expb200_cylinder s, int mdrft, double cylmass)
{
  REQUIRE(s, "cylinder is NULL");
  const int multistep = s->nlev - 1, Mstep = 1 << multistep;
  REQUIRE(mdrft >= 0 && mdrft <= Mstep, "mdrft=%d outside [0,%d]", mdrft, Mstep);
  DeviceGuard gd(s->device);
  InterpArgs A;
  A.nlev = s->nlev;
  A.mfirst = 0;
  for (int lev = 0; lev <= multistep; lev++)
    if (mdrft == 0 || mdrft % (1 << (multistep - lev)) == 0) { A.mfirst = lev; break; }
  for (int M = 0; M < A.mfirst; M++) {
    int d = 1 << (multistep - M);
    int dL = (mdrft / d) * d, dN = dL + d;
    double b = (double)(mdrft - dL) / (double)(dN - dL);
    A.a[M] = 1.0 - b;
    A.b[M] = b;
  }
  cudaStream_t st = s->work_stream();
  cyl_interp_kernel<<<div_up(s->ncoef + 1, 256), 256, 0, st>>>(s->coefL, s->coefN, A, s->ncoef + 1, s->coef_use);
  CU_LAUNCH_CHECK();
  if (cylmass >= 0.0) {
    CU_TRY(cudaStreamSynchronize(st));
    CU_TRY(cudaMemcpy(s->coef_use + s->ncoef, &cylmass, sizeof(double), cudaMemcpyHostToDevice));
  }
  s->field_valid = false;
  return EXPB200_OK;
}

// ~ CylEXP::multistep_update_begin / multistep_update (per switching particle) /
// multistep_update_finish (src/CylEXP.cc:45-157, 159-188; src/Cylinder.cc:1751-1773): for every
// particle whose level changed in the last expb200_adjust_multistep_level, subtract its
// contribution from differ[old level] and add it to differ[new level]; with finish != 0 also
// cosN/sinN[M] += differ[M] (single rank; with several ranks all-reduce the differ buffer
// first, then call expb200_cylinder_multistep_update_finish).
int expb200_cylinder_multistep_update(expb200_cylinder s, expb200_component c, double adb, int finish)
{
  REQUIRE(s && c, "NULL argument");
  REQUIRE(s->device == c->device, "device mismatch");
  DeviceGuard gd(s->device);
  cudaStream_t st = c->stream;
  { cudaStream_t ws = s->work_stream(); if (ws != st) CU_TRY(cudaStreamSynchronize(ws)); }
  s->last_stream = st;
  const size_t n1 = s->ncoef + 1, all = (size_t)s->nlev * n1;
  CU_TRY(cudaMemsetAsync(s->differ, 0, all * sizeof(double), st));
  if (c->n && s->nlev > 1) {
    CylGeom g = geom_for(s, c);
    g.rotate = 0;                                   // the update uses the unrotated position
    const uint32_t kpl = (uint32_t)g.ncell + 2u;
    const uint32_t nkeys = (uint32_t)s->nlev * kpl + 1;
    if (c->seg_cap < (size_t)nkeys + 2) {
      if (c->seg) cudaFree(c->seg);
      c->seg = nullptr;
      CU_TRY(cudaMalloc(&c->seg, ((size_t)nkeys + 2) * sizeof(uint32_t)));
      c->seg_cap = (size_t)nkeys + 2;
    }
    const int CH = cyl_chunk(c->n);
    const size_t slots = (c->n + CH - 1) / CH + (size_t)g.ncell + 1;
    if (slots > s->part_slots) {
      if (s->PA) cudaFree(s->PA);
      s->PA = nullptr; s->part_slots = 0;
      CU_TRY(cudaMalloc(&s->PA, slots * 4 * g.nk * sizeof(double)));
      s->part_slots = slots;
    }
    const int stride = (g.nk + 4) | 1;
    const size_t sm = (size_t)kCylDepositWarps * (32 * stride + 16) * sizeof(double);
    const unsigned nchunks = div_up(c->n, CH);
    for (int side = 0; side < 2; side++) {          // 0: "to" (+), 1: "from" (-)
      cyl_switch_key_kernel<<<grid_for(c->n, 256), 256, 0, st>>>(g, c->posm, c->level, c->prev_level, c->n,
                                                                 side, s->nlev, c->keys[0]);
      CU_LAUNCH_CHECK();
      int cur = 0;
      int rc = radix_sort_pairs(st, c->keys, c->perm, c->hist, c->rank, c->n, nkeys, &cur, c->seg);
      if (rc) return rc;
      c->cur = cur;
      for (int M = 0; M < s->nlev; M++) {
        const int keybase = M * (int)kpl;
        cyl_deposit_kernel<<<div_up(nchunks, kCylDepositWarps), kCylDepositWarps * 32, sm, st>>>(
            g, c->posm, c->perm[cur], c->seg, keybase, CH, adb, stride, 1, s->PA);
        cyl_combine_kernel<<<div_up(g.nnode, 128), 128, 0, st>>>(g, c->seg, keybase, CH, s->PA, s->G);
        // rows get_pot does not fill (m > MLIM, odd m with EVEN_M) stay zero, as in the
        // coefficient pass (src/CylEXP.cc:170-186 loops mm = 0..MMAX over what get_pot filled)
        cyl_contract_kernel<<<(unsigned)s->ncoef, 256, 0, st>>>(g, s->tab[0][0], s->tab[0][1], s->G,
                                                                s->cfg.MLIM, s->cfg.EVEN_M, s->tmp);
        cyl_axpy_kernel<<<div_up(s->ncoef, 256), 256, 0, st>>>(s->differ + (size_t)M * n1, s->tmp,
                                                              side == 0 ? 1.0 : -1.0, s->ncoef);
        CU_LAUNCH_CHECK();
      }
    }
    c->sort_kind = 0;          // the component's regular sort was overwritten
  }
  if (finish) return expb200_cylinder_multistep_update_finish(s);
  return EXPB200_OK;
}

int expb200_cylinder_differ_device(expb200_cylinder s, double **dptr, size_t *count, void **stream)
{
  REQUIRE(s && dptr, "NULL argument");
  *dptr = s->differ;
  if (count) *count = (size_t)s->nlev * (s->ncoef + 1);
  if (stream) *stream = s->work_stream();
  return EXPB200_OK;
}

int expb200_cylinder_multistep_update_finish(expb200_cylinder s)
{
  REQUIRE(s, "cylinder is NULL");
  DeviceGuard gd(s->device);
  const size_t all = (size_t)s->nlev * (s->ncoef + 1);
  cyl_axpy_kernel<<<div_up(all, 256), 256, 0, s->work_stream()>>>(s->coefN, s->differ, 1.0, all);
  CU_LAUNCH_CHECK();
  return EXPB200_OK;
}

// Host access to the per-level tables: which = 0: cosN/sinN, 1: cosL/sinL, 2: differ;
// cos_host, sin_host [(mmax+1)][norder].
int expb200_cylinder_get_level_coefficients(expb200_cylinder s, int which, int mlevel, double *cos_host,
                                            double *sin_host)
{
  REQUIRE(s && cos_host && sin_host && which >= 0 && which <= 4, "bad argument");
  REQUIRE(which >= 3 || (mlevel >= 0 && mlevel < s->nlev), "mlevel=%d outside [0,%d)", mlevel, s->nlev);
  DeviceGuard gd(s->device);
  CU_TRY(cudaStreamSynchronize(s->work_stream()));
  const size_t n1 = s->ncoef + 1, half = s->ncoef / 2;
  const double *src = which == 3 ? s->coef_use : which == 4 ? s->coef
                    : (which == 0 ? s->coefN : which == 1 ? s->coefL : s->differ) + (size_t)mlevel * n1;
  CU_TRY(cudaMemcpy(cos_host, src, half * sizeof(double), cudaMemcpyDeviceToHost));
  CU_TRY(cudaMemcpy(sin_host, src + half, half * sizeof(double), cudaMemcpyDeviceToHost));
  return EXPB200_OK;
}

int expb200_cylinder_get_mass(expb200_cylinder s, int which, int mlevel, double *mass)
{
  REQUIRE(s && mass && which >= 0 && which <= 4, "bad argument");
  REQUIRE(which >= 3 || (mlevel >= 0 && mlevel < s->nlev), "mlevel=%d outside [0,%d)", mlevel, s->nlev);
  DeviceGuard gd(s->device);
  CU_TRY(cudaStreamSynchronize(s->work_stream()));
  const size_t n1 = s->ncoef + 1;
  const double *src = which == 3 ? s->coef_use : which == 4 ? s->coef
                    : (which == 0 ? s->coefN : which == 1 ? s->coefL : s->differ) + (size_t)mlevel * n1;
  CU_TRY(cudaMemcpy(mass, src + s->ncoef, sizeof(double), cudaMemcpyDeviceToHost));
  return EXPB200_OK;
}

} // extern "C"
