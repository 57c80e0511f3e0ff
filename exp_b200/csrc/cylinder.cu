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

// the same for the tail of the sorted permutation (levels >= m), keys relative to level m
__global__ void __launch_bounds__(256)
cyl_key_sub_kernel(CylGeom g, const rec4 *__restrict__ posm, const unsigned *__restrict__ level,
                   const uint32_t *__restrict__ slots, size_t nsub, int m, int maxlev, int mlevel,
                   uint32_t *__restrict__ keys, double *__restrict__ block_mass)
{
  __shared__ double sh[8];
  double msum = 0.0;
  const uint32_t kpl = (uint32_t)g.ncell + 2u;
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < nsub;
       k += (size_t)gridDim.x * blockDim.x) {
    const uint32_t i = slots[k];
    rec4 pm = ld_rec(&posm[i]);
    uint32_t ck = cyl_cell_key(g, pm);
    uint32_t lev = level[i];
    if (lev > (uint32_t)maxlev) lev = maxlev;
    if (lev < (uint32_t)m) lev = m;            // cannot happen while the partition holds
    keys[k] = (lev - (uint32_t)m) * kpl + ck;
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

// The same keys for a compacted list of switching particles (list[k] = store index): no sentinel.
__global__ void __launch_bounds__(256)
cyl_switch_key_list_kernel(CylGeom g, const rec4 *__restrict__ posm, const unsigned *__restrict__ level,
                           const unsigned *__restrict__ prev, const uint32_t *__restrict__ list, size_t nsw,
                           int use_prev, int nlev, uint32_t *__restrict__ keys)
{
  const uint32_t kpl = (uint32_t)g.ncell + 2u;
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < nsw;
       k += (size_t)gridDim.x * blockDim.x) {
    const uint32_t i = list[k];
    uint32_t lev = use_prev ? prev[i] : level[i];
    if (lev >= (uint32_t)nlev) lev = nlev - 1;
    keys[k] = lev * kpl + cyl_cell_key(g, ld_rec(&posm[i]), 1);
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

// ===========================================================================
// Covariance (`covar`), EmpCylSL::accumulate (exputil/EmpCylSL.cc:4069-4074, 4123-4141).
// Per jackknife bin T = indx % sampT and harmonic m the reference accumulates
//   VC[T][m]   += mass vec,            vec_n = norm [(Vc_n cos + Vs_n sin) + i (Vc_n sin - Vs_n cos)]
//   MV[T][m]   += mass vec vec^H       (Hermitian norder x norder)
// with Vc/Vs the bilinear table values of get_pot (Vs = 0 for m = 0).  Unlike the
// coefficients this is quadratic in the particle, so there is nothing to factorise across
// particles: it is a genuine SYRK.  With u = sqrt(mass) [Re vec ; Im vec] (2 norder reals)
//   U U^T = [[aa^T, ab^T],[ba^T, bb^T]],   Re MV = aa^T + bb^T,   Im MV = ba^T - ab^T
// so one real rank-k update per (T, m) on the FP64 tensor cores (DMMA m8n8k4) gives both.
//
// Particles are sorted by (T, cell).  A CTA owns a chunk of consecutive sorted particles and
// the harmonics m = 8 blockIdx.y + warp; it walks the chunk in tiles of 16 particles of one
// bin: every thread evaluates one (m, n) entry of u for the tile's particles from the
// node-major complex pot table Zt[node][m][n] = (potC, potS) (4 coalesced 16-byte loads per
// particle), then warp m multiplies its 2 norder x 16 panel into register accumulators.
// When the bin changes the accumulators go to partial slot (chunk + T) -- unique, because
// both are non-decreasing along the sorted order -- and a second kernel adds a bin's
// partials in chunk order: no atomics, bit-reproducible.
// ===========================================================================
constexpr int kCovLDpad = 4;          // leading dimension of a panel row = tile + 4 doubles: the 64-bit
                                      // fragment loads of a half warp then fall into 16 distinct banks
constexpr int kCovMG = 7;             // harmonics per CTA: one DMMA warp each (warps 0-6)
constexpr int kCovPanel = 512;        // panel threads (warps 8-23); warp 7 stages the geometry
constexpr int kCovThreads = 768;
constexpr int kCovTile = 32;          // particles per tile
constexpr int kCovSplit = 4;          // particle interleave of the panel evaluation (max)

__global__ void __launch_bounds__(256)
cyl_cov_table_kernel(CylGeom g, const double *__restrict__ potC, const double *__restrict__ potS,
                     int mlim, int even_m, double2 *__restrict__ Zt)
{
  const int MN = (g.mmax + 1) * g.norder;
  const size_t total = (size_t)g.nnode * MN;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total;
       e += (size_t)gridDim.x * blockDim.x) {
    const int node = (int)(e / MN), mn = (int)(e % MN), m = mn / g.norder;
    // rows get_pot never fills (m > MLIM, odd m with EVEN_M) count as zero
    const bool off = m > mlim || (even_m && (m & 1));
    double2 z = make_double2(0.0, 0.0);
    if (!off) {
      z.x = potC[(size_t)mn * g.nnode + node];
      if (m) z.y = potS[(size_t)mn * g.nnode + node];
    }
    Zt[e] = z;
  }
}

__global__ void __launch_bounds__(256)
cyl_cov_key_kernel(CylGeom g, const rec4 *__restrict__ posm, const unsigned *__restrict__ level,
                   const uint32_t *__restrict__ slot_of, size_t n, int use_levels, int mlevel,
                   uint32_t sampT, uint64_t indx_base, uint32_t *__restrict__ keys)
{
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    uint32_t key = sampT * (uint32_t)g.ncell;
    if (!use_levels || level[i] == (unsigned)mlevel) {
      const uint32_t ck = cyl_cell_key(g, ld_rec(&posm[i]));
      if (ck < (uint32_t)g.ncell) {
        const uint64_t indx = indx_base + (slot_of ? slot_of[i] : (uint32_t)i);
        key = (uint32_t)(indx % sampT) * (uint32_t)g.ncell + ck;
      }
    }
    keys[i] = key;
  }
}

// bstart[T] = first sorted position whose bin is >= T, T in [0, sampT]
__global__ void cyl_cov_bounds_kernel(const uint32_t *__restrict__ skeys, size_t n, uint32_t ncell,
                                      int sampT, uint32_t *__restrict__ bstart)
{
  const int T = blockIdx.x * blockDim.x + threadIdx.x;
  if (T > sampT) return;
  const uint64_t want = (uint64_t)T * ncell;
  size_t lo = 0, hi = n;
  while (lo < hi) {
    size_t mid = (lo + hi) / 2;
    if ((uint64_t)skeys[mid] < want) lo = mid + 1; else hi = mid;
  }
  bstart[T] = (uint32_t)lo;
}

// Geometry of the used particles in sorted order, one thread per particle (gather of the position
// record, coordinate maps, cell, corner weights, azimuthal factors): GW[k][4] corner weights times
// -4 pi sqrt(mass), GT[k][M1][2] (cos, sin) m phi, GS[k] sqrt(mass), GM[k] mass, GN[k] node index.
__global__ void __launch_bounds__(256)
cyl_cov_geom_kernel(CylGeom g, const rec4 *__restrict__ posm, const uint32_t *__restrict__ perm,
                    size_t nused, double adb, double *__restrict__ GW, double *__restrict__ GT,
                    double *__restrict__ GS, double *__restrict__ GM, int *__restrict__ GN)
{
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nused) return;
  const int M1 = g.mmax + 1;
  const rec4 pm = ld_rec(&posm[perm[k]]);
  double x, y, z, r;
  cyl_classify(g, pm, x, y, z, r, 0);
  double zz = z;
  if (zz / g.ascale > g.rtable) zz = g.rtable * g.ascale;       // EmpCylSL::get_pot clamp
  if (zz / g.ascale < -g.rtable) zz = -g.rtable * g.ascale;
  const CylCell c = cyl_locate(g, r, zz);
  const double mass = pm.w * adb;
  const double sq = sqrt(mass), sw = -4.0 * M_PI * sq;
  double4 wv;
  wv.x = sw * c.c00; wv.y = sw * c.c10; wv.z = sw * c.c01; wv.w = sw * c.c11;
  reinterpret_cast<double4 *>(GW)[k] = wv;
  GN[k] = c.ix * (g.numy + 1) + c.iy;
  GS[k] = sq;
  GM[k] = mass;
  double2 *trg = reinterpret_cast<double2 *>(GT) + k * M1;
  cyl_trig(g.mmax, x, y, r, [&](int m, double cm, double sm) {
    double2 v; v.x = cm; v.y = sm;
    trg[m] = v;
  });
}

// The three stages of a tile of 32 sorted particles -- staging of its geometry records in shared
// memory, the panel u[m][n][j] (8 KB of table gathers per particle, L1/L2), and the rank-k update
// (DMMA) -- run as a software pipeline over the tiles of the chunk with one CTA barrier per step: in
// step i warp 7 copies the geometry of tile i, warps 8-23 evaluate the panel of tile i-1 into one
// half of a double-buffered U, and warps 0-6 (one harmonic each) multiply the panel of tile i-2 from
// the other half.  The DMMA warps hold the accumulators (2 NT^2 doubles) and take REGD registers
// each (setmaxnreg), the sixteen panel warps REGP: the panel stage is bound by instruction issue,
// not by memory (round 2, 1e7 particles, sampT = 100, profiles/r02p_*: all stages in sequence in
// the same 8 warps 16.0 ms; geometry in a producer warp 13.2 ms; this pipeline with 4 panel warps
// 12.6 ms, the panel stage alone 11.2 ms at 0.25 instructions per cycle and scheduler).
template <int NT, int REGD, int REGP>
__global__ void __launch_bounds__(kCovThreads, 1)
cyl_cov_kernel(CylGeom g, const double2 *__restrict__ Zt, const double *__restrict__ GW,
               const double *__restrict__ GT, const double *__restrict__ GS, const double *__restrict__ GM,
               const int *__restrict__ GN, const uint32_t *__restrict__ bstart, int sampT, int CH, int nsplit,
               int MG, double *__restrict__ PC, double *__restrict__ PV, double *__restrict__ PM)
{
  extern __shared__ double cov_smem[];
  constexpr int P = kCovTile, LD = P + kCovLDpad;
  const int N = g.norder, R2 = 2 * N, M1 = g.mmax + 1, MN = M1 * N;
  const int mg0 = blockIdx.y * MG;                             // MG <= kCovMG harmonics per CTA
  const int nml = M1 - mg0 < MG ? M1 - mg0 : MG;               // harmonics of this CTA
  const int items = nml * N;                                   // <= kCovPanel
  const size_t u_doubles = (size_t)MG * R2 * LD;
  double *geo = cov_smem + 2 * u_doubles;                      // after the two panels
  const size_t geo_doubles = (size_t)P * 4 + (size_t)P * M1 * 2 + 2 * P + P / 2;
  auto UB = [&](int b) { return cov_smem + b * u_doubles; };                      // [MG][R2][LD]
  auto WGT = [&](int b) { return geo + b * geo_doubles; };                        // [P][4]
  auto TRG = [&](int b) { return WGT(b) + P * 4; };                               // [P][M1][2]
  auto SQM = [&](int b) { return TRG(b) + (size_t)P * M1 * 2; };                  // [P]
  auto PMS = [&](int b) { return SQM(b) + P; };                                   // [P]
  auto NBS = [&](int b) { return reinterpret_cast<int *>(PMS(b) + P); };          // [P]
  int *info = reinterpret_cast<int *>(geo + 2 * geo_doubles);                     // [4][2]: cnt, T of tile i & 3

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int gq = lane >> 2, t = lane & 3;
  const size_t nused = bstart[sampT];
  const size_t chunk = blockIdx.x;
  const size_t beg = chunk * (size_t)CH;
  if (beg >= nused) return;
  const size_t end = beg + CH < nused ? beg + CH : nused;

  if (w >= 8) {
    // ---- stage 2: panel u[m][n][j], u[m][N+n][j] of tile i-1 (zero beyond cnt); thread -> (split, m, n)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGP));
    const int pt = tid - 256;
    const int split = pt / items, it = pt % items;
    const bool evalr = split < nsplit;
    const int ml = it / N, nn = it % N;
    const int mn = (mg0 + ml) * N + nn;
    double vcr = 0.0, vci = 0.0;            // this thread's share of VC[T][m][n]
    double bmass = 0.0, bcount = 0.0;       // thread 0: massT / numbT of the current bin
    int curT = -1;
    auto flush = [&](int T) {
      const size_t slot = chunk + (size_t)T;
      if (evalr) {
        double *V = PV + ((slot * kCovSplit + split) * MN + mn) * 2;
        V[0] = vcr; V[1] = vci;
        vcr = vci = 0.0;
      }
      if (pt == 0 && mg0 == 0) {
        PM[slot * 2] = bmass; PM[slot * 2 + 1] = bcount;
        bmass = bcount = 0.0;
      }
    };
    for (int i = 0;; i++) {
      const int cnt = i >= 1 ? info[((i - 1) & 3) * 2] : 0;
      if (cnt > 0) {
        const int b = (i - 1) & 1, T0 = info[((i - 1) & 3) * 2 + 1];
        const double *wgt = WGT(b), *trg = TRG(b), *sqm = SQM(b), *pms = PMS(b);
        const int *nbs = NBS(b);
        if (T0 != curT) {
          if (curT >= 0) flush(curT);
          curT = T0;
        }
        if (pt == 0)
          for (int j = 0; j < cnt; j++) { bmass += pms[j]; bcount += 1.0; }
        const int kpad = (cnt + 3) & ~3;
        if (evalr) {
          double *ua = UB(b) + ((size_t)ml * R2 + nn) * LD, *ub = ua + (size_t)N * LD;
          for (int j = split; j < kpad; j += nsplit) {
            double a = 0.0, b2 = 0.0;
            if (j < cnt) {
              const double2 *z = Zt + (size_t)nbs[j] * MN + mn;
              const double2 z00 = __ldg(z), z01 = __ldg(z + MN);
              const double2 z10 = __ldg(z + (size_t)(g.numy + 1) * MN), z11 = __ldg(z + (size_t)(g.numy + 2) * MN);
              const double *wj = wgt + j * 4;
              const double vc = z00.x * wj[0] + z10.x * wj[1] + z01.x * wj[2] + z11.x * wj[3];
              const double vs = z00.y * wj[0] + z10.y * wj[1] + z01.y * wj[2] + z11.y * wj[3];
              const double cm = trg[(j * M1 + mg0 + ml) * 2], sm = trg[(j * M1 + mg0 + ml) * 2 + 1];
              if (g.cov_mode == 0) {
                a = vc * cm + vs * sm;
                b2 = vc * sm - vs * cm;
              } else if (g.cov_mode == 1) {   // sqrt(mass) * |(Vc cos, Vs sin)|: vc, vs carry -4 pi sqrt(mass)
                const double hc = vc * cm, hs = vs * sm;
                a = sqrt(hc * hc + hs * hs) * (0.25 / M_PI);
              } else {                        // sqrt(mass) * |norm| * |(Vc, Vs)|
                a = sqrt(vc * vc + vs * vs);
              }
              vcr += sqm[j] * a;
              vci += sqm[j] * b2;
            }
            ua[j] = a;
            ub[j] = b2;
          }
        }
      }
      __syncthreads();
      if (i >= 2 && info[((i - 2) & 3) * 2] == 0) break;
    }
    if (curT >= 0) flush(curT);
    return;
  }

  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGD));
  if (w == 7) {
    // ---- stage 1: the geometry records of the tile that starts at sorted position k0 -> shared
    // memory (cnt = 0 past the end of the chunk); tiles do not straddle a partition boundary
    size_t k0 = beg;
    int T = 0;
    size_t tend = bstart[1];
    for (int i = 0;; i++) {
      const int b = i & 1;
      int cnt = 0;
      if (k0 < end) {
        while (k0 >= tend) { T++; tend = bstart[T + 1]; }
        const size_t lim = tend < end ? tend : end;
        cnt = lim - k0 < (size_t)P ? (int)(lim - k0) : P;
        const double2 *gw = reinterpret_cast<const double2 *>(GW) + k0 * 2;
        const double2 *gt = reinterpret_cast<const double2 *>(GT) + k0 * M1;
        double2 *sw = reinterpret_cast<double2 *>(WGT(b)), *st = reinterpret_cast<double2 *>(TRG(b));
        const int nw = cnt * 2, nt = cnt * M1;
        double2 r0, r1, q[4];
        r0.x = r0.y = r1.x = r1.y = 0.0;
        if (lane < nw) r0 = __ldg(gw + lane);
        if (lane + 32 < nw) r1 = __ldg(gw + lane + 32);
        double sq = 0.0, ms = 0.0;
        int nb = 0;
        if (lane < cnt) { sq = __ldg(GS + k0 + lane); ms = __ldg(GM + k0 + lane); nb = __ldg(GN + k0 + lane); }
        for (int e0 = 0; e0 < nt; e0 += 128) {
#pragma unroll
          for (int u = 0; u < 4; u++)
            if (e0 + 32 * u + lane < nt) q[u] = __ldg(gt + e0 + 32 * u + lane);
#pragma unroll
          for (int u = 0; u < 4; u++)
            if (e0 + 32 * u + lane < nt) st[e0 + 32 * u + lane] = q[u];
        }
        if (lane < nw) sw[lane] = r0;
        if (lane + 32 < nw) sw[lane + 32] = r1;
        if (lane < cnt) { SQM(b)[lane] = sq; PMS(b)[lane] = ms; NBS(b)[lane] = nb; }
        k0 += cnt;
      }
      if (lane == 0) { info[(i & 3) * 2] = cnt; info[(i & 3) * 2 + 1] = T; }
      __syncthreads();
      if (i >= 2 && info[((i - 2) & 3) * 2] == 0) break;
    }
    return;
  }

  // ---- stage 3: rank-k update of tile i-2 on the tensor cores: warp w owns harmonic mg0 + w
  double acc[NT][NT][2];
#pragma unroll
  for (int a = 0; a < NT; a++)
#pragma unroll
    for (int b = 0; b < NT; b++) acc[a][b][0] = acc[a][b][1] = 0.0;
  int curT = -1;
  auto flush = [&](int T) {
    const size_t slot = chunk + (size_t)T;
    if (w < nml) {
      double *Pm = PC + ((slot * M1 + (mg0 + w)) * R2) * R2;
#pragma unroll
      for (int a = 0; a < NT; a++)
#pragma unroll
        for (int b = a; b < NT; b++) {         // U U^T is symmetric: tiles on and above the diagonal
          const int row = 8 * a + gq, col = 8 * b + 2 * t;
          if (row < R2 && col < R2) {          // R2 is even: col + 1 < R2 as well
            Pm[(size_t)row * R2 + col] = acc[a][b][0];
            Pm[(size_t)row * R2 + col + 1] = acc[a][b][1];
          }
          acc[a][b][0] = acc[a][b][1] = 0.0;
        }
    }
  };
  for (int i = 0;; i++) {
    const int cnt = i >= 2 ? info[((i - 2) & 3) * 2] : 0;
    if (cnt > 0) {
      const int T0 = info[((i - 2) & 3) * 2 + 1];
      if (T0 != curT) {
        if (curT >= 0) flush(curT);
        curT = T0;
      }
      if (w < nml) {
        const int kpad = (cnt + 3) & ~3;
        const double *Um = UB(i & 1) + (size_t)w * R2 * LD;
        for (int kk = 0; kk < kpad; kk += 4) {
          double af[NT];
#pragma unroll
          for (int r = 0; r < NT; r++) {
            const int row = 8 * r + gq;
            af[r] = row < R2 ? Um[(size_t)row * LD + kk + t] : 0.0;
          }
#pragma unroll
          for (int a = 0; a < NT; a++)
#pragma unroll
            for (int b3 = a; b3 < NT; b3++) dmma_m8n8k4(acc[a][b3][0], acc[a][b3][1], af[a], af[b3]);
        }
      }
    }
    __syncthreads();
    if (i >= 2 && info[((i - 2) & 3) * 2] == 0) break;
  }
  if (curT >= 0) flush(curT);
}

// MV / VC / massT / numbT of bin T = sum over the chunks that overlap it, in chunk order
__global__ void __launch_bounds__(128)
cyl_cov_combine_kernel(CylGeom g, const uint32_t *__restrict__ bstart, int CH, int nsplit,
                       const double *__restrict__ PC, const double *__restrict__ PV,
                       const double *__restrict__ PM, double *__restrict__ outMV,
                       double *__restrict__ outVC, double *__restrict__ outMass,
                       unsigned *__restrict__ outCount)
{
  const int T = blockIdx.x, m = blockIdx.y;
  const int N = g.norder, R2 = 2 * N, M1 = g.mmax + 1, MN = M1 * N;
  const uint32_t b0 = bstart[T], b1 = bstart[T + 1];
  const bool any = b1 > b0;
  const size_t c0 = any ? b0 / CH : 1, c1 = any ? (b1 - 1) / CH : 0;
  double *MV = outMV + ((size_t)T * M1 + m) * N * N * 2;
  for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
    const int n = e / N, o = e % N;
    double re = 0.0, im = 0.0;
    for (size_t c = c0; c <= c1; c++) {
      const double *P = PC + (((c + T) * M1 + m) * R2) * R2;
      // only the 8x8 tiles on and above the diagonal were stored; the rest by symmetry
      auto at = [&](int r, int q) { return (r >> 3) <= (q >> 3) ? P[(size_t)r * R2 + q] : P[(size_t)q * R2 + r]; };
      re += at(n, o) + at(N + n, N + o);
      im += at(N + n, o) - at(n, N + o);
    }
    MV[2 * e] = re;
    MV[2 * e + 1] = im;
  }
  double *VC = outVC + ((size_t)T * M1 + m) * N * 2;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    double re = 0.0, im = 0.0;
    for (size_t c = c0; c <= c1; c++)
      for (int sp = 0; sp < nsplit; sp++) {
        const double *V = PV + (((c + T) * kCovSplit + sp) * MN + (size_t)m * N + n) * 2;
        re += V[0];
        im += V[1];
      }
    VC[2 * n] = re;
    VC[2 * n + 1] = im;
  }
  if (m == 0 && threadIdx.x == 0) {
    double ms = 0.0, ct = 0.0;
    for (size_t c = c0; c <= c1; c++) { ms += PM[(c + T) * 2]; ct += PM[(c + T) * 2 + 1]; }
    outMass[T] = ms;
    outCount[T] = (unsigned)(ct + 0.5);
  }
}

template <int NT>
int launch_cyl_cov(cudaStream_t st, dim3 grid, size_t smem, const CylGeom &g, const double2 *Zt,
                   const double *GW, const double *GT, const double *GS, const double *GM, const int *GN,
                   const uint32_t *bstart, int sampT, int CH, int nsplit, int MG, double *PC, double *PV,
                   double *PM)
{
  // registers: 768 threads start with 80 each; the two DMMA warpgroups (accumulators: 2 NT^2
  // doubles) grow to REGD out of what the four panel warpgroups release by shrinking to REGP
  // (setmaxnreg.inc waits for the CTA's own pool: 256 REGD + 512 REGP <= 768 x 80)
  constexpr int REGD = NT <= 5 ? 128 : NT <= 6 ? 144 : 160;
  constexpr int REGP = NT <= 5 ? 56 : NT <= 6 ? 48 : 40;
  static_assert(256 * REGD + 512 * REGP <= 768 * 80, "register budget");
  auto kern = cyl_cov_kernel<NT, REGD, REGP>;
  CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, kCovThreads, smem, st>>>(g, Zt, GW, GT, GS, GM, GN, bstart, sampT, CH, nsplit, MG, PC, PV, PM);
  CU_LAUNCH_CHECK();
  return EXPB200_OK;
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
  const expb200_component_s *owner = nullptr;   // PotAccel::component (expb200_cylinder_set_owner)
  uint64_t orient_ver = 1;          // bumped by set_orientation: the cells depend on it
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
  double2 *Zt = nullptr;            // covariance: node-major complex pot table [nnode][(mmax+1)*norder]
  void *cov_buf[13] = {};      // cached scratch of the covariance call
  size_t cov_cap[13] = {};
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
  // origin: the centre of the component that owns this force (component->ConvertPos,
  // src/Cylinder.cc:1343-1349); the cylinder's force pass has no freeze test, and the
  // coefficient pass always runs on the owner itself
  const expb200_component_s *o = s->owner ? s->owner : c;
  g.cx = o->center[0]; g.cy = o->center[1]; g.cz = o->center[2];
  g.rtrunc2 = c->rtrunc * c->rtrunc;
  return g;
}

int ensure_sorted(expb200_cylinder_s *s, expb200_component_s *c, int mlevel, double adb, bool want_mass)
{
  const uint32_t kpl = (uint32_t)s->geom.ncell + 2u;
  const uint32_t nkeys = (uint32_t)s->nlev * kpl;
  CylGeom g = geom_for(s, c);
  if (s->nlev == 1) mlevel = 0;
  const bool valid = c->sort_kind == 3 && c->sort_owner == s && c->sort_pos_ver == c->pos_ver &&
                     c->sort_lev_ver == c->lev_ver && c->sort_aux_ver == s->orient_ver &&
                     c->sort_origin[0] == g.cx && c->sort_origin[1] == g.cy && c->sort_origin[2] == g.cz &&
                     c->sort_from <= mlevel;
  if (valid && !want_mass) return EXPB200_OK;
  if (c->seg_cap < (size_t)nkeys + 2) {
    if (c->seg) cudaFree(c->seg);
    c->seg = nullptr;
    CU_TRY(cudaMalloc(&c->seg, ((size_t)nkeys + 2) * sizeof(uint32_t)));
    c->seg_cap = (size_t)nkeys + 2;
  }
  // multistep: re-bin only the tail of the levels that can have moved (see sphere.cu: ensure_sorted)
  int m = 0;
  if (!valid && s->nlev > 1 && c->n && c->part_ok && c->part_nlev == s->nlev)
    m = mlevel < c->bound_ok_upto ? mlevel : c->bound_ok_upto;
  if (m > 0) {
    const size_t off = c->lev_start[m], nsub = c->n - off;
    const uint32_t nks = (uint32_t)(s->nlev - m) * kpl;
    uint32_t *seg_sub = c->seg + (size_t)m * kpl;
    const unsigned nbs = grid_for(nsub ? nsub : 1, 256);
    const int W = suffix_key_buffer(c->cur, nks);
    {
      ProfRegion pr(K_KEY, c->stream, 2);
      cyl_key_sub_kernel<<<nbs, 256, 0, c->stream>>>(g, c->posm, c->level, c->perm[c->cur] + off, nsub, m,
                                                     s->nlev - 1, mlevel, c->keys[W] + off, s->block_mass);
      cyl_mass_kernel<<<1, 32, 0, c->stream>>>(s->block_mass, nbs, adb, s->coef + s->ncoef);
      CU_LAUNCH_CHECK();
    }
    int rc;
    if (nsub) {
      ProfRegion pr(K_SORT, c->stream, radix_sort_launches(nsub, nks, false) + 2);
      rc = radix_sort_suffix(c->stream, c->keys, c->perm, c->cur, c->hist, c->rank, off, c->n, nks, seg_sub,
                             SortPayload(), (uint32_t)m * kpl);
    } else {
      rc = segment_offsets(c->stream, c->keys[c->cur], 0, seg_sub, nks, (uint32_t)off);
    }
    if (rc) return rc;
    if ((rc = component_read_bounds(c, kpl, s->nlev, m))) return rc;
    c->sort_kind = 3;
    c->sort_owner = s;
    c->keys_per_level = kpl;
    c->sort_pos_ver = c->pos_ver;
    c->sort_lev_ver = c->lev_ver;
    c->sort_aux_ver = s->orient_ver;
    c->sort_origin[0] = g.cx; c->sort_origin[1] = g.cy; c->sort_origin[2] = g.cz;
    c->sort_from = m;
    c->sort_counters[1]++; c->sort_counters[2] += nsub;
    c->sorts_since_reorder++;
    return EXPB200_OK;
  }
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
  c->sort_aux_ver = s->orient_ver;
  c->sort_origin[0] = g.cx; c->sort_origin[1] = g.cy; c->sort_origin[2] = g.cz;
  c->sort_from = 0;
  c->sort_counters[0]++; c->sort_counters[3] += c->n;
  if (s->nlev > 1 && c->n) return component_read_bounds(c, kpl, s->nlev, -1);
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
  g.cx = g.cy = g.cz = 0.0; g.rtrunc2 = 1.0e300; g.rotate = 0; g.cov_mode = 0;
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
  cudaFree(s->G); cudaFree(s->F); cudaFree(s->PA); cudaFree(s->block_mass); cudaFree(s->Zt);
  for (int q = 0; q < 13; q++) cudaFree(s->cov_buf[q]);
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
  s->orient_ver++;          // cached (level, cell) sorts were made in the old frame
  return EXPB200_OK;
}

int expb200_cylinder_set_owner(expb200_cylinder s, expb200_component owner)
{
  REQUIRE(s, "cylinder is NULL");
  s->owner = owner;
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
    // compact the switching particles first (see expb200_sphere_multistep_update)
    size_t nsw = 0;
    uint32_t *list = c->rank + c->n / 2;
    { int rc = compact_differing(st, c->level, c->prev_level, c->n, c->hist, c->keys[1], &nsw); if (rc) return rc; }
    const bool compact = nsw > 0 && 2 * nsw <= c->n;
    if (compact)
      CU_TRY(cudaMemcpyAsync(list, c->keys[1], nsw * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    const size_t nn = compact ? nsw : c->n;
    const int CH = cyl_chunk(nn);
    const size_t slots = (nn + CH - 1) / CH + (size_t)g.ncell + 1;
    if (slots > s->part_slots) {
      if (s->PA) cudaFree(s->PA);
      s->PA = nullptr; s->part_slots = 0;
      CU_TRY(cudaMalloc(&s->PA, slots * 4 * g.nk * sizeof(double)));
      s->part_slots = slots;
    }
    const int stride = (g.nk + 4) | 1;
    const size_t sm = (size_t)kCylDepositWarps * (32 * stride + 16) * sizeof(double);
    const unsigned nchunks = div_up(nn, CH);
    if (compact && c->sw_cap < nsw) {
      if (c->sw_perm) cudaFree(c->sw_perm);
      c->sw_perm = nullptr; c->sw_cap = 0;
      CU_TRY(cudaMalloc(&c->sw_perm, nsw * sizeof(uint32_t)));
      c->sw_cap = nsw;
    }
    // the compact path leaves perm[cur] (the level partition of the component's own sort) alone
    uint32_t *kl[2] = {c->keys[0], c->keys[1]};
    uint32_t *pl2[2] = {c->sw_perm, c->perm[c->cur ^ 1]};
    for (int side = 0; side < 2 && nsw > 0; side++) {          // 0: "to" (+), 1: "from" (-)
      const uint32_t *sperm = nullptr;
      if (compact) {
        cyl_switch_key_list_kernel<<<grid_for(nsw, 256), 256, 0, st>>>(g, c->posm, c->level, c->prev_level, list,
                                                                       nsw, side, s->nlev, kl[0]);
        CU_LAUNCH_CHECK();
        int cur = 0;
        int rc = radix_sort_pairs(st, kl, pl2, c->hist, c->rank, nsw, nkeys, &cur, c->seg, SortPayload(), list);
        if (rc) return rc;
        sperm = pl2[cur];
      } else {
        cyl_switch_key_kernel<<<grid_for(c->n, 256), 256, 0, st>>>(g, c->posm, c->level, c->prev_level, c->n,
                                                                   side, s->nlev, c->keys[0]);
        CU_LAUNCH_CHECK();
        int cur = 0;
        int rc = radix_sort_pairs(st, c->keys, c->perm, c->hist, c->rank, c->n, nkeys, &cur, c->seg);
        if (rc) return rc;
        c->cur = cur;
        sperm = c->perm[cur];
      }
      for (int M = 0; M < s->nlev; M++) {
        const int keybase = M * (int)kpl;
        cyl_deposit_kernel<<<div_up(nchunks, kCylDepositWarps), kCylDepositWarps * 32, sm, st>>>(
            g, c->posm, sperm, c->seg, keybase, CH, adb, stride, 1, s->PA);
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
    if (nsw > 0) {
      // keys and cell offsets were overwritten; the level partition survives the compact path
      c->sort_pos_ver = 0;
      if (!compact) { c->sort_kind = 0; c->part_ok = false; }
    }
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

// ~ the covariance branch of EmpCylSL::accumulate for levlist[mlevel] (see the kernels above)
static int cylinder_covariance_impl(expb200_cylinder s, expb200_component c, int mlevel, double adb,
                                    int sampT, uint64_t indx_base, unsigned *numbT, double *massT,
                                    double *VC, double *MV, int mode)
{
  REQUIRE(s && c, "NULL argument");
  REQUIRE(s->device == c->device, "device mismatch");
  REQUIRE(mlevel >= 0 && mlevel < s->nlev, "mlevel=%d outside [0,%d]", mlevel, s->nlev - 1);
  const CylGeom &g0 = s->geom;
  REQUIRE(sampT >= 1 && (uint64_t)sampT * g0.ncell < 0x7fffffffull, "sampT=%d out of range", sampT);
  REQUIRE(g0.norder <= 32, "covariance supports norder <= 32");
  DeviceGuard gd(s->device);
  cudaStream_t st = c->stream;
  { cudaStream_t ws = s->work_stream(); if (ws != st) CU_TRY(cudaStreamSynchronize(ws)); }
  s->last_stream = st;
  const int N = g0.norder, R2 = 2 * N, M1 = g0.mmax + 1, MN = M1 * N;
  if (numbT) memset(numbT, 0, (size_t)sampT * sizeof(unsigned));
  if (massT) memset(massT, 0, (size_t)sampT * sizeof(double));
  if (VC) memset(VC, 0, (size_t)sampT * MN * 2 * sizeof(double));
  if (MV) memset(MV, 0, (size_t)sampT * MN * N * 2 * sizeof(double));
  if (c->n == 0) return EXPB200_OK;

  CylGeom g = geom_for(s, c);
  g.cov_mode = mode;
  g.rotate = s->geom.rotate;
  ProfRegion pr(K_COV, st, 0);
  if (!s->Zt) {
    CU_TRY(cudaMalloc(&s->Zt, (size_t)g.nnode * MN * sizeof(double2)));
    cyl_cov_table_kernel<<<grid_for((size_t)g.nnode * MN, 256), 256, 0, st>>>(g, s->tab[0][0], s->tab[0][1],
                                                                              s->cfg.MLIM, s->cfg.EVEN_M, s->Zt);
    CU_LAUNCH_CHECK();
  }
  // (T, cell) keys -> stable sort; the component's own sort is replaced
  const uint32_t nkeys = (uint32_t)sampT * (uint32_t)g.ncell + 1;
  cyl_cov_key_kernel<<<grid_for(c->n, 256), 256, 0, st>>>(g, c->posm, c->level,
                                                          c->identity_order ? nullptr : c->slot_of, c->n,
                                                          s->nlev > 1, mlevel, (uint32_t)sampT, indx_base,
                                                          c->keys[0]);
  CU_LAUNCH_CHECK();
  int cur = 0;
  int rc = radix_sort_pairs(st, c->keys, c->perm, c->hist, c->rank, c->n, nkeys, &cur);
  if (rc) return rc;
  c->cur = cur;
  c->sort_kind = 0; c->part_ok = false;

  uint32_t *bstart = nullptr;
  double *PC = nullptr, *PV = nullptr, *PM = nullptr, *dMV = nullptr, *dVC = nullptr, *dMass = nullptr;
  unsigned *dCount = nullptr;
  // scratch is cached in the cylinder (grow-only, freed with it): the cudaMalloc/cudaFree pairs
  // of a few hundred MB per call cost several times what the kernels do (measured 150 ms
  // against 20 ms at 1e7 particles, sampT = 100)
  auto cleanup = [&]() {};
  auto cov_alloc = [&](void **ptr, size_t bytes, int slot) -> cudaError_t {
    if (s->cov_cap[slot] < bytes) {
      if (s->cov_buf[slot]) cudaFree(s->cov_buf[slot]);
      s->cov_buf[slot] = nullptr; s->cov_cap[slot] = 0;
      const size_t want = bytes + bytes / 8;        // head-room: sizes drift a little from call to call
      cudaError_t e = cudaMalloc(&s->cov_buf[slot], want);
      if (e != cudaSuccess) return e;
      s->cov_cap[slot] = want;
    }
    *ptr = s->cov_buf[slot];
    return cudaSuccess;
  };
#define COV_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { cleanup(); return cuda_fail(_e, #expr, __FILE__, __LINE__); } } while (0)
  COV_TRY(cov_alloc((void **)&bstart, ((size_t)sampT + 1) * sizeof(uint32_t), 0));
  cyl_cov_bounds_kernel<<<div_up((size_t)sampT + 1, 128), 128, 0, st>>>(c->keys[c->cur], c->n, (uint32_t)g.ncell,
                                                                        sampT, bstart);
  uint32_t nused = 0;
  COV_TRY(cudaMemcpyAsync(&nused, bstart + sampT, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  COV_TRY(cudaStreamSynchronize(st));
  if (nused == 0) { cleanup(); return EXPB200_OK; }

  // chunk: a few waves of CTAs over the 148 SMs, 256..4096 particles each
  size_t chq = nused / ((size_t)kSMsB200 * 8);
  int CH = (int)(chq < 256 ? 256 : chq > 4096 ? 4096 : chq);
  const unsigned nchunks = div_up(nused, CH);
  const size_t slots = (size_t)nchunks + sampT + 1;
  // scratch capacity from the component size, so that it does not move with nused
  const size_t slots_cap = (size_t)div_up(c->n, CH) + sampT + 1;
  // harmonics per CTA: one DMMA warp each, and one panel thread per (m, n) of the group
  int MG = kCovPanel / N;
  if (MG > kCovMG) MG = kCovMG;
  if (MG > M1) MG = M1;
  const int mgroups = (M1 + MG - 1) / MG;
  const int items = MG * N;
  int nsplit = kCovPanel / items;
  if (nsplit > kCovSplit) nsplit = kCovSplit;
  COV_TRY(cov_alloc((void **)&PC, slots_cap * M1 * R2 * R2 * sizeof(double), 1));
  COV_TRY(cov_alloc((void **)&PV, slots_cap * kCovSplit * MN * 2 * sizeof(double), 2));
  COV_TRY(cov_alloc((void **)&PM, slots_cap * 2 * sizeof(double), 3));
  COV_TRY(cudaMemsetAsync(PV, 0, slots * kCovSplit * MN * 2 * sizeof(double), st));
  COV_TRY(cov_alloc((void **)&dMV, (size_t)sampT * MN * N * 2 * sizeof(double), 4));
  COV_TRY(cov_alloc((void **)&dVC, (size_t)sampT * MN * 2 * sizeof(double), 5));
  COV_TRY(cov_alloc((void **)&dMass, (size_t)sampT * sizeof(double), 6));
  COV_TRY(cov_alloc((void **)&dCount, (size_t)sampT * sizeof(unsigned), 7));

  // geometry records of the used particles, in sorted order (164 B per particle at mmax = 6)
  double *GW = nullptr, *GT = nullptr, *GS = nullptr, *GM = nullptr;
  int *GN = nullptr;
  // sized for the whole component (nused <= n changes from call to call as particles cross the
  // grid's edge: buffers sized by nused were re-allocated, 1.6 GB at a time, on most steps)
  COV_TRY(cov_alloc((void **)&GW, (size_t)c->n * 4 * sizeof(double), 8));
  COV_TRY(cov_alloc((void **)&GT, (size_t)c->n * M1 * 2 * sizeof(double), 9));
  COV_TRY(cov_alloc((void **)&GS, (size_t)c->n * sizeof(double), 10));
  COV_TRY(cov_alloc((void **)&GM, (size_t)c->n * sizeof(double), 11));
  COV_TRY(cov_alloc((void **)&GN, (size_t)c->n * sizeof(int), 12));
  cyl_cov_geom_kernel<<<div_up(nused, 256), 256, 0, st>>>(g, c->posm, c->perm[c->cur], nused, adb, GW, GT, GS, GM, GN);
  COV_TRY(cudaGetLastError());
  constexpr int P = kCovTile;
  const size_t smem = (2 * (size_t)MG * R2 * (P + kCovLDpad) +
                       2 * ((size_t)P * 4 + (size_t)P * M1 * 2 + 2 * P + P / 2)) * sizeof(double) + 8 * sizeof(int);
  const dim3 grid(nchunks, mgroups);
  const int NT = (R2 + 7) / 8;
#define COV_LAUNCH(K) rc = launch_cyl_cov<K>(st, grid, smem, g, s->Zt, GW, GT, GS, GM, GN, bstart, sampT, CH, nsplit, \
                                             MG, PC, PV, PM)
  switch (NT) {
    case 1: COV_LAUNCH(1); break;  case 2: COV_LAUNCH(2); break;  case 3: COV_LAUNCH(3); break;
    case 4: COV_LAUNCH(4); break;  case 5: COV_LAUNCH(5); break;  case 6: COV_LAUNCH(6); break;
    default: COV_LAUNCH(8); break;
  }
#undef COV_LAUNCH
  if (rc) { cleanup(); return rc; }
  cyl_cov_combine_kernel<<<dim3(sampT, M1), 128, 0, st>>>(g, bstart, CH, nsplit, PC, PV, PM, dMV, dVC, dMass, dCount);
  COV_TRY(cudaGetLastError());
  if (MV) COV_TRY(cudaMemcpyAsync(MV, dMV, (size_t)sampT * MN * N * 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (VC) COV_TRY(cudaMemcpyAsync(VC, dVC, (size_t)sampT * MN * 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (massT) COV_TRY(cudaMemcpyAsync(massT, dMass, (size_t)sampT * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (numbT) COV_TRY(cudaMemcpyAsync(numbT, dCount, (size_t)sampT * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
  COV_TRY(cudaStreamSynchronize(st));
#undef COV_TRY
  cleanup();
  return EXPB200_OK;
}

int expb200_cylinder_covariance(expb200_cylinder s, expb200_component c, int mlevel, double adb,
                                int sampT, uint64_t indx_base, unsigned *numbT, double *massT,
                                double *VC, double *MV)
{
  return cylinder_covariance_impl(s, c, mlevel, adb, sampT, indx_base, numbT, massT, VC, MV, 0);
}

// ~ the PCAVAR / PCAEOF branches of EmpCylSL::accumulate (exputil/EmpCylSL.cc:4089-4119):
//   which = 1 (PCAVAR): covV[T][m][n] += mass modu1,  covM[T][m](n,o) += mass modu1 modu2,
//                       modu = |(Vc cos m phi, Vs sin m phi)|
//   which = 2 (PCAEOF): tvar[m](n,o) += mass modu1 modu2, modu = |norm| |(Vc, Vs)| (pass sampT = 1
//                       and read covM; covV is then sum mass modu)
// Real outputs covV [sampT][(mmax+1)][norder], covM [sampT][(mmax+1)][norder][norder].
int expb200_cylinder_covariance_pca(expb200_cylinder s, expb200_component c, int mlevel, double adb,
                                    int sampT, uint64_t indx_base, int which, unsigned *numbT,
                                    double *massT, double *covV, double *covM)
{
  REQUIRE(s && (which == 1 || which == 2), "which=%d (1: PCAVAR, 2: PCAEOF)", which);
  const size_t MN = (size_t)(s->cfg.mmax + 1) * s->cfg.norder, N = s->cfg.norder;
  std::vector<double> VC(covV ? (size_t)sampT * MN * 2 : 0), MV(covM ? (size_t)sampT * MN * N * 2 : 0);
  int rc = cylinder_covariance_impl(s, c, mlevel, adb, sampT, indx_base, numbT, massT,
                                    covV ? VC.data() : nullptr, covM ? MV.data() : nullptr, which);
  if (rc) return rc;
  if (covV) for (size_t k = 0; k < (size_t)sampT * MN; k++) covV[k] = VC[2 * k];
  if (covM) for (size_t k = 0; k < (size_t)sampT * MN * N; k++) covM[k] = MV[2 * k];
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
