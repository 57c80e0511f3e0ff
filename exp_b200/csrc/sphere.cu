// sphere.cu -- the sphereSL force (SphericalBasis + SLGridSph) on B200.
//
// Replaces SphericalBasis::determine_coefficients_cuda / determine_acceleration_cuda /
// DtoH_coefs / HtoD_coefs and SLGridSph::initialize_cuda (src/SphericalBasis.H:186-270,
// src/cudaSphericalBasis.cu, exputil/cudaSLGridMP2.cu).  Not a port of those kernels:
// see sph_math.cuh for the factorised evaluation this file implements.
//
// Per coefficient call (levlist[mlevel]):
//   key      : radial cell (and level) of every particle            [HBM stream]
//   sort     : one-pass stable radix binning -> permutation, cell offsets and a copy
//              of the position records in sorted order              [HBM stream]
//   deposit  : per particle P_lm(theta) {cos,sin}(m phi) * node weights; lane = particle
//              with the (lmax+1)^2 x 2 sums of a radial cell in registers, reduced across
//              the warp by a fixed butterfly when the cell changes   [FP64 pipe]
//   combine  : per-cell partials -> node sums G[row][i]              [tiny]
//   contract : coef[row][n] = K_row sum_i E[l][n][i] G[row][i]       [tiny]
// Per force call:
//   field    : T[i][slot] = per-cell polynomial coefficients of sum_n E coef   [tiny]
//   force    : per particle Horner forms on T, Legendre + derivative, acc/pot  [FP64 pipe]
// The single-level step additionally fuses kick+drift+key and force+kick.
#include "common.cuh"
#include "sort.cuh"
#include "sph_math.cuh"
#include "sph_host.h"
#include "nccl_shim.h"
#include "profile.cuh"

#include <cmath>
#include <cstdlib>
#include <vector>

namespace expb200 {

__constant__ double c_K[(kMaxL + 1) * (kMaxL + 1)];   // factorial(l,m), src/SphericalBasis.cc:330-335
__constant__ double c_inv[2 * kMaxL + 4];             // 1/k

namespace {

constexpr int kDepositWarps = 4;
constexpr int kForceThreads = 128;
constexpr int kRegDepositMaxL = 6;     // (L+1)^2 x 2 register accumulators fit up to here

template <int L> struct DepositShape {
  static constexpr int NROW = (L + 1) * (L + 1);
  static constexpr int STRIDE = (NROW + 2) | 1;         // odd: conflict-free 64-bit lanes
  static constexpr int RPL = (NROW + 31) / 32;          // rows per lane in the reduce phase
  static constexpr size_t WARP_DOUBLES = 32 * STRIDE + 16;   // + 32 ints of cell ids
};

// ---------------------------------------------------------------------------
// key: (level, radial cell) per particle
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sph_key_kernel(SphGeom g, const rec4 *__restrict__ posm, const unsigned *__restrict__ level,
               size_t n, int use_levels, int maxlev, uint32_t *__restrict__ keys)
{
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    rec4 pm = ld_rec(&posm[i]);
    uint32_t cell = sph_cell_of(g, pm);
    uint32_t lev = 0;
    if (use_levels) {
      lev = level[i];
      if (lev > (uint32_t)maxlev) lev = maxlev;
    }
    keys[i] = lev * (uint32_t)g.numr + cell;
  }
}

// keys of the tail of the sorted permutation (levels >= m), relative to level m: see ensure_sorted
__global__ void __launch_bounds__(256)
sph_key_sub_kernel(SphGeom g, const rec4 *__restrict__ posm, const unsigned *__restrict__ level,
                   const uint32_t *__restrict__ slots, size_t nsub, int m, int maxlev,
                   uint32_t *__restrict__ keys)
{
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < nsub;
       k += (size_t)gridDim.x * blockDim.x) {
    const uint32_t i = slots[k];
    uint32_t lev = level[i];
    if (lev > (uint32_t)maxlev) lev = maxlev;
    if (lev < (uint32_t)m) lev = m;            // cannot happen while the partition holds
    keys[k] = (lev - (uint32_t)m) * (uint32_t)g.numr + sph_cell_of(g, ld_rec(&posm[i]));
  }
}

// Single-level step: first half kick, drift and the radial key of the new position in one
// pass over the particle records (src/step.cc:272-290; arithmetic of src/incvel.cc:68-69,
// src/incpos.cc:63-66).
__global__ void __launch_bounds__(256)
sph_kick_drift_key_kernel(SphGeom g, rec4 *__restrict__ posm, rec4 *__restrict__ vel,
                          const rec4 *__restrict__ accp, size_t n, double dtk, double dtd,
                          uint32_t *__restrict__ keys)
{
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    rec4 v = ld_rec_rw(&vel[i]);
    rec4 a = ld_rec(&accp[i]);
    rec4 p = ld_rec_rw(&posm[i]);
    v.x = fma(a.x, dtk, v.x); v.y = fma(a.y, dtk, v.y); v.z = fma(a.z, dtk, v.z);
    p.x = fma(v.x, dtd, p.x); p.y = fma(v.y, dtd, p.y); p.z = fma(v.z, dtd, p.z);
    st_rec(&vel[i], v);
    st_rec(&posm[i], p);
    keys[i] = sph_cell_of(g, p);
  }
}

// keys for multistep_update: particles whose level changed get (side level, cell with the
// update's own range test), everything else the sentinel key nlev*numr.
__global__ void __launch_bounds__(256)
sph_switch_key_kernel(SphGeom g, const rec4 *__restrict__ posm, const unsigned *__restrict__ level,
                      const unsigned *__restrict__ prev, size_t n, int use_prev, int nlev,
                      uint32_t *__restrict__ keys)
{
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    unsigned a = prev[i], b = level[i];
    uint32_t key = (uint32_t)nlev * (uint32_t)g.numr;
    if (a != b) {
      uint32_t lev = use_prev ? a : b;
      if (lev >= (uint32_t)nlev) lev = nlev - 1;
      key = lev * (uint32_t)g.numr + sph_cell_of(g, ld_rec(&posm[i]), 1);
    }
    keys[i] = key;
  }
}

// The same keys for a compacted list of switching particles (list[k] = store index): no sentinel.
__global__ void __launch_bounds__(256)
sph_switch_key_list_kernel(SphGeom g, const rec4 *__restrict__ posm, const unsigned *__restrict__ level,
                           const unsigned *__restrict__ prev, const uint32_t *__restrict__ list, size_t nsw,
                           int use_prev, int nlev, uint32_t *__restrict__ keys)
{
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < nsw;
       k += (size_t)gridDim.x * blockDim.x) {
    const uint32_t i = list[k];
    uint32_t lev = use_prev ? prev[i] : level[i];
    if (lev >= (uint32_t)nlev) lev = nlev - 1;
    keys[k] = lev * (uint32_t)g.numr + sph_cell_of(g, ld_rec(&posm[i]), 1);
  }
}

// dst[k] += sign * src[k]
__global__ void axpy_kernel(double *__restrict__ dst, const double *__restrict__ src, double sign, size_t n)
{
  size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (k < n) dst[k] += sign * src[k];
}

// compute_multistep_coefficients (src/SphericalBasis.cc:1252-1333):
// out = sum_{M<mfirst} (a_M L[M] + b_M N[M]) + sum_{M>=mfirst} N[M], levels in order.
struct InterpArgs { int nlev, mfirst; double a[32], b[32]; };
__global__ void multistep_interp_kernel(const double *__restrict__ cL, const double *__restrict__ cN,
                                        InterpArgs A, size_t ncoef, double *__restrict__ out)
{
  size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (k >= ncoef) return;
  double s = 0.0;
  for (int M = 0; M < A.mfirst; M++) s += A.a[M] * cL[M * ncoef + k] + A.b[M] * cN[M * ncoef + k];
  for (int M = A.mfirst; M < A.nlev; M++) s += cN[M * ncoef + k];
  out[k] = s;
}

// ---------------------------------------------------------------------------
// deposit (lmax <= 6): one warp per chunk of CH sorted particles, lane = particle.
// Every lane keeps the 2 x NACC sums of the current radial cell in registers; when the
// cell changes (at most numr times over the whole array) or the chunk ends they are
// reduced across the warp and written to partial slot (chunk + cell), which is unique
// because both chunk and cell are non-decreasing along the sorted order.
// One launch ("sweep") covers the azimuthal columns M0..M1 so that the accumulators fit
// the register file without spilling: lmax = 6 takes two sweeps (m <= 2: 58 sums,
// m >= 3: 40 sums).
// ---------------------------------------------------------------------------
template <int L, int M0, int M1>
__device__ __forceinline__ int deposit_row_of(int idx)
{
  int row = -1;
#pragma unroll
  for (int m = M0; m <= M1; m++)
#pragma unroll
    for (int l = m; l <= L; l++) {
      const int c = sph_compact<L, M0>(l, m, 0);
      if (idx == c) row = sph_row(l, m);
      if (m > 0 && idx == c + 1) row = sph_row(l, m) + 1;
    }
  return row;
}

// GATHER: the sorted position stream does not exist yet -- this sweep reads the records of the
// store through the sort permutation (two batches ahead) and writes the stream (ps_out) for the
// sweeps and the force pass that follow; the binning pass then moves keys and slots only
// (expb200_sphere_steps).
struct DepositGather { const uint32_t *perm = nullptr; const rec4 *store = nullptr; rec4 *out = nullptr; };

template <int L, int M0, int M1, int MINB, bool GATHER = false>
__global__ void __launch_bounds__(kDepositWarps * 32, MINB)
sph_deposit_reg_kernel(SphGeom g, const double *__restrict__ p0, const rec4 *__restrict__ ps,
                       const uint32_t *__restrict__ skeys, const uint32_t *__restrict__ seg,
                       int keybase0, int CH, double adb, int batched, double *__restrict__ PA,
                       double *__restrict__ PB, DepositGather G = DepositGather())
{
  constexpr int NROW = (L + 1) * (L + 1);
  constexpr int NACC = sph_compact_count<L, M0, M1>();
  using R = RS<NACC, 0>;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  // batched: blockIdx.y selects one of several consecutive key groups (covariance bins)
  const int keybase = keybase0 + blockIdx.y * g.numr;
  const uint32_t lo = seg[keybase], hi = seg[keybase + g.numr - 1];
  const size_t sbase = batched ? (size_t)blockIdx.y * (g.numr + 2) + (lo - seg[keybase0]) / CH : 0;
  const size_t chunk = (size_t)blockIdx.x * kDepositWarps + w;
  const size_t beg = (size_t)lo + chunk * (size_t)CH;
  if (beg >= hi) return;
  const size_t end = beg + CH < hi ? beg + CH : hi;

  double A[NACC], B[NACC];
#pragma unroll
  for (int r = 0; r < NACC; r++) A[r] = B[r] = 0.0;
  int cur = -1;

  auto flush = [&](int c) {
    const size_t slot = (sbase + chunk + (size_t)c) * NROW;
    R::run(A, lane);
    R::run(B, lane);
#pragma unroll
    for (int i = 0; i < R::FIN; i++) {
      const int idx = R::row(lane, i);
      const int row = idx >= 0 ? deposit_row_of<L, M0, M1>(idx) : -1;
      if (row >= 0) {
        PA[slot + row] = A[i];
        PB[slot + row] = B[i];
      }
    }
#pragma unroll
    for (int r = 0; r < NACC; r++) A[r] = B[r] = 0.0;
  };

  rec4 nxt = rec4{1.0, 0.0, 0.0, 0.0};
  uint32_t nkey = 0, slot1 = 0;
  if (GATHER) {
    if (beg + lane < end) { nxt = ld_rec(&G.store[G.perm[beg + lane]]); nkey = skeys[beg + lane]; }
    if (beg + 32 + lane < end) slot1 = G.perm[beg + 32 + lane];
  } else {
    if (beg + lane < end) { nxt = ld_rec(&ps[beg + lane]); nkey = skeys[beg + lane]; }
  }

  for (size_t k0 = beg; k0 < end; k0 += 32) {
    const bool valid = k0 + lane < end;
    const rec4 pm = nxt;
    int cell = valid ? (int)nkey - keybase : 0x7fffffff;
    {
      const size_t kn = k0 + 32 + lane;      // prefetch the next batch
      if (GATHER) {
        if (valid) st_rec(&G.out[k0 + lane], pm);
        if (kn < end) { nxt = ld_rec(&G.store[slot1]); nkey = skeys[kn]; }
        if (kn + 32 < end) slot1 = G.perm[kn + 32];
      } else if (kn < end) { nxt = ld_rec(&ps[kn]); nkey = skeys[kn]; }
    }
    SphDep d;
    sph_deposit_setup(g, p0, pm, adb, valid ? cell : 0, d);
    if (!valid) d.w1 = d.w2 = 0.0;

    // distinct cells of this batch in increasing order (almost always exactly one)
    int c = __reduce_min_sync(0xffffffffu, cell);
    while (true) {
      if (c != cur) {
        if (cur >= 0) flush(cur);
        cur = c;
      }
      const bool mine = cell == c;
      const double a = mine ? d.w1 : 0.0, b = mine ? d.w2 : 0.0;
      sph_deposit_rows<L, M0, M1>(d, c_inv, [&](int l, int m, int cs, double v) {
        const int q = sph_compact<L, M0>(l, m, cs);
        A[q] = fma(a, v, A[q]);
        B[q] = fma(b, v, B[q]);
      });
      const int nc = __reduce_min_sync(0xffffffffu, (cell > c && valid) ? cell : 0x7fffffff);
      if (nc == 0x7fffffff) break;
      c = nc;
    }
  }
  if (cur >= 0) flush(cur);
}

// ---------------------------------------------------------------------------
// deposit (lmax > 6): one warp per chunk of CH sorted particles.
// phase 1 (lane = particle): rows -> shared tile; phase 2 (lane = row): walk the 32
// particles in order, accumulate w1*Y and w2*Y, flush when the radial cell changes.
// ---------------------------------------------------------------------------
template <int L>
__global__ void __launch_bounds__(kDepositWarps * 32)
sph_deposit_kernel(SphGeom g, const double *__restrict__ p0, const rec4 *__restrict__ ps,
                   const uint32_t *__restrict__ skeys, const uint32_t *__restrict__ seg,
                   int keybase0, int CH, double adb, int batched,
                   double *__restrict__ PA, double *__restrict__ PB)
{
  using S = DepositShape<L>;
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double *tile = smem + (size_t)w * S::WARP_DOUBLES;
  int *cells = reinterpret_cast<int *>(tile + 32 * S::STRIDE);

  const int keybase = keybase0 + blockIdx.y * g.numr;
  const uint32_t lo = seg[keybase], hi = seg[keybase + g.numr - 1];
  const size_t sbase = batched ? (size_t)blockIdx.y * (g.numr + 2) + (lo - seg[keybase0]) / CH : 0;
  const size_t chunk = (size_t)blockIdx.x * kDepositWarps + w;
  const size_t beg = (size_t)lo + chunk * (size_t)CH;
  if (beg >= hi) return;
  const size_t end = beg + CH < hi ? beg + CH : hi;

  double accA[S::RPL], accB[S::RPL];
#pragma unroll
  for (int q = 0; q < S::RPL; q++) accA[q] = accB[q] = 0.0;
  int cur = -1;

  auto flush = [&](int c) {
    size_t slot = (sbase + chunk + (size_t)c) * S::NROW;
#pragma unroll
    for (int q = 0; q < S::RPL; q++) {
      int r = lane + 32 * q;
      if (r < S::NROW) {
        PA[slot + r] = accA[q];
        PB[slot + r] = accB[q];
      }
      accA[q] = accB[q] = 0.0;
    }
  };

  double *my = tile + lane * S::STRIDE;
  for (size_t k0 = beg; k0 < end; k0 += 32) {
    const size_t k = k0 + lane;
    int cell = -1;
    double w1 = 0.0, w2 = 0.0;
    if (k < end) {
      rec4 pm = ld_rec(&ps[k]);
      cell = (int)skeys[k] - keybase;
      SphDep d;
      sph_deposit_setup(g, p0, pm, adb, cell, d);
      w1 = d.w1; w2 = d.w2;
      sph_deposit_rows<L, 0, L>(d, c_inv, [&](int l, int m, int cs, double v) { my[sph_row(l, m) + cs] = v; });
    } else {
      for (int r = 0; r < S::NROW; r++) my[r] = 0.0;
    }
    my[S::NROW] = w1;
    my[S::NROW + 1] = w2;
    cells[lane] = cell;
    __syncwarp();

    const int cnt = end - k0 < 32 ? (int)(end - k0) : 32;
    for (int jj = 0; jj < cnt; jj++) {
      const int c = cells[jj];                 // warp-uniform
      if (c < 0) continue;
      if (c != cur) {
        if (cur >= 0) flush(cur);
        cur = c;
      }
      const double *row = tile + jj * S::STRIDE;
      const double a = row[S::NROW], b = row[S::NROW + 1];
#pragma unroll
      for (int q = 0; q < S::RPL; q++) {
        int r = lane + 32 * q;
        double v = r < S::NROW ? row[r] : 0.0;
        accA[q] += a * v;
        accB[q] += b * v;
      }
    }
    __syncwarp();
  }
  if (cur >= 0) flush(cur);
}

// ---------------------------------------------------------------------------
// combine: G[row][i] = sum over the chunks holding cell i of PA + same for cell i-1 of PB,
// chunks visited in increasing order (fixed order).
// ---------------------------------------------------------------------------
__global__ void sph_combine_kernel(SphGeom g, const uint32_t *__restrict__ seg, int keybase0, int CH,
                                   int batched, const double *__restrict__ PA,
                                   const double *__restrict__ PB, double *__restrict__ G)
{
  const int i = blockIdx.x;
  const int nrow = g.nrow;
  const int keybase = keybase0 + blockIdx.y * g.numr;
  const uint32_t lo = seg[keybase];
  const size_t sbase = batched ? (size_t)blockIdx.y * (g.numr + 2) + (lo - seg[keybase0]) / CH : 0;
  PA += sbase * nrow;
  PB += sbase * nrow;
  G += (size_t)blockIdx.y * nrow * g.numr;
  for (int row = threadIdx.x; row < nrow; row += blockDim.x) {
    double sum = 0.0;
    if (i <= g.numr - 2) {
      uint32_t s0 = seg[keybase + i], s1 = seg[keybase + i + 1];
      if (s1 > s0) {
        size_t j0 = (s0 - lo) / CH, j1 = (s1 - 1 - lo) / CH;
        for (size_t j = j0; j <= j1; j++) sum += PA[(j + i) * nrow + row];
      }
    }
    if (i >= 1) {
      const int c = i - 1;
      uint32_t s0 = seg[keybase + c], s1 = seg[keybase + c + 1];
      if (s1 > s0) {
        size_t j0 = (s0 - lo) / CH, j1 = (s1 - 1 - lo) / CH;
        double sb = 0.0;
        for (size_t j = j0; j <= j1; j++) sb += PB[(j + c) * nrow + row];
        sum += sb;
      }
    }
    G[(size_t)row * g.numr + i] = sum;
  }
}

// Fixed-order CTA sum of one double per thread (256 threads).
__device__ __forceinline__ double block_sum_256(double v, double *sh)
{
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < 8; k++) t += sh[k];
  }
  return t;   // valid in thread 0
}

// contract: coef[row][n] = K_lm sum_i E[l][n][i] * G[row][i]; one CTA per (row, n).
__global__ void __launch_bounds__(256)
sph_contract_kernel(SphGeom g, const double *__restrict__ Ep, const double *__restrict__ G,
                    const int *__restrict__ row_l, const int *__restrict__ row_m, int M0_only,
                    double *__restrict__ coef)
{
  __shared__ double sh[8];
  const int row = blockIdx.x, n = blockIdx.y;
  const int l = row_l[row], m = row_m[row];
  const bool dead = M0_only && m != 0;              // src/SphericalBasis.cc:549
  G += (size_t)blockIdx.z * g.nrow * g.numr;        // batched over covariance bins
  coef += (size_t)blockIdx.z * g.nrow * g.nmax;
  const double *Gr = G + (size_t)row * g.numr;
  const double *E = Ep + ((size_t)l * g.nmax + n) * g.numr;
  double s = 0.0;
  for (int i = threadIdx.x; i < g.numr; i += 256) s += E[i] * Gr[i];
  double t = block_sum_256(s, sh);
  if (threadIdx.x == 0) coef[(size_t)row * g.nmax + n] = dead ? 0.0 : c_K[l * (kMaxL + 1) + m] * t;
}

// field: T[i][slot][0..4] from H[i][slot] = K_lm (sum_n E[l][n][i] coef[rowc][n],
// sum_n E[l][n][i] coef[rows][n]) at nodes i-1, i, i+1 (sph_poly_entry).
// slot_rowc/slot_rows encode which coefficient rows the reference's force loop reads for
// this (l,m), including its NO_L0/NO_L1/EVEN_L/EVEN_M/M0_only behaviour; -1 = no term.
__global__ void __launch_bounds__(128)
sph_field_kernel(SphGeom g, const double *__restrict__ Ep, const double *__restrict__ p0,
                 const double *__restrict__ coef, const int *__restrict__ slot_l,
                 const int *__restrict__ slot_m, const int *__restrict__ slot_rowc,
                 const int *__restrict__ slot_rows, double2 *__restrict__ T)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int slot = blockIdx.y;
  if (i >= g.numr) return;
  const int l = slot_l[slot], m = slot_m[slot], rc = slot_rowc[slot], rs = slot_rows[slot];
  double2 H[3];
  for (int k = 0; k < 3; k++) H[k] = make_double2(0.0, 0.0);
  if (rc >= 0) {
    const double K = c_K[l * (kMaxL + 1) + m];
    const double *cc = coef + (size_t)rc * g.nmax;
    const double *cs = rs >= 0 ? coef + (size_t)rs * g.nmax : nullptr;
    for (int k = 0; k < 3; k++) {
      const int node = i - 1 + k;
      if (node < 0 || node >= g.numr) continue;
      const double *E = Ep + (size_t)l * g.nmax * g.numr + node;
      double c = 0.0, s = 0.0;
      for (int n = 0; n < g.nmax; n++) {
        double e = E[(size_t)n * g.numr];
        c += e * cc[n];
        if (cs) s += e * cs[n];
      }
      H[k] = make_double2(K * c, K * s);
    }
  }
  double2 out[kSphPoly];
  sph_poly_entry(g, p0, i, H[0], H[1], H[2], out);
  double2 *dst = T + ((size_t)i * g.nslot + slot) * kSphPoly;
  for (int k = 0; k < kSphPoly; k++) dst[k] = out[k];
}

// ---- TMA bulk copy (global -> shared) completing on an mbarrier --------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tma_bulk_load(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned phase)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}

constexpr int kForceRows = 6;     // table rows (radial cells) a CTA stages in shared memory

// What the force kernels do with a particle's result.
//   mode 0: acc/pot += (Component::AddAcc)
//   mode 1: fused single-level step: acc/pot = value (they would have been zeroed first) and
//           the second half kick vel += acc*dtk
//   mode 2: mode 1, then the NEXT step's first half kick, drift and radial key (the step
//           driver runs several steps back to back: src/step.cc:272-322 in a loop)
struct ForceOut {
  rec4 *accp, *vel, *posm;
  uint32_t *keys;
  double dtk, dtd;
  int mode;
};

// pm: the particle's position record (already in registers); v: its velocity record,
// loaded by the caller before the evaluation so that its latency is hidden (modes 1, 2).
__device__ __forceinline__ void force_store(const SphGeom &g, const ForceOut &F, uint32_t i, bool ok,
                                            const double *o, rec4 pm, rec4 v)
{
  if (F.mode == 0) {
    if (!ok) return;
    rec4 a = ld_rec_rw(&F.accp[i]);
    a.x += o[0]; a.y += o[1]; a.z += o[2]; a.w += o[3];
    st_rec(&F.accp[i], a);
    return;
  }
  const rec4 a = rec4{0.0 + o[0], 0.0 + o[1], 0.0 + o[2], 0.0 + o[3]};
  st_rec(&F.accp[i], a);
  v.x = fma(a.x, F.dtk, v.x); v.y = fma(a.y, F.dtk, v.y); v.z = fma(a.z, F.dtk, v.z);
  if (F.mode == 2) {
    v.x = fma(a.x, F.dtk, v.x); v.y = fma(a.y, F.dtk, v.y); v.z = fma(a.z, F.dtk, v.z);
    rec4 p = pm;
    p.x = fma(v.x, F.dtd, p.x); p.y = fma(v.y, F.dtd, p.y); p.z = fma(v.z, F.dtd, p.z);
    st_rec(&F.posm[i], p);
    F.keys[i] = sph_cell_of(g, p);
  }
  st_rec(&F.vel[i], v);
}

// force_store with the record to combine with (mode 0: acc/pot, modes 1-2: velocity) already
// fetched by the caller
__device__ __forceinline__ void force_store_side(const SphGeom &g, const ForceOut &F, uint32_t i, bool ok,
                                                 const double *o, rec4 pm, rec4 v)
{
  if (F.mode == 0) {
    if (!ok) return;
    v.x += o[0]; v.y += o[1]; v.z += o[2]; v.w += o[3];
    st_rec(&F.accp[i], v);
    return;
  }
  force_store(g, F, i, ok, o, pm, v);
}

// force: one thread per particle, sorted order (the 128 particles of a CTA span a few
// radial cells), positions streamed from the sorted copy.  The CTA finds the range of
// table rows its particles read and, when it is at most kForceRows, stages those rows of T
// (contiguous, 2240 B each at lmax=6) into shared memory with one TMA bulk copy; otherwise
// (sparse outskirts, the unsorted dead-cell tail) it reads T through L1.
// mode 0: acc/pot += (Component::AddAcc); mode 1 (fused step): acc/pot = value (they would
// have been zeroed first).
template <int L>
__global__ void __launch_bounds__(kForceThreads, 4)
sph_force_kernel(SphGeom g, const double2 *__restrict__ T, const rec4 *ps, ForceOut F,
                 const uint32_t *__restrict__ perm, const uint32_t *__restrict__ seg, int keybase,
                 size_t n)
{
  constexpr int ROW = (L + 1) * (L + 2) / 2 * kSphPoly;      // double2 per table row
  __shared__ __align__(128) double2 rows[kForceRows * ROW];
  __shared__ __align__(8) uint64_t bar;
  __shared__ int s_lo, s_hi;

  if (threadIdx.x == 0) {
    s_lo = 0x7fffffff;
    s_hi = -1;
    mbar_init(&bar, 1);
  }
  __syncthreads();

  const size_t lo = seg[keybase];
  const size_t k = lo + blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const bool valid = k < n;
  rec4 pm = rec4{g.cx + 1.0, g.cy, g.cz, 0.0};
  uint32_t i = 0;
  rec4 vv = rec4{0.0, 0.0, 0.0, 0.0};
  if (valid) {
    pm = ld_rec_rw(&ps[k]);
    i = perm[k];
    if (F.mode) vv = ld_rec_rw(&F.vel[i]);
  }

  SphForce S;
  const bool ok = sph_force_setup<true>(g, pm, S) && valid;
  {
    int wmin = __reduce_min_sync(0xffffffffu, ok ? S.cell : 0x7fffffff);
    int wmax = __reduce_max_sync(0xffffffffu, ok ? S.j : -1);      // j >= cell
    if ((threadIdx.x & 31) == 0) {
      atomicMin(&s_lo, wmin);
      atomicMax(&s_hi, wmax);
    }
  }
  __syncthreads();
  const int rlo = s_lo, rhi = s_hi;
  const bool staged = rhi >= rlo && rhi - rlo < kForceRows;
  const double2 *Tc, *Td;
  if (staged) {
    if (threadIdx.x == 0)
      tma_bulk_load(rows, T + (size_t)rlo * ROW, (uint32_t)((rhi - rlo + 1) * ROW * sizeof(double2)), &bar);
    Tc = rows + (ok ? S.cell - rlo : 0) * ROW;
    Td = rows + (ok ? S.j - rlo : 0) * ROW;
    mbar_wait(&bar, 0);
  } else {
    Tc = T + (size_t)(ok ? S.cell : 0) * ROW;
    Td = T + (size_t)(ok ? S.j : 0) * ROW;
  }

  double ax = 0.0, ay = 0.0, az = 0.0, pot = 0.0;
  const bool far = __any_sync(0xffffffffu, ok && S.ioff);
  if (ok) {
    if (staged) {
      auto ld = [](const double2 *a) { return *a; };
      if (far) sph_force_eval<L, true>(c_inv, S, Tc, Td, ld, ax, ay, az, pot);
      else     sph_force_eval<L, false>(c_inv, S, Tc, Td, ld, ax, ay, az, pot);
    } else {
      auto ld = [](const double2 *a) { return __ldg(a); };
      if (far) sph_force_eval<L, true>(c_inv, S, Tc, Td, ld, ax, ay, az, pot);
      else     sph_force_eval<L, false>(c_inv, S, Tc, Td, ld, ax, ay, az, pot);
    }
  }
  if (!valid) return;
  const double o[4] = {ax, ay, az, pot};
  force_store(g, F, i, ok, o, pm, vv);
}

// force, two particles per lane (adjacent in the sorted order, so they almost always read the
// same table rows): halves the table bytes delivered per particle, which is what bounds the
// one-particle form.  CTAs of 64 threads = 128 particles, staging as in sph_force_kernel.
constexpr int kForce2Threads = 64;

template <int L>
__global__ void __launch_bounds__(kForce2Threads, 5)
sph_force2_kernel(SphGeom g, const double2 *__restrict__ T, const rec4 *ps, ForceOut F,
                  const uint32_t *__restrict__ perm, const uint32_t *__restrict__ seg, int keybase,
                  size_t n)
{
  constexpr int ROW = (L + 1) * (L + 2) / 2 * kSphPoly;      // double2 per table row
  __shared__ __align__(128) double2 rows[kForceRows * ROW];
  __shared__ __align__(8) uint64_t bar;
  __shared__ int s_lo, s_hi;

  if (threadIdx.x == 0) {
    s_lo = 0x7fffffff;
    s_hi = -1;
    mbar_init(&bar, 1);
  }
  __syncthreads();

  const size_t lo = seg[keybase];
  const size_t k0 = lo + (blockIdx.x * (size_t)kForce2Threads + threadIdx.x) * 2;
  const bool validA = k0 < n, validB = k0 + 1 < n;
  rec4 pmA = rec4{g.cx + 1.0, g.cy, g.cz, 0.0}, pmB = pmA;
  uint32_t iA = 0, iB = 0;
  rec4 vA = rec4{0.0, 0.0, 0.0, 0.0}, vB = vA;
  if (validA) { pmA = ld_rec_rw(&ps[k0]); iA = perm[k0]; if (F.mode) vA = ld_rec_rw(&F.vel[iA]); }
  if (validB) { pmB = ld_rec_rw(&ps[k0 + 1]); iB = perm[k0 + 1]; if (F.mode) vB = ld_rec_rw(&F.vel[iB]); }

  SphForce A, B;
  const bool okA = sph_force_setup<true>(g, pmA, A) && validA;
  const bool okB = sph_force_setup<true>(g, pmB, B) && validB;
  {
    int mn = min(okA ? A.cell : 0x7fffffff, okB ? B.cell : 0x7fffffff);
    int mx = max(okA ? A.j : -1, okB ? B.j : -1);
    mn = __reduce_min_sync(0xffffffffu, mn);
    mx = __reduce_max_sync(0xffffffffu, mx);
    if ((threadIdx.x & 31) == 0) {
      atomicMin(&s_lo, mn);
      atomicMax(&s_hi, mx);
    }
  }
  __syncthreads();
  const int rlo = s_lo, rhi = s_hi;
  const bool staged = rhi >= rlo && rhi - rlo < kForceRows;
  if (staged) {
    if (threadIdx.x == 0)
      tma_bulk_load(rows, T + (size_t)rlo * ROW, (uint32_t)((rhi - rlo + 1) * ROW * sizeof(double2)), &bar);
    mbar_wait(&bar, 0);
  }

  double oa[4] = {0.0, 0.0, 0.0, 0.0}, ob[4] = {0.0, 0.0, 0.0, 0.0};
  const bool pair = okA && okB && A.cell == B.cell && A.j == B.j && !A.ioff && !B.ioff;
  if (pair) {
    if (staged)
      sph_force_eval2<L>(c_inv, A, B, rows + (A.cell - rlo) * ROW, rows + (A.j - rlo) * ROW,
                         [](const double2 *a) { return *a; }, oa, ob);
    else
      sph_force_eval2<L>(c_inv, A, B, T + (size_t)A.cell * ROW, T + (size_t)A.j * ROW,
                         [](const double2 *a) { return __ldg(a); }, oa, ob);
  } else {
    // rare: the pair straddles a cell boundary, or a particle beyond rmax (generic loads)
    const double2 *base = staged ? rows - (size_t)rlo * ROW : T;
    auto ld = [](const double2 *a) { return *a; };
    if (okA) sph_force_eval<L, true>(c_inv, A, base + (size_t)A.cell * ROW, base + (size_t)A.j * ROW, ld,
                                     oa[0], oa[1], oa[2], oa[3]);
    if (okB) sph_force_eval<L, true>(c_inv, B, base + (size_t)B.cell * ROW, base + (size_t)B.j * ROW, ld,
                                     ob[0], ob[1], ob[2], ob[3]);
  }

  if (validA) force_store(g, F, iA, okA, oa, pmA, vA);
  if (validB) force_store(g, F, iB, okB, ob, pmB, vB);
}

// force, persistent and software-pipelined (the default): a CTA walks tiles of TILE = 2*TH
// sorted particles (tile t = absolute sorted positions [t*TILE, (t+1)*TILE), so every bulk
// copy is 16-byte aligned) and keeps a ring of NST stages in shared memory.  While the lanes
// evaluate tile t from one stage, three TMA bulk copies fill the next stage for the CTA's
// next tile: the position records (contiguous in the sorted copy), the permutation slice,
// and the table rows the tile reads.  The row range is known WITHOUT looking at the
// particles: the tile is sorted by key, so its first and last sorted key bound the radial
// cells (sph_tile_desc_kernel, a pre-pass of one thread per tile).  The lane's two velocity
// records are gathered with cp.async into shared memory, and the position records are
// re-read from the stage after the evaluation, so no register (and no spill) waits on a
// load.  Nothing the evaluation needs is then a dependent global load: the long-scoreboard
// stalls of the one-shot kernel (sph_force2_kernel: positions -> cells -> atomics -> TMA ->
// wait, all serial, per CTA) are off the critical path.  Every lane still checks its own
// rows against the staged range and falls back to global loads when outside (tiles that
// straddle levels, the dead-cell tail).
template <int L, int TH, int ROWS> struct FPStage {
  static constexpr int kFPTile = 2 * TH;
  static constexpr int ROW = (L + 1) * (L + 2) / 2 * kSphPoly;
  double2 rows[ROWS * ROW];
  rec4 ps[kFPTile];
  uint32_t perm[kFPTile];
};
template <int L, int TH, int ROWS, int NST> struct FPShared {
  FPStage<L, TH, ROWS> stage[NST];
  rec4 vbuf[2 * TH];            // the lane's two velocity (mode 0: acc/pot) records, via cp.async
};
struct FPDesc { int rlo, nrows; };

__device__ __forceinline__ void tma_bulk_copy(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint64_t *bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// register-free global -> shared copies (LDGSTS): the data lands in shared memory while the
// lane computes, so no register (and no spill) waits on the load
__device__ __forceinline__ void cp_async_rec(rec4 *dst_smem, const rec4 *src)
{
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n"
               "cp.async.cg.shared.global [%2], [%3], 16;"
               ::"r"(smem_u32(dst_smem)), "l"(src), "r"(smem_u32(dst_smem) + 16),
                 "l"(reinterpret_cast<const char *>(src) + 16) : "memory");
}
__device__ __forceinline__ void cp_async_u64(void *dst_smem, const void *src)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}

// Row range of every tile of the pipelined force kernel, from the first and last sorted key of
// the tile (the tile is sorted by key): out[t] = (first table row, number of rows), rows = 0
// when the tile straddles two levels, touches the dead cell or needs more than `rows_max` rows
// (its lanes then read the table through L1/L2).
__global__ void __launch_bounds__(256)
sph_tile_desc_kernel(const uint32_t *__restrict__ skeys, const uint32_t *__restrict__ seg, int keybase,
                     size_t n, int tile, int numr, int rows_max, int2 *__restrict__ out)
{
  const size_t lo = seg[keybase];
  const size_t t_end = (n + tile - 1) / tile;
  const size_t t = lo / tile + blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= t_end) return;
  const size_t a = t * tile < lo ? lo : t * tile;
  const size_t b = (t + 1) * tile <= n ? (t + 1) * tile - 1 : n - 1;
  const uint32_t kf = skeys[a < b ? a : b], kl = skeys[b];
  const int lf = (int)(kf / (uint32_t)numr), ll = (int)(kl / (uint32_t)numr);
  const int cf = (int)(kf - (uint32_t)lf * numr), cl = (int)(kl - (uint32_t)ll * numr);
  const int rlo = cf, rhi = cl < 1 ? 1 : cl;       // stencil centre j = max(cell, 1)
  int nrows = rhi - rlo + 1;
  if (lf != ll || cl >= numr - 1 || nrows > rows_max || nrows < 1) nrows = 0;
  out[t] = make_int2(rlo, nrows);
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int L, int TH, int MINB, int ROWS, int NST>
__global__ void __launch_bounds__(TH, MINB)
sph_force2p_kernel(SphGeom g, const double2 *__restrict__ T, const rec4 *ps, ForceOut F,
                   const uint32_t *__restrict__ perm, const int2 *__restrict__ tdesc,
                   const uint32_t *__restrict__ seg, int keybase, size_t n)
{
  using Stage = FPStage<L, TH, ROWS>;
  constexpr int kFPTile = 2 * TH;
  constexpr int ROW = Stage::ROW;
  extern __shared__ __align__(128) unsigned char fp_smem[];
  FPShared<L, TH, ROWS, NST> &SH = *reinterpret_cast<FPShared<L, TH, ROWS, NST> *>(fp_smem);
  __shared__ __align__(8) uint64_t full[NST];
  __shared__ FPDesc desc[NST];
  __shared__ __align__(8) int2 ndesc;    // descriptor of the tile after next (cp.async)

  const size_t lo = seg[keybase];
  const size_t t_end = (n + kFPTile - 1) / kFPTile;
  const size_t t0 = lo / kFPTile + blockIdx.x;
  const size_t stride = gridDim.x;
  if (t0 >= t_end) return;
  const int tid = threadIdx.x;

  // thread 0 starts the copies of tile t into stage s.  The particle arrays are allocated in
  // whole tiles (component.cu pads every array), so a tile is always copied in full; lanes
  // outside [lo, n) ignore what they get.
  auto issue = [&](int s, size_t t, int2 nd) {
    const int rlo = nd.x, nrows = nd.y;
    const size_t k0 = t * kFPTile;
    desc[s].rlo = rlo;
    desc[s].nrows = nrows;
    const uint32_t rbytes = (uint32_t)nrows * ROW * (uint32_t)sizeof(double2);
    mbar_expect(&full[s], rbytes + kFPTile * (uint32_t)(sizeof(rec4) + sizeof(uint32_t)));
    if (nrows) tma_bulk_copy(SH.stage[s].rows, T + (size_t)rlo * ROW, rbytes, &full[s]);
    tma_bulk_copy(SH.stage[s].ps, ps + k0, kFPTile * (uint32_t)sizeof(rec4), &full[s]);
    tma_bulk_copy(SH.stage[s].perm, perm + k0, kFPTile * (uint32_t)sizeof(uint32_t), &full[s]);
  };

  if (tid == 0) {
    for (int q = 0; q < NST; q++) {
      mbar_init(&full[q], 1);
    }
    for (int q = 0; q < NST; q++)
      if (t0 + q * stride < t_end) issue(q, t0 + q * stride, __ldg(tdesc + t0 + q * stride));
  }
  __syncthreads();

  const rec4 *side = F.mode ? F.vel : F.accp;     // the record force_store combines the result with
  rec4 *myv = &SH.vbuf[2 * tid];
  // The warps of a CTA stay in lockstep (one block barrier per tile): they run the same
  // long straight-line code on the four sub-partitions at the same time and share its
  // instruction fetches; letting them drift (per-stage "empty" barriers) measured slower.
  int s = 0, ph = 0;
  for (size_t t = t0; t < t_end; t += stride) {
    __syncthreads();               // the stage used by the previous tile is drained
    if (tid == 0) {
      const size_t tn = t + (NST - 1) * stride;    // refill the stage the previous tile used
      if (t != t0 && tn < t_end) {
        cp_async_wait_all();
        issue(s == 0 ? NST - 1 : s - 1, tn, ndesc);
      }
      if (tn + stride < t_end) cp_async_u64(&ndesc, tdesc + tn + stride);
    }
    mbar_wait(&full[s], ph);
    const FPDesc d = desc[s];
    const Stage &S = SH.stage[s];

    const size_t k0 = t * kFPTile + 2 * (size_t)tid;
    const bool validA = k0 >= lo && k0 < n, validB = k0 + 1 >= lo && k0 + 1 < n;
    uint32_t iA, iB;
    SphForce A, B;
    bool okA, okB;
    {
      const rec4 dummy = rec4{g.cx + 1.0, g.cy, g.cz, 0.0};
      const rec4 pmA = validA ? S.ps[2 * tid] : dummy, pmB = validB ? S.ps[2 * tid + 1] : dummy;
      const uint2 ii = *reinterpret_cast<const uint2 *>(&S.perm[2 * tid]);
      iA = validA ? ii.x : 0; iB = validB ? ii.y : 0;
      if (validA) cp_async_rec(myv, side + iA);
      if (validB) cp_async_rec(myv + 1, side + iB);
      okA = sph_force_setup<true>(g, pmA, A) && validA;
      okB = sph_force_setup<true>(g, pmB, B) && validB;
    }
    const int rlo = d.rlo, rhi = d.rlo + d.nrows - 1;
    const bool inA = A.cell >= rlo && A.j <= rhi, inB = B.cell >= rlo && B.j <= rhi;

    double oa[4] = {0.0, 0.0, 0.0, 0.0}, ob[4] = {0.0, 0.0, 0.0, 0.0};
    const bool pair = okA && okB && A.cell == B.cell && A.j == B.j && !A.ioff && !B.ioff;
    if (pair && inA) {
      sph_force_eval2<L>(c_inv, A, B, S.rows + (A.cell - rlo) * ROW, S.rows + (A.j - rlo) * ROW,
                         [](const double2 *a) { return *a; }, oa, ob);
    } else if (pair) {
      sph_force_eval2<L>(c_inv, A, B, T + (size_t)A.cell * ROW, T + (size_t)A.j * ROW,
                         [](const double2 *a) { return __ldg(a); }, oa, ob);
    } else {
      // rare: the pair straddles a cell boundary, or a particle beyond rmax (generic loads)
      auto ld = [](const double2 *a) { return *a; };
      const double2 *bA = inA ? S.rows - (ptrdiff_t)rlo * ROW : T;
      const double2 *bB = inB ? S.rows - (ptrdiff_t)rlo * ROW : T;
      if (okA) sph_force_eval<L, true>(c_inv, A, bA + (size_t)A.cell * ROW, bA + (size_t)A.j * ROW, ld,
                                       oa[0], oa[1], oa[2], oa[3]);
      if (okB) sph_force_eval<L, true>(c_inv, B, bB + (size_t)B.cell * ROW, bB + (size_t)B.j * ROW, ld,
                                       ob[0], ob[1], ob[2], ob[3]);
    }

    // the position records are re-read from the stage rather than held in registers across
    // the evaluation
    cp_async_wait_all();
    if (validA) force_store_side(g, F, iA, okA, oa, S.ps[2 * tid], myv[0]);
    if (validB) force_store_side(g, F, iB, okB, ob, S.ps[2 * tid + 1], myv[1]);
    if (++s == NST) { s = 0; ph ^= 1; }
  }
}

// ---------------------------------------------------------------------------
// force, warp-specialised (EXPB200_FORCE_VARIANT=3; kept as a measured alternative).
//
// What bounded sph_force2p_kernel (profiles/r01f_ncu_full_summary.txt): the dense (l,m)
// evaluation runs at the FP64 pipe's rate but was only ~41 % of a warp's time; the geometric
// set-up (sqrt / reciprocal chains), the epilogue (projection, kicks, drift, next key, global
// gathers and stores) and the tile barrier are latency-bound, and at 168 registers only three
// warps per scheduler were there to cover them.  Here the two kinds of work run in different
// warps of one CTA:
//   warps 0-3 ("evaluation", 168 registers each after setmaxnreg.inc) do nothing but the
//     (l,m) loops, two particles per lane, from operands already waiting in shared memory;
//   warps 4-7 ("helpers", 88 registers after setmaxnreg.dec) do the set-up of the NEXT tile,
//     the epilogue of the PREVIOUS one, and issue the TMA copies.
// Two such CTAs are resident per SM: every scheduler holds two evaluation warps that keep the
// FP64 pipe busy back to back and two helper warps whose latencies nobody waits for.
//
// A CTA walks tiles of 256 sorted particles (tile i of the CTA = "window" i).  In window i:
//   evaluation warp w : particles 64w..64w+63 of tile i   : prep[i&1] + rows[i&1] -> res[i&1]
//   helper warp w     : epilogue of tile i-1 (res, keep, ps of i-1 -> global memory),
//                       then set-up of tile i+1 (ps[(i+1)%3] -> prep[(i+1)&1], keep[(i+1)&1])
//   helper thread 224 : TMA of the table rows of tile i+1 and of the particles of tile i+2
// with one CTA barrier per window (the evaluation is the long pole; the helpers have slack),
// so every buffer is written in one window and read in the next.  The arithmetic per particle
// is that of sph_force_setup / sph_force_sums(2) / sph_force_project, i.e. of the other force
// kernels.
// ---------------------------------------------------------------------------
constexpr int kWSTile = 256;
constexpr int kWSPairs = kWSTile / 2;
constexpr int kWSThreads = 256;
constexpr int kWSSlots = 4;           // particle-record ring: tiles i-1 (epilogue), i, i+1 (set-up), i+2 (in flight)

template <int L, int ROWS> struct WSShared {
  static constexpr int ROW = (L + 1) * (L + 2) / 2 * kSphPoly;
  double2 rows[2][ROWS * ROW];
  rec4 ps[kWSSlots][kWSTile];
  uint32_t perm[kWSSlots][kWSTile];
  // set-up -> evaluation; .x / .y = first / second particle of a lane
  double2 costh[2][kWSPairs], c1[2][kWSPairs], s1[2][kWSPairs], p[2][kWSPairs], q[2][kWSPairs];
  int2 info[2][kWSPairs];                             // cell | (j != cell) << 16 | ok << 17 | ioff << 18
  // evaluation -> epilogue
  double2 sums[2][4][kWSPairs];
  // set-up -> epilogue (ir: 1/r of the particle itself; beyond rmax the evaluation derives the
  // continuation factors from it and the epilogue uses 1/rmax)
  double2 ir[2][kWSPairs], iR2[2][kWSPairs], somx2i[2][kWSPairs], fac[2][kWSPairs];
};

__device__ __forceinline__ double &d2_at(double2 &v, int k) { return reinterpret_cast<double *>(&v)[k]; }
__device__ __forceinline__ int &i2_at(int2 &v, int k) { return reinterpret_cast<int *>(&v)[k]; }

template <int L, int ROWS, int REGE, int REGH>
__global__ void __launch_bounds__(kWSThreads, 2)
sph_force_ws_kernel(SphGeom g, const double2 *__restrict__ T, const rec4 *ps, ForceOut F,
                    const uint32_t *__restrict__ perm, const int2 *__restrict__ tdesc,
                    const uint32_t *__restrict__ seg, int keybase, size_t n)
{
  static_assert(REGE + REGH == 256 && REGE % 8 == 0, "the two roles share 2 x 128 registers per thread pair");
  using SHT = WSShared<L, ROWS>;
  constexpr int ROW = SHT::ROW;
  extern __shared__ __align__(128) unsigned char ws_smem[];
  SHT &SH = *reinterpret_cast<SHT *>(ws_smem);
  __shared__ __align__(8) uint64_t rows_full[2], ps_full[kWSSlots];
  __shared__ FPDesc desc[2];

  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  const bool is_eval = wid < 4;
  const int pi = (wid & 3) * 32 + lane;             // the lane's particle pair within a tile
  const bool issuer = tid == kWSThreads - 32;

  const size_t lo = seg[keybase];
  const size_t t_end = (n + kWSTile - 1) / kWSTile;
  const size_t t0 = lo / kWSTile + blockIdx.x;
  const size_t stride = gridDim.x;
  const int ntiles = t0 < t_end ? (int)((t_end - t0 + stride - 1) / stride) : 0;

  if (tid == 0) {
    for (int k = 0; k < 2; k++) mbar_init(&rows_full[k], 1);
    for (int k = 0; k < kWSSlots; k++) mbar_init(&ps_full[k], 1);
  }
  __syncthreads();
  if (ntiles == 0) return;

  auto issue_rows = [&](int i) {
    const int s = i & 1;
    const int2 nd = __ldg(tdesc + t0 + (size_t)i * stride);
    desc[s].rlo = nd.x;
    desc[s].nrows = nd.y;
    const uint32_t bytes = (uint32_t)nd.y * ROW * (uint32_t)sizeof(double2);
    mbar_expect(&rows_full[s], bytes);
    if (nd.y) tma_bulk_copy(SH.rows[s], T + (size_t)nd.x * ROW, bytes, &rows_full[s]);
  };
  auto issue_ps = [&](int i) {
    const int s = i % kWSSlots;
    const size_t k0 = (t0 + (size_t)i * stride) * kWSTile;
    mbar_expect(&ps_full[s], kWSTile * (uint32_t)(sizeof(rec4) + sizeof(uint32_t)));
    tma_bulk_copy(SH.ps[s], ps + k0, kWSTile * (uint32_t)sizeof(rec4), &ps_full[s]);
    tma_bulk_copy(SH.perm[s], perm + k0, kWSTile * (uint32_t)sizeof(uint32_t), &ps_full[s]);
  };

  // ---- helper: set-up of tile i -------------------------------------------------------------
  auto setup_one = [&](int b, int s, size_t gk, int k) {
    const bool valid = gk >= lo && gk < n;
    const rec4 pm = valid ? SH.ps[s][2 * pi + k] : rec4{g.cx + 1.0, g.cy, g.cz, 0.0};
    SphForce S;
    const bool ok = sph_force_setup<true>(g, pm, S) && valid;
    d2_at(SH.costh[b][pi], k) = S.costh;
    d2_at(SH.c1[b][pi], k) = S.c1;
    d2_at(SH.s1[b][pi], k) = S.s1;
    d2_at(SH.p[b][pi], k) = S.p;
    d2_at(SH.q[b][pi], k) = S.q;
    i2_at(SH.info[b][pi], k) = ok ? (S.cell | ((S.j != S.cell) << 16) | (1 << 17) | ((int)S.ioff << 18)) : 0;
    // beyond rmax S.ir is 1/rmax and r0i = (1/r)/scale: keep 1/r itself
    d2_at(SH.ir[b][pi], k) = S.ioff ? S.r0i * g.scale : S.ir;
    d2_at(SH.iR2[b][pi], k) = S.iR2;
    d2_at(SH.somx2i[b][pi], k) = S.somx2i;
    d2_at(SH.fac[b][pi], k) = S.fac;
  };
  auto setup = [&](int i) {
    const int b = i & 1, s = i % kWSSlots;
    mbar_wait(&ps_full[s], (i / kWSSlots) & 1);
    const size_t k0 = (t0 + (size_t)i * stride) * kWSTile + 2 * (size_t)pi;
    setup_one(b, s, k0, 0);
    setup_one(b, s, k0 + 1, 1);
  };

  // ---- helper: epilogue of tile i -----------------------------------------------------------
  // The records the result is combined with (velocity, or acc/pot in mode 0) are gathered most
  // of a window ahead (prefetch_side(i) right after the epilogue of tile i-1, used by
  // epilogue(i) in the next window), so that no helper waits on a global load.
  const rec4 *side = F.mode ? F.vel : F.accp;
  rec4 sideA = rec4{0.0, 0.0, 0.0, 0.0}, sideB = sideA;
  auto prefetch_side = [&](int i) {
    const int s = i % kWSSlots;
    const size_t k0 = (t0 + (size_t)i * stride) * kWSTile + 2 * (size_t)pi;
    const uint2 ii = *reinterpret_cast<const uint2 *>(&SH.perm[s][2 * pi]);
    if (k0 >= lo && k0 < n) sideA = ld_rec_rw(side + ii.x);
    if (k0 + 1 >= lo && k0 + 1 < n) sideB = ld_rec_rw(side + ii.y);
  };
  auto epilogue_one = [&](int b, int s, size_t gk, int k, uint32_t idx, const rec4 &v) {
    if (!(gk >= lo && gk < n)) return;
    const int inf = i2_at(SH.info[b][pi], k);
    const bool ok = (inf >> 17) & 1, ioff = (inf >> 18) & 1;
    const rec4 pm = SH.ps[s][2 * pi + k];
    double o[4] = {0.0, 0.0, 0.0, 0.0};
    if (ok) {
      SphForce S;
      S.x = pm.x - g.cx; S.y = pm.y - g.cy; S.z = pm.z - g.cz;
      S.R2 = S.x * S.x + S.y * S.y;
      S.ir = ioff ? g.inv_rmax : d2_at(SH.ir[b][pi], k);
      S.iR2 = d2_at(SH.iR2[b][pi], k);
      S.somx2i = d2_at(SH.somx2i[b][pi], k);
      S.fac = d2_at(SH.fac[b][pi], k);
      const double sm[4] = {d2_at(SH.sums[b][0][pi], k), d2_at(SH.sums[b][1][pi], k),
                            d2_at(SH.sums[b][2][pi], k), d2_at(SH.sums[b][3][pi], k)};
      sph_force_project(S, ioff, sm, o);
    }
    force_store_side(g, F, idx, ok, o, pm, v);
  };
  auto epilogue = [&](int i) {
    const int b = i & 1, s = i % kWSSlots;
    const size_t k0 = (t0 + (size_t)i * stride) * kWSTile + 2 * (size_t)pi;
    const uint2 ii = *reinterpret_cast<const uint2 *>(&SH.perm[s][2 * pi]);
    epilogue_one(b, s, k0, 0, ii.x, sideA);
    epilogue_one(b, s, k0 + 1, 1, ii.y, sideB);
  };

  // ---- evaluation of tile i -------------------------------------------------------------------
  auto evaluate = [&](int i) {
    const int b = i & 1;
    mbar_wait(&rows_full[b], (i >> 1) & 1);
    const FPDesc d = desc[b];
    const double2 *rows = SH.rows[b];
    const int2 inf = SH.info[b][pi];
    const bool okA = (inf.x >> 17) & 1, okB = (inf.y >> 17) & 1;
    SphForce A, B;
    {
      const double2 ct = SH.costh[b][pi], c1 = SH.c1[b][pi], s1 = SH.s1[b][pi], pp = SH.p[b][pi], qq = SH.q[b][pi];
      A.costh = ct.x; B.costh = ct.y;
      A.xd = sph_clamp_x(ct.x); B.xd = sph_clamp_x(ct.y);
      A.c1 = c1.x; B.c1 = c1.y; A.s1 = s1.x; B.s1 = s1.y;
      A.p = pp.x; B.p = pp.y; A.q = qq.x; B.q = qq.y;
    }
    A.cell = inf.x & 0xffff; A.j = A.cell + ((inf.x >> 16) & 1); A.ioff = (inf.x >> 18) & 1;
    B.cell = inf.y & 0xffff; B.j = B.cell + ((inf.y >> 16) & 1); B.ioff = (inf.y >> 18) & 1;
    A.ratio = B.ratio = 1.0; A.r0i = B.r0i = 0.0;
    const int rlo = d.rlo, rhi = d.rlo + d.nrows - 1;
    const bool inA = A.cell >= rlo && A.j <= rhi, inB = B.cell >= rlo && B.j <= rhi;
    double sa[4] = {0.0, 0.0, 0.0, 0.0}, sb[4] = {0.0, 0.0, 0.0, 0.0};
    const bool pair = okA && okB && A.cell == B.cell && A.j == B.j && !A.ioff && !B.ioff;
    if (pair && inA) {
      sph_force_sums2<L>(c_inv, A, B, rows + (A.cell - rlo) * ROW, rows + (A.j - rlo) * ROW,
                         [](const double2 *a) { return *a; }, sa, sb);
    } else if (pair) {
      sph_force_sums2<L>(c_inv, A, B, T + (size_t)A.cell * ROW, T + (size_t)A.j * ROW,
                         [](const double2 *a) { return __ldg(a); }, sa, sb);
    } else {
      // rare: the pair straddles a cell boundary, or a particle beyond rmax (generic loads);
      // the continuation factors as sph_force_setup forms them
      if (A.ioff) { const double ir0 = SH.ir[b][pi].x; A.ratio = g.rmax * ir0; A.r0i = ir0 * g.inv_scale; }
      if (B.ioff) { const double ir0 = SH.ir[b][pi].y; B.ratio = g.rmax * ir0; B.r0i = ir0 * g.inv_scale; }
      auto ld = [](const double2 *a) { return *a; };
      const double2 *bA = inA ? rows - (ptrdiff_t)rlo * ROW : T;
      const double2 *bB = inB ? rows - (ptrdiff_t)rlo * ROW : T;
      if (okA) sph_force_sums<L, true>(c_inv, A, bA + (size_t)A.cell * ROW, bA + (size_t)A.j * ROW, ld, sa);
      if (okB) sph_force_sums<L, true>(c_inv, B, bB + (size_t)B.cell * ROW, bB + (size_t)B.j * ROW, ld, sb);
    }
#pragma unroll
    for (int k = 0; k < 4; k++) SH.sums[b][k][pi] = make_double2(sa[k], sb[k]);
  };

  // The two roles never share code after this point, so that ptxas allocates registers for
  // each branch within its own setmaxnreg budget.  Both execute 1 + (ntiles + 1) CTA barriers.
  if (is_eval) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGE));
    asm volatile("bar.sync 0;" ::: "memory");               // prep of tile 0 is there
    for (int i = 0; i <= ntiles; i++) {
      if (i < ntiles) evaluate(i);
      asm volatile("bar.sync 0;" ::: "memory");
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGH));
    if (issuer) {
      issue_rows(0);
      issue_ps(0);
      if (ntiles > 1) issue_ps(1);
    }
    setup(0);
    prefetch_side(0);
    asm volatile("bar.sync 0;" ::: "memory");
    for (int i = 0; i <= ntiles; i++) {
      if (issuer) {
        if (i + 1 < ntiles) issue_rows(i + 1);      // rows slot (i+1)&1 was drained in window i-1
        if (i + 2 < ntiles) issue_ps(i + 2);        // ps slot (i+2)%4 == (i-2)%4: its epilogue ran in window i-1
      }
      if (i >= 1) epilogue(i - 1);
      if (i < ntiles) prefetch_side(i);             // for epilogue(i) in the next window
      if (i + 1 < ntiles) setup(i + 1);
      asm volatile("bar.sync 0;" ::: "memory");
    }
  }
}

// ===========================================================================
// Covariance (pcavar), src/SphericalBasis.cc:501-509, 525-545, 563-581.
// Per jackknife bin T = indx % sampT the reference accumulates, over the bin's particles,
//   coefT[T][lm](n)   = the bin's share of the coefficients (complex: cos + i sin)
//   coefM[T][lm](n,o) = sum_p mass fac0^2 (K_lm P_lm)^2 potd(l,n) potd(l,o)
// potd is linear in the two table nodes of the particle's radial cell, so coefM factorises
// like the coefficients do: with a = P0 x1, b = P0 x2 and W = mass fac0^2 (K P)^2
//   Saa[i] = sum_{p in cell i} W a^2,  Sab[i] = sum W a b,  Sbb[i] = sum W b^2
//   coefM(n,o) = sum_i E_n(i) [ D_i E_o(i) + O_i E_o(i+1) + O_{i-1} E_o(i-1) ]
//   D_i = Saa[i] + Sbb[i-1],  O_i = Sab[i]
// i.e. one dense  E^T (Tri E)  product per (T, l, m) with K = numr: the DMMA tensor cores
// do it (mma.sync m8n8k4 f64).  The per-particle work is 3 sums per (l,m).
// ===========================================================================

// keys (T, cell): T = host slot % sampT; particles of other levels -> sentinel
__global__ void __launch_bounds__(256)
sph_cov_key_kernel(SphGeom g, const rec4 *__restrict__ posm, const unsigned *__restrict__ level,
                   const uint32_t *__restrict__ slot_of, size_t n, int use_levels, int mlevel,
                   uint32_t sampT, uint64_t indx_base, uint32_t *__restrict__ keys)
{
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    uint32_t key = sampT * (uint32_t)g.numr;
    if (!use_levels || level[i] == (unsigned)mlevel) {
      const uint64_t indx = indx_base + (slot_of ? slot_of[i] : (uint32_t)i);
      key = (uint32_t)(indx % sampT) * (uint32_t)g.numr + sph_cell_of(g, ld_rec(&posm[i]));
    }
    keys[i] = key;
  }
}

// quadratic deposit: lane = particle, 3 sums per (l,m) + the mass in registers, flushed
// (fixed butterfly) when the (T, cell) key changes.  Partial slot layout as the deposit's.
template <int L>
__global__ void __launch_bounds__(kDepositWarps * 32, 2)
sph_cov_deposit_kernel(SphGeom g, const double *__restrict__ p0, const rec4 *__restrict__ ps,
                       const uint32_t *__restrict__ skeys, const uint32_t *__restrict__ seg,
                       int keybase0, int CH, double adb, double *__restrict__ PS)
{
  constexpr int NS = (L + 1) * (L + 2) / 2;
  constexpr int NACC = 3 * NS + 1;
  using R = RS<NACC, 0>;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int keybase = keybase0 + blockIdx.y * g.numr;
  const uint32_t lo = seg[keybase], hi = seg[keybase + g.numr - 1];
  const size_t sbase = (size_t)blockIdx.y * (g.numr + 2) + (lo - seg[keybase0]) / CH;
  const size_t chunk = (size_t)blockIdx.x * kDepositWarps + w;
  const size_t beg = (size_t)lo + chunk * (size_t)CH;
  if (beg >= hi) return;
  const size_t end = beg + CH < hi ? beg + CH : hi;

  double A[NACC];
#pragma unroll
  for (int r = 0; r < NACC; r++) A[r] = 0.0;
  int cur = -1;
  auto flush = [&](int c) {
    const size_t slot = (sbase + chunk + (size_t)c) * NACC;
    R::run(A, lane);
#pragma unroll
    for (int i = 0; i < R::FIN; i++) {
      const int idx = R::row(lane, i);
      if (idx >= 0) PS[slot + idx] = A[i];
    }
#pragma unroll
    for (int r = 0; r < NACC; r++) A[r] = 0.0;
  };

  for (size_t k0 = beg; k0 < end; k0 += 32) {
    const bool valid = k0 + lane < end;
    rec4 pm = rec4{1.0, 0.0, 0.0, 0.0};
    int cell = 0x7fffffff;
    if (valid) { pm = ld_rec(&ps[k0 + lane]); cell = (int)skeys[k0 + lane] - keybase; }
    SphDep d;
    sph_deposit_setup(g, p0, pm, adb, valid ? cell : 0, d);
    // d.w1 = mass adb fac0 P0 x1 = mass adb fac0 a: W a^2 = w1^2 / (mass adb), etc.
    const double mass = pm.w * adb;
    const double im = valid && mass != 0.0 ? 1.0 / mass : 0.0;
    const double waa = valid ? d.w1 * d.w1 * im : 0.0, wab = valid ? d.w1 * d.w2 * im : 0.0,
                 wbb = valid ? d.w2 * d.w2 * im : 0.0;
    int c = __reduce_min_sync(0xffffffffu, cell);
    while (true) {
      if (c != cur) {
        if (cur >= 0) flush(cur);
        cur = c;
      }
      const bool mine = cell == c;
      const double aa = mine ? waa : 0.0, ab = mine ? wab : 0.0, bb = mine ? wbb : 0.0;
      sph_for_each_plm<L, 0, L>(d.costh, c_inv, [](int) {}, [&](int l, int m, double plm, double) {
        const double f = (g.cov_unit_m && m) ? 1.0 : c_K[l * (kMaxL + 1) + m] * plm;
        const double f2 = f * f;
        const int q = 3 * sph_slot(L, l, m);
        A[q] = fma(aa, f2, A[q]);
        A[q + 1] = fma(ab, f2, A[q + 1]);
        A[q + 2] = fma(bb, f2, A[q + 2]);
      }, [](int) {});
      A[3 * NS] += mine ? mass : 0.0;
      const int nc = __reduce_min_sync(0xffffffffu, (cell > c && valid) ? cell : 0x7fffffff);
      if (nc == 0x7fffffff) break;
      c = nc;
    }
  }
  if (cur >= 0) flush(cur);
}

// combine: D[ty][slot][i] = Saa[i] + Sbb[i-1], O[ty][slot][i] = Sab[i]; the mass of cell i
// goes to Mc[ty][i].  Chunks visited in increasing order.
__global__ void sph_cov_combine_kernel(SphGeom g, const uint32_t *__restrict__ seg, int keybase0, int CH,
                                       const double *__restrict__ PS, double *__restrict__ DO,
                                       double *__restrict__ Mc)
{
  const int i = blockIdx.x, ty = blockIdx.y;
  const int ns = g.nslot, nacc = 3 * ns + 1;
  const int keybase = keybase0 + ty * g.numr;
  const uint32_t lo = seg[keybase];
  const size_t sbase = (size_t)ty * (g.numr + 2) + (lo - seg[keybase0]) / CH;
  PS += sbase * nacc;
  double *D = DO + (size_t)ty * ns * g.numr * 2, *O = D + (size_t)ns * g.numr;
  auto cell_sum = [&](int c, int idx) {
    double sum = 0.0;
    if (c < 0 || c > g.numr - 2) return sum;
    uint32_t s0 = seg[keybase + c], s1 = seg[keybase + c + 1];
    if (s1 > s0) {
      size_t j0 = (s0 - lo) / CH, j1 = (s1 - 1 - lo) / CH;
      for (size_t j = j0; j <= j1; j++) sum += PS[(j + c) * nacc + idx];
    }
    return sum;
  };
  for (int slot = threadIdx.x; slot < ns; slot += blockDim.x) {
    D[(size_t)slot * g.numr + i] = cell_sum(i, 3 * slot) + cell_sum(i - 1, 3 * slot + 2);
    O[(size_t)slot * g.numr + i] = cell_sum(i, 3 * slot + 1);
  }
  if (threadIdx.x == blockDim.x - 1) Mc[(size_t)ty * g.numr + i] = cell_sum(i, 3 * ns);
}

__global__ void sph_cov_mass_kernel(SphGeom g, const double *__restrict__ Mc, double *__restrict__ massT)
{
  if (threadIdx.x) return;
  double t = 0.0;
  for (int i = 0; i < g.numr; i++) t += Mc[(size_t)blockIdx.x * g.numr + i];
  massT[blockIdx.x] = t;
}

// contraction on the FP64 tensor cores: one CTA per (slot, T); warp w takes the nodes
// i = 4 (w + 4 s), s = 0, 1, ...; the four partial nmax x nmax tiles are summed in warp order.
// A fragment (row n = 8 rt + lane/4, k = lane%4) is E[l][n][i]; the B fragment of column tile
// ct is the tridiagonally weighted E of the same lane mapping, so one set of loads serves both.
constexpr int kCovTiles = 4;          // nmax <= 32
__global__ void __launch_bounds__(128)
sph_cov_contract_kernel(SphGeom g, const double *__restrict__ Ep, const double *__restrict__ DO,
                        const int *__restrict__ slot_l, const int *__restrict__ slot_m, int M0_only,
                        double *__restrict__ Mout)
{
  __shared__ double part[4][kCovTiles * 8][kCovTiles * 8 + 1];
  const int slot = blockIdx.x, ty = blockIdx.y;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int gq = lane >> 2, t = lane & 3;
  const int N = g.nmax, R = g.numr, NT = (N + 7) / 8;
  const int l = slot_l[slot];
  const double *D = DO + ((size_t)ty * g.nslot * 2 + slot) * R;
  const double *O = D + (size_t)g.nslot * R;
  const double *E = Ep + (size_t)l * N * R;
  double acc[kCovTiles][kCovTiles][2];
#pragma unroll
  for (int a = 0; a < kCovTiles; a++)
#pragma unroll
    for (int b = 0; b < kCovTiles; b++) acc[a][b][0] = acc[a][b][1] = 0.0;

  const bool dead = M0_only && slot_m[slot] != 0;
  if (!dead) {
    for (int i0 = 4 * w; i0 < R; i0 += 16) {
      const int i = i0 + t;
      const bool in = i < R;
      const double Di = in ? D[i] : 0.0;
      const double Oi = (in && i <= R - 2) ? O[i] : 0.0;
      const double Om = (in && i >= 1) ? O[i - 1] : 0.0;
      double af[kCovTiles], bf[kCovTiles];
#pragma unroll
      for (int r = 0; r < kCovTiles; r++) {
        const int n = 8 * r + gq;
        double ec = 0.0, em = 0.0, ep = 0.0;
        if (r < NT && n < N && in) {
          const double *e = E + (size_t)n * R + i;
          ec = e[0];
          if (i >= 1) em = e[-1];
          if (i <= R - 2) ep = e[1];
        }
        af[r] = ec;
        bf[r] = Di * ec + Oi * ep + Om * em;
      }
#pragma unroll
      for (int a = 0; a < kCovTiles; a++)
#pragma unroll
        for (int b = 0; b < kCovTiles; b++)
          if (a < NT && b < NT) dmma_m8n8k4(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
    }
  }
#pragma unroll
  for (int a = 0; a < kCovTiles; a++)
#pragma unroll
    for (int b = 0; b < kCovTiles; b++) {
      part[w][8 * a + gq][8 * b + 2 * t] = acc[a][b][0];
      part[w][8 * a + gq][8 * b + 2 * t + 1] = acc[a][b][1];
    }
  __syncthreads();
  double *M = Mout + ((size_t)ty * g.nslot + slot) * N * N;
  for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
    const int n = e / N, o = e % N;
    M[e] = ((part[0][n][o] + part[1][n][o]) + part[2][n][o]) + part[3][n][o];
  }
}

// host layout: coefT[T][iC][n][2] (iC = l(l+1)/2 + m; re = cos row, im = sin row),
// coefM[T][iC][n][o]
__global__ void sph_cov_pack_kernel(SphGeom g, const double *__restrict__ coefT_rows,
                                    const double *__restrict__ Mslots, double *__restrict__ outT,
                                    double *__restrict__ outM)
{
  const int iC = blockIdx.x, ty = blockIdx.y;
  const int N = g.nmax, ldim = g.nslot;
  int l = 0;
  while ((l + 1) * (l + 2) / 2 <= iC) l++;
  const int m = iC - l * (l + 1) / 2;
  const int row = sph_row(l, m), slot = sph_slot(g.lmax, l, m);
  const double *rows = coefT_rows + (size_t)ty * g.nrow * N;
  double *T = outT + ((size_t)ty * ldim + iC) * N * 2;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    T[2 * n] = rows[(size_t)row * N + n];
    T[2 * n + 1] = m ? rows[(size_t)(row + 1) * N + n] : 0.0;
  }
  const double *Ms = Mslots + ((size_t)ty * ldim + slot) * N * N;
  double *M = outM + ((size_t)ty * ldim + iC) * N * N;
  for (int e = threadIdx.x; e < N * N; e += blockDim.x) M[e] = Ms[e];
}

template <int L> struct LaunchCov {
  static int run(cudaStream_t st, dim3 grid, const SphGeom &g, const double *p0, const rec4 *ps,
                 const uint32_t *skeys, const uint32_t *seg, int keybase, int CH, double adb, double *PS)
  {
    sph_cov_deposit_kernel<L><<<grid, kDepositWarps * 32, 0, st>>>(g, p0, ps, skeys, seg, keybase, CH, adb, PS);
    CU_LAUNCH_CHECK();
    return EXPB200_OK;
  }
};

int launch_cov_deposit(int lmax, cudaStream_t st, dim3 grid, const SphGeom &g, const double *p0,
                       const rec4 *ps, const uint32_t *skeys, const uint32_t *seg, int keybase, int CH,
                       double adb, double *PS)
{
  switch (lmax) {
    case 0: return LaunchCov<0>::run(st, grid, g, p0, ps, skeys, seg, keybase, CH, adb, PS);
    case 1: return LaunchCov<1>::run(st, grid, g, p0, ps, skeys, seg, keybase, CH, adb, PS);
    case 2: return LaunchCov<2>::run(st, grid, g, p0, ps, skeys, seg, keybase, CH, adb, PS);
    case 3: return LaunchCov<3>::run(st, grid, g, p0, ps, skeys, seg, keybase, CH, adb, PS);
    case 4: return LaunchCov<4>::run(st, grid, g, p0, ps, skeys, seg, keybase, CH, adb, PS);
    case 5: return LaunchCov<5>::run(st, grid, g, p0, ps, skeys, seg, keybase, CH, adb, PS);
    case 6: return LaunchCov<6>::run(st, grid, g, p0, ps, skeys, seg, keybase, CH, adb, PS);
    default: set_error("covariance: lmax=%d not instantiated (max 6)", lmax); return EXPB200_EINVAL;
  }
}

template <int L> struct Launch {
  static int deposit(cudaStream_t st, unsigned nchunks, const SphGeom &g, const double *p0,
                     const rec4 *ps, const uint32_t *skeys, const uint32_t *seg, int keybase,
                     int CH, double adb, double *PA, double *PB, int nbatch, DepositGather G)
  {
    const dim3 nblocks(div_up(nchunks, kDepositWarps), nbatch > 0 ? nbatch : 1);
    const int batched = nbatch > 0;
    bool gathered = false;
    if constexpr (L <= 4) {
      sph_deposit_reg_kernel<L, 0, L, 3><<<nblocks, kDepositWarps * 32, 0, st>>>(g, p0, ps, skeys, seg,
                                                                                 keybase, CH, adb, batched, PA, PB);
    } else if constexpr (L <= kRegDepositMaxL) {
      static const int variant = getenv("EXPB200_DEPOSIT_VARIANT") ? atoi(getenv("EXPB200_DEPOSIT_VARIANT")) : 1;
#define DEP(M0, M1, MB) sph_deposit_reg_kernel<L, M0, M1, MB><<<nblocks, kDepositWarps * 32, 0, st>>>(g, p0, ps, skeys, seg, keybase, CH, adb, batched, PA, PB)
      if (variant == 0) { DEP(0, 2, 3); DEP(3, L, 3); }
      else if (variant == 1 && G.perm && !batched) {
        // first sweep materialises the sorted stream (G.out == ps), the second streams it
        sph_deposit_reg_kernel<L, 0, 2, 2, true><<<nblocks, kDepositWarps * 32, 0, st>>>(
            g, p0, ps, skeys, seg, keybase, CH, adb, batched, PA, PB, G);
        gathered = true;
        DEP(3, L, 3);
      }
      else if (variant == 1) { DEP(0, 2, 2); DEP(3, L, 3); }
      else if (variant == 2) { DEP(0, 1, 3); DEP(2, 3, 3); DEP(4, L, 3); }
      else { DEP(0, 1, 4); DEP(2, 3, 4); DEP(4, L, 4); }
#undef DEP
    } else {
      size_t sm = kDepositWarps * DepositShape<L>::WARP_DOUBLES * sizeof(double);
      static bool attr[64] = {};
      int dev = 0;
      cudaGetDevice(&dev);
      if (!attr[dev & 63]) {
        CU_TRY(cudaFuncSetAttribute(sph_deposit_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        attr[dev & 63] = true;
      }
      sph_deposit_kernel<L><<<nblocks, kDepositWarps * 32, sm, st>>>(g, p0, ps, skeys, seg, keybase,
                                                                      CH, adb, batched, PA, PB);
    }
    CU_LAUNCH_CHECK();
    if (G.perm && !gathered) { set_error("deposit: gather requested on a path without it"); return EXPB200_ESTATE; }
    return EXPB200_OK;
  }
  template <int TH, int MINB, int ROWS, int NST>
  static int force_pipelined(cudaStream_t st, const SphGeom &g, const double2 *T, const rec4 *ps,
                             const ForceOut &F, const uint32_t *perm, const uint32_t *skeys,
                             const uint32_t *seg, int keybase, size_t n, int2 *tdesc)
  {
    constexpr size_t sm = sizeof(FPShared<L, TH, ROWS, NST>);
    static int ctas[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    int &occ = ctas[dev & 63];
    if (!occ) {
      CU_TRY(cudaFuncSetAttribute(sph_force2p_kernel<L, TH, MINB, ROWS, NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      int nb = 0, nsm = 0;
      CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, sph_force2p_kernel<L, TH, MINB, ROWS, NST>, TH, sm));
      CU_TRY(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
      occ = (nb > 0 ? nb : 1) * nsm;
    }
    const size_t ntiles = div_up(n, (size_t)(2 * TH));
    const unsigned grid = (unsigned)(ntiles < (size_t)occ ? ntiles : (size_t)occ);
    sph_tile_desc_kernel<<<div_up(ntiles, 256), 256, 0, st>>>(skeys, seg, keybase, n, 2 * TH, g.numr, ROWS, tdesc);
    sph_force2p_kernel<L, TH, MINB, ROWS, NST><<<grid, TH, sm, st>>>(g, T, ps, F, perm, tdesc, seg, keybase, n);
    CU_LAUNCH_CHECK();
    return EXPB200_OK;
  }
  template <int ROWS, int REGE>
  static int force_ws(cudaStream_t st, const SphGeom &g, const double2 *T, const rec4 *ps,
                      const ForceOut &F, const uint32_t *perm, const uint32_t *skeys,
                      const uint32_t *seg, int keybase, size_t n, int2 *tdesc)
  {
    auto kern = sph_force_ws_kernel<L, ROWS, REGE, 256 - REGE>;
    constexpr size_t sm = sizeof(WSShared<L, ROWS>);
    static int ctas[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    int &occ = ctas[dev & 63];
    if (!occ) {
      CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      int nb = 0, nsm = 0;
      CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kWSThreads, sm));
      CU_TRY(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
      occ = (nb > 0 ? nb : 1) * nsm;
    }
    const size_t ntiles = div_up(n, (size_t)kWSTile);
    const unsigned grid = (unsigned)(ntiles < (size_t)occ ? ntiles : (size_t)occ);
    sph_tile_desc_kernel<<<div_up(ntiles, 256), 256, 0, st>>>(skeys, seg, keybase, n, kWSTile, g.numr, ROWS, tdesc);
    kern<<<grid, kWSThreads, sm, st>>>(g, T, ps, F, perm, tdesc, seg, keybase, n);
    CU_LAUNCH_CHECK();
    return EXPB200_OK;
  }
  static int force(cudaStream_t st, const SphGeom &g, const double2 *T, const rec4 *ps, const ForceOut &F,
                   const uint32_t *perm, const uint32_t *skeys, const uint32_t *seg, int keybase, size_t n,
                   int2 *tdesc)
  {
    // 2: persistent pipelined kernel (default), 3: warp-specialised persistent kernel (measured
    // 0.70 ms against 0.69 ms at 1e7 particles: not faster),
    // 1: one-shot two-particle kernel, 0: one particle per lane
    static const int variant = getenv("EXPB200_FORCE_VARIANT") ? atoi(getenv("EXPB200_FORCE_VARIANT")) : 2;
    if constexpr (L <= 8) {
      // the pipelined kernels read the sorted keys while (mode 2) writing the next step's keys:
      // they must be different buffers
      if (variant >= 3 && skeys && tdesc && !(F.mode == 2 && skeys == F.keys)) {
        static const int split = getenv("EXPB200_WS_REGS") ? atoi(getenv("EXPB200_WS_REGS")) : 0;
        if (split == 1) return force_ws<4, 160>(st, g, T, ps, F, perm, skeys, seg, keybase, n, tdesc);
        if (split == 2) return force_ws<4, 152>(st, g, T, ps, F, perm, skeys, seg, keybase, n, tdesc);
        if (split == 3) return force_ws<4, 144>(st, g, T, ps, F, perm, skeys, seg, keybase, n, tdesc);
        if (split == 4) return force_ws<4, 136>(st, g, T, ps, F, perm, skeys, seg, keybase, n, tdesc);
        if (split == 5) return force_ws<4, 128>(st, g, T, ps, F, perm, skeys, seg, keybase, n, tdesc);
        return force_ws<4, 168>(st, g, T, ps, F, perm, skeys, seg, keybase, n, tdesc);
      }
      if (variant >= 2 && skeys && tdesc && !(F.mode == 2 && skeys == F.keys)) {
        // CTAs of two warps (tiles of 128 particles), 6 per SM, 4 table rows per stage, two stages:
        // the fastest of the shapes measured at 1e7 particles (128x3, 192x2, 384x1 were slower)
        return force_pipelined<64, 6, 4, 2>(st, g, T, ps, F, perm, skeys, seg, keybase, n, tdesc);
      }
    }
    if (variant >= 1 && L <= 8)
      sph_force2_kernel<L><<<div_up(n, 2 * kForce2Threads), kForce2Threads, 0, st>>>(g, T, ps, F, perm, seg,
                                                                                   keybase, n);
    else
      sph_force_kernel<L><<<div_up(n, kForceThreads), kForceThreads, 0, st>>>(g, T, ps, F, perm, seg,
                                                                             keybase, n);
    CU_LAUNCH_CHECK();
    return EXPB200_OK;
  }
};

#define DISPATCH_L(lmax, CALL)                                                \
  switch (lmax) {                                                             \
    case 0: return Launch<0>::CALL;   case 1: return Launch<1>::CALL;         \
    case 2: return Launch<2>::CALL;   case 3: return Launch<3>::CALL;         \
    case 4: return Launch<4>::CALL;   case 5: return Launch<5>::CALL;         \
    case 6: return Launch<6>::CALL;   case 7: return Launch<7>::CALL;         \
    case 8: return Launch<8>::CALL;   case 9: return Launch<9>::CALL;         \
    case 10: return Launch<10>::CALL; case 11: return Launch<11>::CALL;       \
    case 12: return Launch<12>::CALL;                                         \
    default: set_error("lmax=%d not instantiated (max %d)", lmax, kMaxL); return EXPB200_EINVAL; \
  }

int launch_deposit(int lmax, cudaStream_t st, unsigned nchunks, const SphGeom &g, const double *p0,
                   const rec4 *ps, const uint32_t *skeys, const uint32_t *seg, int keybase, int CH,
                   double adb, double *PA, double *PB, int nbatch = 0, DepositGather G = DepositGather())
{
  DISPATCH_L(lmax, deposit(st, nchunks, g, p0, ps, skeys, seg, keybase, CH, adb, PA, PB, nbatch, G))
}

// true where launch_deposit's first sweep can build the sorted position stream itself
bool deposit_can_gather(int lmax)
{
  static const int variant = getenv("EXPB200_DEPOSIT_VARIANT") ? atoi(getenv("EXPB200_DEPOSIT_VARIANT")) : 1;
  static const bool off = getenv("EXPB200_NO_DEPOSIT_GATHER") != nullptr;
  return !off && variant == 1 && lmax > 4 && lmax <= kRegDepositMaxL;
}

int launch_force(int lmax, cudaStream_t st, const SphGeom &g, const double2 *T, const rec4 *ps,
                 const ForceOut &F, const uint32_t *perm, const uint32_t *skeys, const uint32_t *seg,
                 int keybase, size_t n, int2 *tdesc)
{
  DISPATCH_L(lmax, force(st, g, T, ps, F, perm, skeys, seg, keybase, n, tdesc))
}

} // namespace
} // namespace expb200

using namespace expb200;

struct expb200_sphere_s {
  int device = 0;
  expb200_sphere_config cfg;
  SphGeom geom;              // centre fields filled per call from the component
  const expb200_component_s *owner = nullptr;   // PotAccel::component (expb200_sphere_set_owner)
  int nlev = 1;
  size_t ncoef = 0;

  double *Ep = nullptr;      // [l][n][i] = ef / sqrt(ev)
  double *p0 = nullptr;      // [numr]
  double *coef = nullptr;    // device expcoef0: last accumulated vector
  double *coef_use = nullptr;// coefficients installed for the force
  double2 *T = nullptr;      // field table [numr][nslot][kSphPoly] (sph_poly_entry)
  double *G = nullptr;       // node sums [nrow][numr]
  double *PA = nullptr, *PB = nullptr;
  size_t part_slots = 0;
  void *cov_buf[11] = {};    // cached scratch of the covariance call
  size_t cov_cap[11] = {};
  int2 *tdesc = nullptr;     // per-tile row ranges of the pipelined force kernel
  size_t tdesc_cap = 0;
  int *row_l = nullptr, *row_m = nullptr;
  int *slot_l = nullptr, *slot_m = nullptr, *slot_rowc = nullptr, *slot_rows = nullptr;
  bool field_valid = false;
  bool have_accum = false;
  uint32_t *used_dev = nullptr;
  cudaStream_t last_stream = nullptr;
  cudaStream_t own_stream = nullptr;   // used when no live component stream is known
  // the stream this sphere's table operations are ordered on
  cudaStream_t work_stream()
  {
    if (last_stream && expb200::stream_is_live(last_stream)) return last_stream;
    last_stream = nullptr;
    return own_stream;
  }
  // multistep: expcoefN / expcoefL per level and the level-switch differences
  double *coefN = nullptr, *coefL = nullptr, *differ = nullptr, *tmp = nullptr;   // [nlev][ncoef]
};

namespace {

int deposit_chunk(size_t n)
{
  // one warp per chunk; 12 warps are resident per SM (register-limited), aim at >= 4 waves
  // of chunks and at most 4096 particles each
  size_t q = n / ((size_t)kSMsB200 * 12 * 4 * 32);
  if (q < 1) q = 1;
  if (q > 128) q = 128;
  return (int)q * 32;
}

SphGeom geom_for(const expb200_sphere_s *s, const expb200_component_s *c)
{
  SphGeom g = s->geom;
  // expansion origin: the centre of the component that OWNS this force (PotAccel::component);
  // freeze: the centre and rtrunc of the component whose particles are evaluated (cC)
  const expb200_component_s *o = s->owner ? s->owner : c;
  g.cx = o->center[0]; g.cy = o->center[1]; g.cz = o->center[2];
  g.fx = c->center[0]; g.fy = c->center[1]; g.fz = c->center[2];
  g.ext = (g.cx != g.fx || g.cy != g.fy || g.cz != g.fz) ? 1 : 0;
  g.rtrunc2 = c->rtrunc * c->rtrunc;
  return g;
}

// Bins the keys in c->keys[0]: permutation, cell offsets and the sorted copy of posm.
// With allow_reorder the particle store is physically re-ordered into this order when it has
// never been (random order: every access through the permutation would be a random
// 32-byte DRAM transaction) or every kReorderEvery sorts (particles drift across cells).
// sorts between two physical re-orderings of the store (EXPB200_REORDER_EVERY overrides the
// single-level default for experiments)
int reorder_every(int nlev)
{
  static const int env = getenv("EXPB200_REORDER_EVERY") ? atoi(getenv("EXPB200_REORDER_EVERY")) : 0;
  static const int envm = getenv("EXPB200_REORDER_EVERY_MULTI") ? atoi(getenv("EXPB200_REORDER_EVERY_MULTI")) : 0;
  if (nlev > 1) return envm > 0 ? envm : kReorderEveryMulti;
  return env > 0 ? env : kReorderEvery;
}

int sort_keyed(expb200_sphere_s *s, expb200_component_s *c, uint32_t nkeys, bool allow_reorder,
               bool defer_stream = false)
{
  // defer_stream: leave the sorted position stream to the first deposit sweep (c->psort_pending),
  // unless this sort is due to re-order the store, which takes the stream as the new store
  const bool reorder_due = allow_reorder && (c->never_reordered ||
      c->sorts_since_reorder + 1 >= reorder_every(s->nlev));
  if (reorder_due) defer_stream = false;
  int cur = 0;
  {
    ProfRegion pr(K_SORT, c->stream, radix_sort_launches(c->n, nkeys, true, !defer_stream));
    SortPayload pl;
    pl.src = defer_stream ? nullptr : c->posm;
    pl.dst = defer_stream ? nullptr : c->posm_sorted;
    int rc = radix_sort_pairs(c->stream, c->keys, c->perm, c->hist, c->rank, c->n, nkeys, &cur, c->seg, pl);
    if (rc) return rc;
  }
  c->cur = cur;
  c->psort = c->posm_sorted;
  c->psort_pending = defer_stream;
  c->sorts_since_reorder++;
  if (allow_reorder && (c->never_reordered || c->sorts_since_reorder >= reorder_every(s->nlev)))
    return component_reorder(c);
  return EXPB200_OK;
}

void mark_sorted(expb200_sphere_s *s, expb200_component_s *c)
{
  c->sort_kind = 2;
  c->sort_owner = s;
  c->keys_per_level = s->cfg.numr;
  c->sort_pos_ver = c->pos_ver;
  c->sort_lev_ver = c->lev_ver;
  c->sort_from = 0;
  const expb200_component_s *o = s->owner ? s->owner : c;
  for (int k = 0; k < 3; k++) c->sort_origin[k] = o->center[k];
}

int ensure_seg(expb200_component_s *c, uint32_t nkeys)
{
  if (c->seg_cap < (size_t)nkeys + 2) {
    if (c->seg) cudaFree(c->seg);
    c->seg = nullptr;
    CU_TRY(cudaMalloc(&c->seg, ((size_t)nkeys + 2) * sizeof(uint32_t)));
    c->seg_cap = (size_t)nkeys + 2;
  }
  return EXPB200_OK;
}

// (level, radial cell) sort of component c for sphere s with the cell order current for the levels
// >= mlevel, reused while neither positions nor levels changed.  Multistep: when the level partition
// of the permutation is known (a full sort established it and only levels >= some lo have changed
// level since), only the tail of the levels >= min(mlevel, lo) is re-binned -- on a sub-step only
// the active levels have moved (src/step.cc:210-262) -- instead of the whole component.
int ensure_sorted(expb200_sphere_s *s, expb200_component_s *c, int mlevel = 0)
{
  SphGeom g = geom_for(s, c);
  if (s->nlev == 1) mlevel = 0;
  if (c->sort_kind == 2 && c->sort_owner == s && c->sort_pos_ver == c->pos_ver &&
      c->sort_lev_ver == c->lev_ver && c->sort_origin[0] == g.cx && c->sort_origin[1] == g.cy &&
      c->sort_origin[2] == g.cz && c->sort_from <= mlevel)
    return EXPB200_OK;
  const uint32_t numr = (uint32_t)s->cfg.numr;
  const uint32_t nkeys = (uint32_t)s->nlev * numr;
  { int rc = ensure_seg(c, nkeys); if (rc) return rc; }
  int m = 0;
  if (s->nlev > 1 && c->n && c->part_ok && c->part_nlev == s->nlev)
    m = mlevel < c->bound_ok_upto ? mlevel : c->bound_ok_upto;
  if (m > 0) {
    const size_t off = c->lev_start[m], nsub = c->n - off;
    const uint32_t nks = (uint32_t)(s->nlev - m) * numr;
    uint32_t *seg_sub = c->seg + (size_t)m * numr;
    if (nsub) {
      const int W = suffix_key_buffer(c->cur, nks);
      { ProfRegion pr(K_KEY, c->stream);
        sph_key_sub_kernel<<<grid_for(nsub, 256), 256, 0, c->stream>>>(g, c->posm, c->level, c->perm[c->cur] + off,
                                                                       nsub, m, s->nlev - 1, c->keys[W] + off);
      }
      CU_LAUNCH_CHECK();
      ProfRegion pr(K_SORT, c->stream, radix_sort_launches(nsub, nks, false, true) + 2);
      SortPayload pl;
      pl.src = c->posm;
      pl.dst = c->posm_sorted;
      int rc = radix_sort_suffix(c->stream, c->keys, c->perm, c->cur, c->hist, c->rank, off, c->n, nks, seg_sub, pl,
                                 (uint32_t)m * numr);
      if (rc) return rc;
      c->psort = c->posm_sorted;         // [0, off) of the stream is stale: levels < m are not read
    } else {
      int rc = segment_offsets(c->stream, c->keys[c->cur], 0, seg_sub, nks, (uint32_t)off);
      if (rc) return rc;
    }
    int rc = component_read_bounds(c, numr, s->nlev, m);
    if (rc) return rc;
    mark_sorted(s, c);
    c->sort_from = m;
    c->sort_counters[1]++; c->sort_counters[2] += nsub;
    c->sorts_since_reorder++;
    return EXPB200_OK;
  }
  if (c->n) {
    { ProfRegion pr(K_KEY, c->stream);
    sph_key_kernel<<<grid_for(c->n, 256), 256, 0, c->stream>>>(g, c->posm, c->level, c->n,
                                                               s->nlev > 1, s->nlev - 1, c->keys[0]);
    }
    CU_LAUNCH_CHECK();
    int rc = sort_keyed(s, c, nkeys, true);
    if (rc) return rc;
    c->sort_counters[0]++; c->sort_counters[3] += c->n;
    if (s->nlev > 1 && (rc = component_read_bounds(c, numr, s->nlev, -1))) return rc;
  } else {
    CU_TRY(cudaMemsetAsync(c->seg, 0, ((size_t)nkeys + 1) * sizeof(uint32_t), c->stream));
  }
  mark_sorted(s, c);
  return EXPB200_OK;
}

int ensure_partials(expb200_sphere_s *s, size_t n, int CH)
{
  size_t slots = (n + CH - 1) / CH + (size_t)s->cfg.numr + 1;
  if (slots <= s->part_slots) return EXPB200_OK;
  if (s->PA) cudaFree(s->PA);
  if (s->PB) cudaFree(s->PB);
  s->PA = s->PB = nullptr;
  s->part_slots = 0;
  CU_TRY(cudaMalloc(&s->PA, slots * s->geom.nrow * sizeof(double)));
  CU_TRY(cudaMalloc(&s->PB, slots * s->geom.nrow * sizeof(double)));
  s->part_slots = slots;
  return EXPB200_OK;
}

int ensure_tdesc(expb200_sphere_s *s, size_t n)
{
  const size_t need = n / 128 + 2;          // tiles are >= 128 particles
  if (need <= s->tdesc_cap) return EXPB200_OK;
  if (s->tdesc) cudaFree(s->tdesc);
  s->tdesc = nullptr;
  s->tdesc_cap = 0;
  CU_TRY(cudaMalloc(&s->tdesc, need * sizeof(int2)));
  s->tdesc_cap = need;
  return EXPB200_OK;
}

int build_field(expb200_sphere_s *s, cudaStream_t st)
{
  const SphGeom &g = s->geom;
  dim3 grid(div_up(g.numr, 128), g.nslot);
  ProfRegion pr(K_FIELD, st);
  sph_field_kernel<<<grid, 128, 0, st>>>(g, s->Ep, s->p0, s->coef_use, s->slot_l, s->slot_m,
                                         s->slot_rowc, s->slot_rows, s->T);
  CU_LAUNCH_CHECK();
  s->field_valid = true;
  return EXPB200_OK;
}

// deposit + combine + contract of levlist[mlevel] on an already sorted component -> s->coef
int accumulate_sorted(expb200_sphere_s *s, expb200_component_s *c, int mlevel, double adb)
{
  const int CH = deposit_chunk(c->n);
  int rc = ensure_partials(s, c->n, CH);
  if (rc) return rc;
  SphGeom g = geom_for(s, c);
  const int keybase = (s->nlev > 1 ? mlevel : 0) * s->cfg.numr;
  if (c->n) {
    unsigned nchunks = div_up(c->n, CH);
    ProfRegion pr(K_DEPOSIT, c->stream);
    DepositGather G;
    if (c->psort_pending) { G.perm = c->perm[c->cur]; G.store = c->posm; G.out = c->posm_sorted; }
    rc = launch_deposit(s->cfg.lmax, c->stream, nchunks, g, s->p0, c->psort, c->keys[c->cur],
                        c->seg, keybase, CH, adb, s->PA, s->PB, 0, G);
    if (rc) return rc;
    c->psort_pending = false;
  }
  { ProfRegion pr(K_COMBINE, c->stream);
    sph_combine_kernel<<<g.numr, 64, 0, c->stream>>>(g, c->seg, keybase, CH, 0, s->PA, s->PB, s->G); }
  { ProfRegion pr(K_CONTRACT, c->stream);
    sph_contract_kernel<<<dim3(g.nrow, g.nmax), 256, 0, c->stream>>>(g, s->Ep, s->G, s->row_l, s->row_m,
                                                                     s->cfg.M0_only, s->coef); }
  CU_LAUNCH_CHECK();
  s->have_accum = true;
  s->last_stream = c->stream;
  return EXPB200_OK;
}

int accumulate(expb200_sphere_s *s, expb200_component_s *c, int mlevel, double adb)
{
  { cudaStream_t ws = s->work_stream(); if (ws != c->stream) CU_TRY(cudaStreamSynchronize(ws)); }
  int rc = ensure_sorted(s, c, mlevel);
  if (rc) return rc;
  return accumulate_sorted(s, c, mlevel, adb);
}

} // namespace

extern "C" {

int expb200_sphere_create(int device, const expb200_sphere_config *cfg, const double *ev,
                          const double *ef, const double *p0, expb200_sphere *out)
{
  REQUIRE(cfg && ev && ef && p0 && out, "NULL argument");
  REQUIRE(cfg->lmax >= 0 && cfg->lmax <= kMaxL, "lmax=%d outside [0,%d]", cfg->lmax, kMaxL);
  REQUIRE(cfg->nmax >= 1 && cfg->nmax <= 1024, "nmax=%d out of range", cfg->nmax);
  REQUIRE(cfg->numr >= 4, "numr=%d too small", cfg->numr);
  REQUIRE(cfg->cmap >= 0 && cfg->cmap <= 2, "cmap=%d unknown", cfg->cmap);
  REQUIRE(cfg->multistep >= 0 && cfg->multistep < 30, "multistep=%d out of range", cfg->multistep);
  REQUIRE(cfg->scale > 0.0, "scale must be positive");
  int ndev = 0;
  CU_TRY(cudaGetDeviceCount(&ndev));
  REQUIRE(device >= 0 && device < ndev, "device %d not present (%d visible)", device, ndev);
  DeviceGuard gd(device);

  auto *s = new expb200_sphere_s;
  s->device = device;
  s->cfg = *cfg;
  s->nlev = cfg->multistep + 1;
  const int L = cfg->lmax, N = cfg->nmax, R = cfg->numr;
  s->geom = make_sph_geom(*cfg);
  SphGeom &g = s->geom;
  s->ncoef = (size_t)g.nrow * N;

  // constants for this device
  {
    std::vector<double> K, inv;
    make_sph_constants(K, inv);
    CU_TRY(cudaMemcpyToSymbol(c_K, K.data(), K.size() * sizeof(double)));
    CU_TRY(cudaMemcpyToSymbol(c_inv, inv.data(), inv.size() * sizeof(double)));
  }

  // tables: fold 1/sqrt(ev) into ef
  {
    std::vector<double> E((size_t)(L + 1) * N * R);
    for (int l = 0; l <= L; l++)
      for (int n = 0; n < N; n++) {
        double q = std::sqrt(ev[l * N + n]);
        const double *src = ef + ((size_t)l * N + n) * R;
        double *dst = E.data() + ((size_t)l * N + n) * R;
        for (int i = 0; i < R; i++) dst[i] = src[i] / q;
      }
    CU_TRY(cudaMalloc(&s->Ep, E.size() * sizeof(double)));
    CU_TRY(cudaMemcpy(s->Ep, E.data(), E.size() * sizeof(double), cudaMemcpyHostToDevice));
    CU_TRY(cudaMalloc(&s->p0, R * sizeof(double)));
    CU_TRY(cudaMemcpy(s->p0, p0, R * sizeof(double), cudaMemcpyHostToDevice));
  }

  // row -> (l,m) and the force loop's slot -> coefficient rows map (sph_host.h)
  {
    SphMaps M = make_sph_maps(*cfg);
    auto up = [&](int **d, const std::vector<int> &h) -> int {
      CU_TRY(cudaMalloc(d, h.size() * sizeof(int)));
      CU_TRY(cudaMemcpy(*d, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice));
      return 0;
    };
    int rc;
    if ((rc = up(&s->row_l, M.row_l)) || (rc = up(&s->row_m, M.row_m)) || (rc = up(&s->slot_l, M.slot_l)) ||
        (rc = up(&s->slot_m, M.slot_m)) ||
        (rc = up(&s->slot_rowc, M.rowc)) || (rc = up(&s->slot_rows, M.rows)))
      return rc;
  }

  CU_TRY(cudaMalloc(&s->coef, s->ncoef * sizeof(double)));
  CU_TRY(cudaMalloc(&s->coef_use, s->ncoef * sizeof(double)));
  CU_TRY(cudaMemset(s->coef, 0, s->ncoef * sizeof(double)));
  CU_TRY(cudaMemset(s->coef_use, 0, s->ncoef * sizeof(double)));
  CU_TRY(cudaMalloc(&s->T, (size_t)R * g.nslot * kSphPoly * sizeof(double2)));
  CU_TRY(cudaMalloc(&s->G, (size_t)g.nrow * R * sizeof(double)));
  CU_TRY(cudaMalloc(&s->used_dev, 2 * sizeof(uint32_t)));
  CU_TRY(cudaStreamCreateWithFlags(&s->own_stream, cudaStreamNonBlocking));
  {
    size_t bytes = (size_t)s->nlev * s->ncoef * sizeof(double);
    CU_TRY(cudaMalloc(&s->coefN, bytes));
    CU_TRY(cudaMalloc(&s->coefL, bytes));
    CU_TRY(cudaMalloc(&s->differ, bytes));
    CU_TRY(cudaMalloc(&s->tmp, s->ncoef * sizeof(double)));
    CU_TRY(cudaMemset(s->coefN, 0, bytes));
    CU_TRY(cudaMemset(s->coefL, 0, bytes));
    CU_TRY(cudaMemset(s->differ, 0, bytes));
  }
  *out = s;
  return EXPB200_OK;
}

int expb200_sphere_set_owner(expb200_sphere s, expb200_component owner)
{
  REQUIRE(s, "sphere is NULL");
  s->owner = owner;
  return EXPB200_OK;
}

int expb200_sphere_destroy(expb200_sphere s)
{
  if (!s) return EXPB200_OK;
  DeviceGuard g(s->device);
  cudaDeviceSynchronize();
  cudaFree(s->Ep); cudaFree(s->p0); cudaFree(s->coef); cudaFree(s->coef_use); cudaFree(s->T);
  cudaFree(s->G); cudaFree(s->PA); cudaFree(s->PB); cudaFree(s->row_l); cudaFree(s->row_m);
  cudaFree(s->slot_l); cudaFree(s->slot_m); cudaFree(s->slot_rowc); cudaFree(s->slot_rows); cudaFree(s->used_dev);
  cudaFree(s->coefN); cudaFree(s->coefL); cudaFree(s->differ); cudaFree(s->tmp); cudaFree(s->tdesc);
  for (int q = 0; q < 11; q++) cudaFree(s->cov_buf[q]);
  if (s->own_stream) cudaStreamDestroy(s->own_stream);
  delete s;
  return EXPB200_OK;
}

int expb200_sphere_determine_coefficients(expb200_sphere s, expb200_component c, int mlevel,
                                          double adb, double *coef_host, uint64_t *used)
{
  REQUIRE(s && c, "NULL argument");
  REQUIRE(s->device == c->device, "sphere on device %d, component on device %d", s->device, c->device);
  REQUIRE(mlevel >= 0 && mlevel < s->nlev, "mlevel=%d outside [0,%d]", mlevel, s->nlev - 1);
  DeviceGuard g(s->device);
  int rc = accumulate(s, c, mlevel, adb);
  if (rc) return rc;
  if (coef_host)
    CU_TRY(cudaMemcpyAsync(coef_host, s->coef, s->ncoef * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (used) {
    const int keybase = (s->nlev > 1 ? mlevel : 0) * s->cfg.numr;
    uint32_t lohi[2];
    CU_TRY(cudaMemcpyAsync(&lohi[0], c->seg + keybase, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaMemcpyAsync(&lohi[1], c->seg + keybase + s->cfg.numr - 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    *used = lohi[1] - lohi[0];
  } else if (coef_host) {
    CU_TRY(cudaStreamSynchronize(c->stream));
  }
  return EXPB200_OK;
}

int expb200_sphere_coef_device(expb200_sphere s, double **dptr, void **stream)
{
  REQUIRE(s && dptr, "NULL argument");
  *dptr = s->coef;
  if (stream) *stream = s->work_stream();
  return EXPB200_OK;
}

int expb200_sphere_set_coefficients(expb200_sphere s, const double *coef_host)
{
  REQUIRE(s, "sphere is NULL");
  DeviceGuard g(s->device);
  cudaStream_t st = s->work_stream();
  if (coef_host) {
    CU_TRY(cudaStreamSynchronize(st));
    CU_TRY(cudaMemcpy(s->coef_use, coef_host, s->ncoef * sizeof(double), cudaMemcpyHostToDevice));
  } else {
    REQUIRE(s->have_accum, "no coefficients accumulated on the device yet");
    CU_TRY(cudaMemcpyAsync(s->coef_use, s->coef, s->ncoef * sizeof(double), cudaMemcpyDeviceToDevice, st));
  }
  s->field_valid = false;
  return EXPB200_OK;
}

int expb200_sphere_determine_acceleration(expb200_sphere s, expb200_component c, int mlevel)
{
  REQUIRE(s && c, "NULL argument");
  REQUIRE(s->device == c->device, "sphere on device %d, component on device %d", s->device, c->device);
  REQUIRE(mlevel >= 0 && mlevel < s->nlev, "mlevel=%d outside [0,%d]", mlevel, s->nlev - 1);
  DeviceGuard g(s->device);
  if (c->n == 0) return EXPB200_OK;
  int rc = ensure_sorted(s, c, mlevel);
  if (rc) return rc;
  { cudaStream_t ws = s->work_stream(); if (ws != c->stream) CU_TRY(cudaStreamSynchronize(ws)); }
  if (!s->field_valid) {
    rc = build_field(s, c->stream);
    if (rc) return rc;
  }
  s->last_stream = c->stream;
  SphGeom gm = geom_for(s, c);
  const int keybase = (s->nlev > 1 ? mlevel : 0) * s->cfg.numr;
  ProfRegion pr(K_FORCE, c->stream, 2);   // tile pre-pass + force
  ForceOut F{c->accp, c->vel, c->posm, c->keys[0], 0.0, 0.0, 0};
  if ((rc = ensure_tdesc(s, c->n))) return rc;
  return launch_force(s->cfg.lmax, c->stream, gm, s->T, c->psort, F, c->perm[c->cur], c->keys[c->cur],
                      c->seg, keybase, c->n, s->tdesc);
}

// ---- multistep coefficient tables -------------------------------------------------
// ~ the bookkeeping of determine_coefficients_particles (src/SphericalBasis.cc:785-788,
// 864-903): swap expcoefL[mlevel] <-> expcoefN[mlevel], then expcoefN[mlevel] = the
// (all-reduced) vector just accumulated.
int expb200_sphere_multistep_commit(expb200_sphere s, int mlevel)
{
  REQUIRE(s, "sphere is NULL");
  REQUIRE(mlevel >= 0 && mlevel < s->nlev, "mlevel=%d outside [0,%d]", mlevel, s->nlev - 1);
  REQUIRE(s->have_accum, "no coefficients accumulated yet");
  DeviceGuard g(s->device);
  cudaStream_t st = s->work_stream();
  double *N = s->coefN + (size_t)mlevel * s->ncoef, *L = s->coefL + (size_t)mlevel * s->ncoef;
  CU_TRY(cudaMemcpyAsync(L, N, s->ncoef * sizeof(double), cudaMemcpyDeviceToDevice, st));
  CU_TRY(cudaMemcpyAsync(N, s->coef, s->ncoef * sizeof(double), cudaMemcpyDeviceToDevice, st));
  return EXPB200_OK;
}

// ~ SphericalBasis::compute_multistep_coefficients (src/SphericalBasis.cc:1231-1333) with
// the interpolation tables of initialize_multistep (src/multistep.cc:656-671); installs
// the result as the coefficients used by the next force evaluation.
int expb200_sphere_compute_multistep_coefficients(expb200_sphere s, int mdrft)
{
  REQUIRE(s, "sphere is NULL");
  const int multistep = s->nlev - 1, Mstep = 1 << multistep;
  REQUIRE(mdrft >= 0 && mdrft <= Mstep, "mdrft=%d outside [0,%d]", mdrft, Mstep);
  DeviceGuard g(s->device);
  InterpArgs A;
  A.nlev = s->nlev;
  // mfirst[mdrft]: lowest active level (src/multistep.cc:646-654)
  A.mfirst = 0;
  for (int lev = 0; lev <= multistep; lev++)
    if (mdrft == 0 || mdrft % (1 << (multistep - lev)) == 0) { A.mfirst = lev; break; }
  for (int M = 0; M < A.mfirst; M++) {
    int d = 1 << (multistep - M);                 // dstep for level M
    int dL = (mdrft / d) * d, dN = dL + d;        // dstepL[M][mdrft], dstepN[M][mdrft]
    double b = (double)(mdrft - dL) / (double)(dN - dL);
    A.a[M] = 1.0 - b;
    A.b[M] = b;
  }
  cudaStream_t st = s->work_stream();
  multistep_interp_kernel<<<div_up(s->ncoef, 256), 256, 0, st>>>(s->coefL, s->coefN, A, s->ncoef, s->coef_use);
  CU_LAUNCH_CHECK();
  s->field_valid = false;
  return EXPB200_OK;
}

// Host access to the per-level tables (which = 0: expcoefN, 1: expcoefL, 2: differ).
int expb200_sphere_get_level_coefficients(expb200_sphere s, int which, int mlevel, double *host)
{
  REQUIRE(s && host && which >= 0 && which <= 4, "bad argument");
  REQUIRE(which >= 3 || (mlevel >= 0 && mlevel < s->nlev), "mlevel=%d outside [0,%d)", mlevel, s->nlev);
  DeviceGuard g(s->device);
  CU_TRY(cudaStreamSynchronize(s->work_stream()));
  const double *src = which == 3 ? s->coef_use : which == 4 ? s->coef
                    : (which == 0 ? s->coefN : which == 1 ? s->coefL : s->differ) + (size_t)mlevel * s->ncoef;
  CU_TRY(cudaMemcpy(host, src, s->ncoef * sizeof(double), cudaMemcpyDeviceToHost));
  return EXPB200_OK;
}

int expb200_sphere_set_level_coefficients(expb200_sphere s, int which, int mlevel, const double *host)
{
  REQUIRE(s && host && which >= 0 && which <= 1 && mlevel >= 0 && mlevel < s->nlev, "bad argument");
  DeviceGuard g(s->device);
  CU_TRY(cudaStreamSynchronize(s->work_stream()));
  double *dst = (which == 0 ? s->coefN : s->coefL) + (size_t)mlevel * s->ncoef;
  CU_TRY(cudaMemcpy(dst, host, s->ncoef * sizeof(double), cudaMemcpyHostToDevice));
  return EXPB200_OK;
}

// ~ multistep_update_begin / multistep_update (per switching particle) /
// multistep_update_finish (src/SphericalBasis.cc:1013-1228): for every particle whose
// level changed in the last expb200_adjust_multistep_level, subtract its basis
// contribution from differ[old level] and add it to differ[new level]; with
// finish != 0 also expcoefN[M] += differ[M] (single rank; with several ranks all-reduce
// differ first, then call expb200_sphere_multistep_update_finish).
int expb200_sphere_multistep_update(expb200_sphere s, expb200_component c, double adb, int finish)
{
  REQUIRE(s && c, "NULL argument");
  REQUIRE(s->device == c->device, "device mismatch");
  DeviceGuard g(s->device);
  cudaStream_t st = c->stream;
  { cudaStream_t ws = s->work_stream(); if (ws != st) CU_TRY(cudaStreamSynchronize(ws)); }
  s->last_stream = st;
  const size_t all = (size_t)s->nlev * s->ncoef;
  CU_TRY(cudaMemsetAsync(s->differ, 0, all * sizeof(double), st));
  if (c->n && s->nlev > 1) {
    SphGeom gm = geom_for(s, c);
    const uint32_t nkeys = (uint32_t)s->nlev * (uint32_t)s->cfg.numr + 1;
    int rc = ensure_seg(c, nkeys);
    if (rc) return rc;
    // the switching particles are few: compact their store indices (ascending, deterministic) and
    // bin only those -- two full (level, cell) sorts of the component per substep otherwise
    // (round 2: half of the disk+halo step's sort time, profiles/r02q_bench_diskhalo.json)
    size_t nsw = 0;
    uint32_t *list = c->rank + c->n / 2;             // the sort of nsw <= n/2 items uses rank[0, nsw)
    rc = compact_differing(st, c->level, c->prev_level, c->n, c->hist, c->keys[1], &nsw);
    if (rc) return rc;
    const bool compact = nsw > 0 && 2 * nsw <= c->n;
    if (compact)
      CU_TRY(cudaMemcpyAsync(list, c->keys[1], nsw * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    const size_t nn = compact ? nsw : c->n;          // items binned per side
    const int CH = deposit_chunk(nn);
    if ((rc = ensure_partials(s, nn, CH))) return rc;
    if (compact && c->sw_cap < nsw) {
      if (c->sw_perm) cudaFree(c->sw_perm);
      c->sw_perm = nullptr; c->sw_cap = 0;
      CU_TRY(cudaMalloc(&c->sw_perm, nsw * sizeof(uint32_t)));
      c->sw_cap = nsw;
    }
    // the compact path leaves perm[cur] (the level partition of the component's own sort) alone:
    // its permutation buffers are the scratch one and the spare half
    uint32_t *kl[2] = {c->keys[0], c->keys[1]};
    uint32_t *pl2[2] = {c->sw_perm, c->perm[c->cur ^ 1]};
    for (int side = 0; side < 2 && nsw > 0; side++) {          // 0: "to" (+), 1: "from" (-)
      const uint32_t *skeys = nullptr;
      if (compact) {
        sph_switch_key_list_kernel<<<grid_for(nsw, 256), 256, 0, st>>>(gm, c->posm, c->level, c->prev_level, list,
                                                                       nsw, side, s->nlev, kl[0]);
        CU_LAUNCH_CHECK();
        ProfRegion pr(K_SORT, st, radix_sort_launches(nsw, nkeys, true, true));
        SortPayload pl;
        pl.src = c->posm;
        pl.dst = c->posm_sorted;
        int cur = 0;
        rc = radix_sort_pairs(st, kl, pl2, c->hist, c->rank, nsw, nkeys, &cur, c->seg, pl, list);
        if (rc) return rc;
        skeys = kl[cur];
        c->psort = c->posm_sorted;
      } else {
        sph_switch_key_kernel<<<grid_for(c->n, 256), 256, 0, st>>>(gm, c->posm, c->level, c->prev_level,
                                                                   c->n, side, s->nlev, c->keys[0]);
        CU_LAUNCH_CHECK();
        if ((rc = sort_keyed(s, c, nkeys, false))) return rc;
        skeys = c->keys[c->cur];
      }
      for (int M = 0; M < s->nlev; M++) {
        const int keybase = M * s->cfg.numr;
        unsigned nchunks = div_up(nn, CH);
        rc = launch_deposit(s->cfg.lmax, st, nchunks, gm, s->p0, c->psort, skeys,
                            c->seg, keybase, CH, adb, s->PA, s->PB);
        if (rc) return rc;
        sph_combine_kernel<<<gm.numr, 64, 0, st>>>(gm, c->seg, keybase, CH, 0, s->PA, s->PB, s->G);
        sph_contract_kernel<<<dim3(gm.nrow, gm.nmax), 256, 0, st>>>(gm, s->Ep, s->G, s->row_l, s->row_m, 0, s->tmp);
        axpy_kernel<<<div_up(s->ncoef, 256), 256, 0, st>>>(s->differ + (size_t)M * s->ncoef, s->tmp,
                                                          side == 0 ? 1.0 : -1.0, s->ncoef);
        CU_LAUNCH_CHECK();
      }
    }
    if (nsw > 0) {
      // keys, cell offsets and the sorted position stream were overwritten: the cell order has
      // to be rebuilt before its next use; the level partition survives the compact path
      c->sort_pos_ver = 0;
      if (!compact) { c->sort_kind = 0; c->part_ok = false; }
    }
  }
  if (finish) return expb200_sphere_multistep_update_finish(s);
  return EXPB200_OK;
}

int expb200_sphere_differ_device(expb200_sphere s, double **dptr, size_t *count, void **stream)
{
  REQUIRE(s && dptr, "NULL argument");
  *dptr = s->differ;
  if (count) *count = (size_t)s->nlev * s->ncoef;
  if (stream) *stream = s->work_stream();
  return EXPB200_OK;
}

int expb200_sphere_multistep_update_finish(expb200_sphere s)
{
  REQUIRE(s, "sphere is NULL");
  DeviceGuard g(s->device);
  const size_t all = (size_t)s->nlev * s->ncoef;
  axpy_kernel<<<div_up(all, 256), 256, 0, s->work_stream()>>>(s->coefN, s->differ, 1.0, all);
  CU_LAUNCH_CHECK();
  return EXPB200_OK;
}

// nsteps single-level KDK steps of a self-gravitating component back to back
// (do_step, multistep == 0 branch, src/step.cc:272-322, in the main loop of src/expand.cc).
// Per step: [kick(dt/2) + drift(dt) + key] | bin | deposit/combine/contract | [all-reduce] |
// field | force (overwriting acc/pot: no zero pass) + kick(dt/2); between two steps of the
// same call the first bracket rides in the previous step's force kernel.
int expb200_sphere_steps(expb200_sphere s, expb200_component c, double dt, int nsteps, double adb,
                         void *nccl_comm)
{
  REQUIRE(s && c, "NULL argument");
  REQUIRE(s->nlev == 1, "expb200_sphere_step(s) is the single-level driver (multistep=0)");
  REQUIRE(nsteps >= 0, "nsteps=%d", nsteps);
  DeviceGuard g(s->device);
  int rc;
  { cudaStream_t ws = s->work_stream(); if (ws != c->stream) CU_TRY(cudaStreamSynchronize(ws)); }
  const uint32_t nkeys = (uint32_t)s->cfg.numr;
  if ((rc = ensure_seg(c, nkeys))) return rc;
  if (c->n == 0 || nsteps == 0) return EXPB200_OK;
  if ((rc = ensure_tdesc(s, c->n))) return rc;
  SphGeom gm = geom_for(s, c);
  for (int step = 0; step < nsteps; step++) {
    if (step == 0) {
      ProfRegion pr(K_KDK, c->stream);
      sph_kick_drift_key_kernel<<<grid_for(c->n, 256), 256, 0, c->stream>>>(gm, c->posm, c->vel, c->accp, c->n,
                                                                            0.5 * dt, dt, c->keys[0]);
      CU_LAUNCH_CHECK();
    }
    c->pos_ver++;
    if ((rc = sort_keyed(s, c, nkeys, true, deposit_can_gather(s->cfg.lmax)))) return rc;
    mark_sorted(s, c);
    if ((rc = accumulate_sorted(s, c, 0, adb))) return rc;
    if (nccl_comm) {
      ProfRegion pr(K_ALLREDUCE, c->stream, 0);
      if ((rc = nccl_allreduce_sum_f64(nccl_comm, s->coef, s->ncoef, c->stream))) return rc;
    }
    CU_TRY(cudaMemcpyAsync(s->coef_use, s->coef, s->ncoef * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    if ((rc = build_field(s, c->stream))) return rc;
    const bool last = step == nsteps - 1;
    // the force kernel reads the sorted keys (tile descriptors) and, when another step follows,
    // writes the next step's keys: after a two-pass sort the sorted keys sit in keys[0], so the
    // new ones go through the sort's scratch array (the same kernel runs either way, which keeps
    // step() x k and steps(k) bit-identical)
    const bool alias = !last && c->keys[c->cur] == c->keys[0];
    ForceOut F{c->accp, c->vel, c->posm, alias ? c->rank : c->keys[0], 0.5 * dt, dt, last ? 1 : 2};
    {
      ProfRegion pr(K_FORCE, c->stream, 2);   // tile pre-pass + force
      rc = launch_force(s->cfg.lmax, c->stream, gm, s->T, c->psort, F, c->perm[c->cur], c->keys[c->cur],
                        c->seg, 0, c->n, s->tdesc);
      if (rc) return rc;
    }
    if (alias)
      CU_TRY(cudaMemcpyAsync(c->keys[0], c->rank, c->n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c->stream));
  }
  return EXPB200_OK;
}

int expb200_sphere_step(expb200_sphere s, expb200_component c, double dt, double adb, void *nccl_comm)
{
  return expb200_sphere_steps(s, c, dt, 1, adb, nccl_comm);
}

// ~ the `compute && pcavar` branch of determine_coefficients_thread
// (src/SphericalBasis.cc:501-509, 525-545, 563-581) for levlist[mlevel]: per jackknife bin
// T = (indx_base + host slot) % sampT the number, mass, coefficient share and covariance
// matrices of the bin's particles.  Host outputs (any may be NULL):
//   countT[sampT], massT[sampT], coefT[sampT][ldim][nmax][2], coefM[sampT][ldim][nmax][nmax]
// with ldim = (lmax+1)(lmax+2)/2, index l(l+1)/2 + m (the reference's iC).
static int sphere_covariance_impl(expb200_sphere s, expb200_component c, int mlevel, double adb, int sampT,
                                  uint64_t indx_base, unsigned *countT, double *massT, double *coefT,
                                  double *coefM, int unit_m)
{
  REQUIRE(s && c, "NULL argument");
  REQUIRE(s->device == c->device, "device mismatch");
  REQUIRE(mlevel >= 0 && mlevel < s->nlev, "mlevel=%d outside [0,%d]", mlevel, s->nlev - 1);
  REQUIRE(sampT >= 1 && (uint64_t)sampT * s->cfg.numr < 0x7fffffffull, "sampT=%d out of range", sampT);
  REQUIRE(s->cfg.nmax <= 8 * kCovTiles, "covariance supports nmax <= %d", 8 * kCovTiles);
  REQUIRE(s->cfg.lmax <= 6, "covariance supports lmax <= 6");
  DeviceGuard gd(s->device);
  cudaStream_t st = c->stream;
  { cudaStream_t ws = s->work_stream(); if (ws != st) CU_TRY(cudaStreamSynchronize(ws)); }
  s->last_stream = st;
  SphGeom gm = geom_for(s, c);
  gm.cov_unit_m = unit_m;
  const int R = gm.numr, N = gm.nmax, ldim = gm.nslot;
  const uint32_t nkeys = (uint32_t)sampT * R + 1;
  int rc = ensure_seg(c, nkeys);
  if (rc) return rc;
  if (countT) for (int T = 0; T < sampT; T++) countT[T] = 0;
  if (massT) for (int T = 0; T < sampT; T++) massT[T] = 0.0;
  if (coefT) memset(coefT, 0, (size_t)sampT * ldim * N * 2 * sizeof(double));
  if (coefM) memset(coefM, 0, (size_t)sampT * ldim * N * N * sizeof(double));
  if (c->n == 0) return EXPB200_OK;

  ProfRegion pr(K_COV, st, 0);
  sph_cov_key_kernel<<<grid_for(c->n, 256), 256, 0, st>>>(gm, c->posm, c->level,
                                                          c->identity_order ? nullptr : c->slot_of, c->n,
                                                          s->nlev > 1, mlevel, (uint32_t)sampT, indx_base,
                                                          c->keys[0]);
  CU_LAUNCH_CHECK();
  if ((rc = sort_keyed(s, c, nkeys, false))) return rc;
  c->sort_kind = 0; c->part_ok = false;                       // the component's regular sort was overwritten

  // bin boundaries on the host: seg[T*R] (first particle of bin T) and seg[T*R + R - 1]
  std::vector<uint32_t> hseg((size_t)nkeys + 1);
  CU_TRY(cudaMemcpyAsync(hseg.data(), c->seg, hseg.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  size_t maxlen = 1;
  for (int T = 0; T < sampT; T++) {
    size_t len = hseg[(size_t)T * R + R - 1] - hseg[(size_t)T * R];
    if (countT) countT[T] = (unsigned)len;
    if (len > maxlen) maxlen = len;
  }
  int CH = (int)(maxlen / 8 / 32) * 32;
  if (CH < 32) CH = 32;
  if (CH > 2048) CH = 2048;
  const unsigned nchunks = div_up(maxlen, CH);

  // bins are processed TB at a time; scratch sized for one batch
  const size_t per_bin = (size_t)(R + 2) + nchunks + 1;              // partial slots per bin
  const size_t bytes_per_bin = per_bin * (2 * gm.nrow + 3 * ldim + 1) * sizeof(double) +
                               ((size_t)gm.nrow * R + 2 * (size_t)ldim * R + R + (size_t)gm.nrow * N +
                                (size_t)ldim * N * N) * sizeof(double);
  int TB = (int)((size_t)1 << 30) / (int)(bytes_per_bin > ((size_t)1 << 30) ? ((size_t)1 << 30) : bytes_per_bin);
  if (TB < 1) TB = 1;
  if (TB > sampT) TB = sampT;
  if (TB > 65535) TB = 65535;
  double *PA = nullptr, *PB = nullptr, *PS = nullptr, *G = nullptr, *DO = nullptr, *Mc = nullptr,
         *rows = nullptr, *Ms = nullptr, *mT = nullptr, *outT = nullptr, *outM = nullptr;
  // scratch is cached in the sphere (grow-only, freed with it): cudaMalloc/cudaFree of these
  // buffers per call costs more than the kernels
  auto cleanup = [&]() {};
  auto cov_alloc = [&](double **ptr, size_t bytes, int slot) -> cudaError_t {
    if (s->cov_cap[slot] < bytes) {
      if (s->cov_buf[slot]) cudaFree(s->cov_buf[slot]);
      s->cov_buf[slot] = nullptr; s->cov_cap[slot] = 0;
      const size_t want = bytes + bytes / 8;        // head-room: bin populations drift from call to call
      cudaError_t e = cudaMalloc(&s->cov_buf[slot], want);
      if (e != cudaSuccess) return e;
      s->cov_cap[slot] = want;
    }
    *ptr = static_cast<double *>(s->cov_buf[slot]);
    return cudaSuccess;
  };
#define COV_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { cleanup(); return cuda_fail(_e, #expr, __FILE__, __LINE__); } } while (0)
  const size_t slots = per_bin * TB + 2;
  COV_TRY(cov_alloc(&PA, slots * gm.nrow * sizeof(double), 0));
  COV_TRY(cov_alloc(&PB, slots * gm.nrow * sizeof(double), 1));
  COV_TRY(cov_alloc(&PS, slots * (3 * ldim + 1) * sizeof(double), 2));
  COV_TRY(cov_alloc(&G, (size_t)TB * gm.nrow * R * sizeof(double), 3));
  COV_TRY(cov_alloc(&DO, (size_t)TB * 2 * ldim * R * sizeof(double), 4));
  COV_TRY(cov_alloc(&Mc, (size_t)TB * R * sizeof(double), 5));
  COV_TRY(cov_alloc(&rows, (size_t)TB * gm.nrow * N * sizeof(double), 6));
  COV_TRY(cov_alloc(&Ms, (size_t)TB * ldim * N * N * sizeof(double), 7));
  COV_TRY(cov_alloc(&mT, (size_t)TB * sizeof(double), 8));
  COV_TRY(cov_alloc(&outT, (size_t)TB * ldim * N * 2 * sizeof(double), 9));
  COV_TRY(cov_alloc(&outM, (size_t)TB * ldim * N * N * sizeof(double), 10));

  for (int T0 = 0; T0 < sampT; T0 += TB) {
    const int nb = sampT - T0 < TB ? sampT - T0 : TB;
    const int keybase = T0 * R;
    // linear sums of every bin: the ordinary deposit, batched over the bins
    rc = launch_deposit(s->cfg.lmax, st, nchunks, gm, s->p0, c->psort, c->keys[c->cur], c->seg, keybase, CH,
                        adb, PA, PB, nb);
    if (rc) { cleanup(); return rc; }
    sph_combine_kernel<<<dim3(R, nb), 64, 0, st>>>(gm, c->seg, keybase, CH, 1, PA, PB, G);
    sph_contract_kernel<<<dim3(gm.nrow, N, nb), 256, 0, st>>>(gm, s->Ep, G, s->row_l, s->row_m,
                                                              s->cfg.M0_only, rows);
    // quadratic sums, then the E^T (Tri E) products on the tensor cores
    rc = launch_cov_deposit(s->cfg.lmax, st, dim3(div_up(nchunks, kDepositWarps), nb), gm, s->p0, c->psort,
                            c->keys[c->cur], c->seg, keybase, CH, adb, PS);
    if (rc) { cleanup(); return rc; }
    sph_cov_combine_kernel<<<dim3(R, nb), 32, 0, st>>>(gm, c->seg, keybase, CH, PS, DO, Mc);
    sph_cov_mass_kernel<<<nb, 32, 0, st>>>(gm, Mc, mT);
    sph_cov_contract_kernel<<<dim3(ldim, nb), 128, 0, st>>>(gm, s->Ep, DO, s->slot_l, s->slot_m,
                                                            s->cfg.M0_only, Ms);
    sph_cov_pack_kernel<<<dim3(ldim, nb), 128, 0, st>>>(gm, rows, Ms, outT, outM);
    COV_TRY(cudaGetLastError());
    if (massT) COV_TRY(cudaMemcpyAsync(massT + T0, mT, nb * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (coefT) COV_TRY(cudaMemcpyAsync(coefT + (size_t)T0 * ldim * N * 2, outT, (size_t)nb * ldim * N * 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (coefM) COV_TRY(cudaMemcpyAsync(coefM + (size_t)T0 * ldim * N * N, outM, (size_t)nb * ldim * N * N * sizeof(double), cudaMemcpyDeviceToHost, st));
    COV_TRY(cudaStreamSynchronize(st));
  }
#undef COV_TRY
  cleanup();
  return EXPB200_OK;
}

int expb200_sphere_covariance(expb200_sphere s, expb200_component c, int mlevel, double adb, int sampT,
                              uint64_t indx_base, unsigned *countT, double *massT, double *coefT,
                              double *coefM)
{
  return sphere_covariance_impl(s, c, mlevel, adb, sampT, indx_base, countT, massT, coefT, coefM, 0);
}

// ~ the `compute && pcaeof` branch (src/SphericalBasis.cc:536-545, 573-581):
// tvar[lm](n,o) = sum_p wk[n] wk[o] / mass with wk = potd(l,n) facL mass fac0 for m = 0 and
// wk = potd(l,n) mass fac0 (no angular factor) for m > 0, over levlist[mlevel];
// tvar [ldim][nmax][nmax] is overwritten.  Same pipeline as the pcavar matrices with one bin.
int expb200_sphere_covariance_eof(expb200_sphere s, expb200_component c, int mlevel, double adb, double *tvar)
{
  REQUIRE(tvar, "tvar is NULL");
  return sphere_covariance_impl(s, c, mlevel, adb, 1, 0, nullptr, nullptr, nullptr, tvar, 1);
}

} // extern "C"
