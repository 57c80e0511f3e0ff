// component.cu -- device-resident particle store, leapfrog and multistep level
// assignment.  Replaces the reference's Component CUDA seams
// (src/cudaComponent.cu), incr_position_cuda / incr_velocity_cuda
// (src/cudaIncpos.cu, src/cudaIncvel.cu) and cuda_compute_levels
// (src/cudaMultistep.cu); the arithmetic follows the CPU path
// (src/incpos.cc:63-66, src/incvel.cc:68-69, src/multistep.cc:52-236).
#include "common.cuh"
#include "sort.cuh"
#include "profile.cuh"
#include "nccl_shim.h"

#include <algorithm>
#include <cstring>
#include <thread>
#include <tuple>
#include <cmath>
#include <mutex>
#include <vector>

namespace expb200 {

// ---------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------
static thread_local std::string g_err;

void set_error(const char *fmt, ...)
{
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line)
{
  set_error("CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
  return e == cudaErrorMemoryAllocation ? EXPB200_ENOMEM : EXPB200_ECUDA;
}

static std::mutex g_stream_mu;
static std::vector<cudaStream_t> g_live_streams;

void stream_register(cudaStream_t st, bool live)
{
  std::lock_guard<std::mutex> lk(g_stream_mu);
  for (size_t k = 0; k < g_live_streams.size(); k++)
    if (g_live_streams[k] == st) { g_live_streams.erase(g_live_streams.begin() + k); break; }
  if (live) g_live_streams.push_back(st);
}

bool stream_is_live(cudaStream_t st)
{
  std::lock_guard<std::mutex> lk(g_stream_mu);
  for (auto s : g_live_streams) if (s == st) return true;
  return false;
}

namespace {

constexpr size_t kStageChunk = 1u << 20;   // particles per staging chunk
constexpr size_t kBounceChunk = 1u << 18;  // particles per pinned bounce buffer (14 MB)
constexpr int kStageDoubles = 10;          // pos3 vel3 acc3 pot1 (download) / pos3 vel3 mass1 (upload)

// ---------------------------------------------------------------------------
// pack / unpack between host AoS field arrays and the 32-byte device records
// ---------------------------------------------------------------------------
__global__ void pack_kernel(rec4 *__restrict__ posm, rec4 *__restrict__ vel, rec4 *__restrict__ accp,
                            const double *__restrict__ pos, const double *__restrict__ v,
                            const double *__restrict__ mass, const double *__restrict__ potext,
                            size_t first, size_t count, int reset_acc, const uint32_t *__restrict__ where)
{
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count;
       i += (size_t)gridDim.x * blockDim.x) {
    size_t s = where ? where[first + i] : first + i;
    st_rec(&posm[s], rec4{pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], mass[i]});
    double pe = potext ? potext[i] : 0.0;
    if (v) st_rec(&vel[s], rec4{v[3 * i], v[3 * i + 1], v[3 * i + 2], pe});
    else   st_rec(&vel[s], rec4{0.0, 0.0, 0.0, pe});
    if (reset_acc) st_rec(&accp[s], rec4{0.0, 0.0, 0.0, 0.0});
  }
}

__global__ void unpack_kernel(const rec4 *__restrict__ posm, const rec4 *__restrict__ vel,
                              const rec4 *__restrict__ accp, double *__restrict__ pos,
                              double *__restrict__ v, double *__restrict__ acc,
                              double *__restrict__ pot, size_t first, size_t count,
                              const uint32_t *__restrict__ where)
{
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count;
       i += (size_t)gridDim.x * blockDim.x) {
    size_t s = where ? where[first + i] : first + i;
    if (pos) { rec4 p = ld_rec_rw(&posm[s]); pos[3 * i] = p.x; pos[3 * i + 1] = p.y; pos[3 * i + 2] = p.z; }
    if (v)   { rec4 q = ld_rec_rw(&vel[s]);  v[3 * i] = q.x;   v[3 * i + 1] = q.y;   v[3 * i + 2] = q.z; }
    if (acc || pot) {
      rec4 a = ld_rec_rw(&accp[s]);
      if (acc) { acc[3 * i] = a.x; acc[3 * i + 1] = a.y; acc[3 * i + 2] = a.z; }
      if (pot) pot[i] = a.w;
    }
  }
}

// ---- storage re-ordering helpers ---------------------------------------------------
__global__ void gather_rec_kernel(rec4 *__restrict__ dst, const rec4 *__restrict__ src,
                                  const uint32_t *__restrict__ perm, size_t n)
{
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x)
    st_rec(&dst[k], ld_rec(&src[perm[k]]));
}
__global__ void scatter_rec_kernel(rec4 *__restrict__ dst, const rec4 *__restrict__ src,
                                   const uint32_t *__restrict__ where, size_t n)
{
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    st_rec(&dst[where[i]], ld_rec(&src[i]));
}
// dst[k] = src[idx ? idx[k] : k]
__global__ void gather_u32_kernel(uint32_t *__restrict__ dst, const uint32_t *__restrict__ src,
                                  const uint32_t *__restrict__ idx, size_t n)
{
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x)
    dst[k] = src[idx ? idx[k] : k];
}
// dst[idx[i]] = src[i]
__global__ void scatter_u32_kernel(uint32_t *__restrict__ dst, const uint32_t *__restrict__ src,
                                   const uint32_t *__restrict__ idx, size_t n)
{
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[idx[i]] = src[i];
}
// where[slot_of[k]] = k; slot_of == NULL means the identity
__global__ void invert_kernel(uint32_t *__restrict__ where, const uint32_t *__restrict__ slot_of, size_t n)
{
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x)
    where[slot_of[k]] = (uint32_t)k;
}
__global__ void iota_kernel(uint32_t *__restrict__ a, size_t n)
{
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x)
    a[k] = (uint32_t)k;
}

__global__ void fill_u32_kernel(unsigned *a, size_t n, unsigned v)
{
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) a[i] = v;
}
__global__ void fill_f32_kernel(float *a, size_t n, float v)
{
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) a[i] = v;
}

// ---------------------------------------------------------------------------
// a21: leapfrog.  pos += vel*dt (src/incpos.cc:63-66), vel += acc*dt
// (src/incvel.cc:68-69) over levlist[mlevel] (mlevel>=0) or every particle.
// Natural order, one 32-byte record in and out per particle: HBM-bound.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
drift_kernel(rec4 *__restrict__ posm, const rec4 *__restrict__ vel,
             const unsigned *__restrict__ level, size_t n, double dt, int mlevel, int dim)
{
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    if (mlevel >= 0 && level[i] != (unsigned)mlevel) continue;
    rec4 p = ld_rec_rw(&posm[i]);
    rec4 v = ld_rec_rw(&vel[i]);
    p.x = fma(v.x, dt, p.x);
    if (dim > 1) p.y = fma(v.y, dt, p.y);
    if (dim > 2) p.z = fma(v.z, dt, p.z);
    st_rec(&posm[i], p);
  }
}

__global__ void __launch_bounds__(256)
kick_kernel(rec4 *__restrict__ vel, const rec4 *__restrict__ accp,
            const unsigned *__restrict__ level, size_t n, double dt, int mlevel, int dim)
{
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    if (mlevel >= 0 && level[i] != (unsigned)mlevel) continue;
    rec4 v = ld_rec_rw(&vel[i]);
    rec4 a = ld_rec_rw(&accp[i]);
    v.x = fma(a.x, dt, v.x);
    if (dim > 1) v.y = fma(a.y, dt, v.y);
    if (dim > 2) v.z = fma(a.z, dt, v.z);
    st_rec(&vel[i], v);
  }
}

// The kicks and drifts of all the levels >= lo of a sub-step in one pass: particle i of level M
// gets v += a dtk[M], then x += v dtd[M] (each skipped when its table is absent), the arithmetic of
// kick_kernel / drift_kernel.
struct LevelDt { double k[32], d[32]; int kick, drift; };

__global__ void __launch_bounds__(256)
move_levels_kernel(rec4 *__restrict__ posm, rec4 *__restrict__ vel, const rec4 *__restrict__ accp,
                   const unsigned *__restrict__ level, size_t n, LevelDt D, int lo, int nlev)
{
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    const unsigned lev = level[i];
    if (lev < (unsigned)lo || lev >= (unsigned)nlev) continue;
    rec4 v = ld_rec_rw(&vel[i]);
    if (D.kick) {
      const rec4 a = ld_rec_rw(&accp[i]);
      const double dt = D.k[lev];
      v.x = fma(a.x, dt, v.x);
      v.y = fma(a.y, dt, v.y);
      v.z = fma(a.z, dt, v.z);
      st_rec(&vel[i], v);
    }
    if (D.drift) {
      rec4 p = ld_rec_rw(&posm[i]);
      const double dt = D.d[lev];
      p.x = fma(v.x, dt, p.x);
      p.y = fma(v.y, dt, p.y);
      p.z = fma(v.z, dt, p.z);
      st_rec(&posm[i], p);
    }
  }
}

// ~ zeroPotAccKernel (src/cudaComponent.cu:729-748)
__global__ void __launch_bounds__(256)
zero_kernel(rec4 *__restrict__ accp, const unsigned *__restrict__ level, size_t n, int minlev)
{
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    if (minlev > 0 && level[i] < (unsigned)minlev) continue;
    st_rec(&accp[i], rec4{0.0, 0.0, 0.0, 0.0});
  }
}

// ---------------------------------------------------------------------------
// a22: adjust_multistep_level_thread (src/multistep.cc:52-236).
// Bit-exact with the CPU path: every operation is an explicit IEEE round-to-
// nearest intrinsic (no FMA contraction), dtreq rounds through float, and
// floor(log(dtime/dtreq)/log(2)) is evaluated as the exact binary exponent,
// except when the ratio is exactly 2^k, where the host libm's own answer
// (pow2lev[k], computed with log() on the host at call time) is used.
// ---------------------------------------------------------------------------
struct LevelArgs {
  double dtime, fS, fD, fV, fA, fP;
  int multistep, shiftlevl, mfirst, minlev, noswitch, reset_dtreq, apply, dim;
  int pow2lev[64];
};

__global__ void __launch_bounds__(256)
levels_kernel(const rec4 *__restrict__ vel, const rec4 *__restrict__ accp,
              const float *__restrict__ scale, unsigned *__restrict__ level,
              float *__restrict__ dtreq, size_t n, LevelArgs A,
              unsigned long long *__restrict__ stats)
{
  const double eps = 1.0e-10;
  unsigned long long examined = 0, switched = 0, offlo = 0, offhi = 0;
  double mindt = 1.0e300, maxdt = 0.0;

  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    unsigned plev = level[i];
    if ((int)plev < A.minlev) continue;

    rec4 v = ld_rec_rw(&vel[i]);
    rec4 a = ld_rec_rw(&accp[i]);

    double dtr = __dmul_rn(v.x, a.x), vtot = __dmul_rn(v.x, v.x), atot = __dmul_rn(a.x, a.x);
    if (A.dim > 1) {
      dtr = __dadd_rn(dtr, __dmul_rn(v.y, a.y));
      vtot = __dadd_rn(vtot, __dmul_rn(v.y, v.y));
      atot = __dadd_rn(atot, __dmul_rn(a.y, a.y));
    }
    if (A.dim > 2) {
      dtr = __dadd_rn(dtr, __dmul_rn(v.z, a.z));
      vtot = __dadd_rn(vtot, __dmul_rn(v.z, v.z));
      atot = __dadd_rn(atot, __dmul_rn(a.z, a.z));
    }
    double ptot = fabs(__dadd_rn(a.w, v.w));          // |pot + potext|

    double dsr = (double)scale[i];
    double dts;
    if (dsr > 0) dts = __ddiv_rn(__dmul_rn(A.fS, dsr), fabs(__dadd_rn(__dsqrt_rn(vtot), eps)));
    else         dts = __ddiv_rn(1.0, eps);

    double dtd = __ddiv_rn(__dmul_rn(A.fD, 1.0), __dsqrt_rn(__dadd_rn(vtot, eps)));
    double dtv = __dmul_rn(A.fV, __dsqrt_rn(__ddiv_rn(vtot, __dadd_rn(atot, eps))));
    double dta = __ddiv_rn(__dmul_rn(A.fA, ptot), __dadd_rn(fabs(dtr), eps));
    double dtA = __dmul_rn(A.fP, __dsqrt_rn(__ddiv_rn(ptot, __dadd_rn(atot, eps))));

    double best = dtd;
    if (dtv < best) best = dtv;
    if (dts < best) best = dts;
    if (dta > 0.0 && dta < best) best = dta;
    if (dtA > 0.0 && dtA < best) best = dtA;
    double dt = best > eps ? best : eps;              // std::max(eps, best)

    float dq;
    if (A.noswitch) {
      dq = dtreq[i];
      if (A.reset_dtreq) dq = __double2float_rn(1.7976931348623157e308);
      if (dt < (double)dq) dq = __double2float_rn(dt);
    } else {
      dq = __double2float_rn(dt);
    }
    dtreq[i] = dq;

    if (A.apply) {
      unsigned nlev = plev;
      double dqd = (double)dq;
      if (dqd > A.dtime) {
        nlev = 0;
        offhi++;
        maxdt = fmax(maxdt, dqd);
      } else {
        double x = __ddiv_rn(A.dtime, dqd);           // >= 1
        long long bits = __double_as_longlong(x);
        int e = (int)((bits >> 52) & 0x7ff) - 1023;
        bool pow2 = (bits & 0x000fffffffffffffLL) == 0;
        if (e > 62) e = 62;                           // inf / absurd ratios: clamp, multistep wins below
        if (e < 0) e = 0;
        nlev = (unsigned)(pow2 ? A.pow2lev[e] : e);
      }
      if (A.shiftlevl) {
        if (nlev > plev) {
          if (nlev - plev > (unsigned)A.shiftlevl) nlev = plev + A.shiftlevl;
        } else if (plev > nlev) {
          if (plev - nlev > (unsigned)A.shiftlevl) nlev = plev - A.shiftlevl;
        }
      }
      if (nlev > (unsigned)A.multistep) {
        nlev = A.multistep;
        offlo++;
        mindt = fmin(mindt, dqd);
      }
      if ((int)nlev < A.mfirst) nlev = A.mfirst;
      if (nlev != plev) {
        level[i] = nlev;
        switched++;
      }
      examined++;
    }
  }

  // CTA reduction of the counters, then one integer atomic per CTA
  __shared__ unsigned long long sh[4][8];
  __shared__ double shd[2][8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    examined += __shfl_down_sync(0xffffffffu, examined, o);
    switched += __shfl_down_sync(0xffffffffu, switched, o);
    offlo += __shfl_down_sync(0xffffffffu, offlo, o);
    offhi += __shfl_down_sync(0xffffffffu, offhi, o);
    mindt = fmin(mindt, __shfl_down_sync(0xffffffffu, mindt, o));
    maxdt = fmax(maxdt, __shfl_down_sync(0xffffffffu, maxdt, o));
  }
  if (lane == 0) { sh[0][w] = examined; sh[1][w] = switched; sh[2][w] = offlo; sh[3][w] = offhi;
                   shd[0][w] = mindt; shd[1][w] = maxdt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long e = 0, s = 0, lo = 0, hi = 0;
    double mn = 1.0e300, mx = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); k++) {
      e += sh[0][k]; s += sh[1][k]; lo += sh[2][k]; hi += sh[3][k];
      mn = fmin(mn, shd[0][k]); mx = fmax(mx, shd[1][k]);
    }
    if (e)  atomicAdd(&stats[0], e);
    if (s)  atomicAdd(&stats[1], s);
    if (lo) atomicAdd(&stats[2], lo);
    if (hi) atomicAdd(&stats[3], hi);
    // positive doubles order like their bit patterns
    atomicMin(&stats[4], (unsigned long long)__double_as_longlong(mn));
    atomicMax(&stats[5], (unsigned long long)__double_as_longlong(mx));
  }
}

// ---- centre of mass (Component::fix_positions) ------------------------------------
// Per multistep level M >= mlevel the sums over the non-frozen particles of
// {mass, mass*pos[k], mass*vel[k], mass*acc[k]} (k < dim), src/Component.cc:3280-3352.
// One warp per tile of 32 particles; for every level present in the tile (ballot) the ten
// values are reduced with the fixed reduce-scatter butterfly (RS<10>) and the lane that
// ends up owning a value adds it to the warp's shared-memory accumulator.  CTA partials are
// written to `part` and summed in block order by com_final_kernel: no atomics, so the
// result is bit-reproducible.
constexpr int kComWarps = 8;
constexpr int kComVals = 10;
constexpr int kComMaxLev = 16;

__global__ void __launch_bounds__(kComWarps * 32)
com_sum_kernel(const rec4 *__restrict__ posm, const rec4 *__restrict__ vel, const rec4 *__restrict__ accp,
               const unsigned *__restrict__ level, size_t n, int mlevel, int nlev, int dim,
               double cx, double cy, double cz, double rtrunc2, double *__restrict__ part)
{
  __shared__ double acc[kComWarps][kComMaxLev][kComVals];
  using R = RS<kComVals, 0>;
  static_assert(R::FIN == 1, "one value per lane after the reduce-scatter");
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int q = threadIdx.x; q < kComWarps * kComMaxLev * kComVals; q += blockDim.x) (&acc[0][0][0])[q] = 0.0;
  __syncthreads();
  const int myrow = R::row(lane, 0);
  const size_t step = (size_t)gridDim.x * kComWarps * 32;
  for (size_t base = ((size_t)blockIdx.x * kComWarps + w) * 32; base < n; base += step) {
    const size_t i = base + lane;
    int lev = -1;
    rec4 p = rec4{0, 0, 0, 0}, v = p, a = p;
    if (i < n) {
      p = ld_rec(&posm[i]);
      const double dx = p.x - cx, dy = p.y - cy, dz = p.z - cz;
      const int l = (int)level[i];
      if (!(dx * dx + dy * dy + dz * dz > rtrunc2) && l >= mlevel && l < nlev) {   // Component::freeze
        lev = l;
        v = ld_rec(&vel[i]);
        a = ld_rec(&accp[i]);
      }
    }
    for (int mm = mlevel; mm < nlev; mm++) {
      if (!__ballot_sync(0xffffffffu, lev == mm)) continue;
      const double m = lev == mm ? p.w : 0.0;
      double A[kComVals];
      A[0] = m;
      A[1] = m * p.x; A[2] = dim > 1 ? m * p.y : 0.0; A[3] = dim > 2 ? m * p.z : 0.0;
      A[4] = m * v.x; A[5] = dim > 1 ? m * v.y : 0.0; A[6] = dim > 2 ? m * v.z : 0.0;
      A[7] = m * a.x; A[8] = dim > 1 ? m * a.y : 0.0; A[9] = dim > 2 ? m * a.z : 0.0;
      R::run(A, lane);
      if (myrow >= 0) acc[w][mm][myrow] += A[0];
    }
  }
  __syncthreads();
  for (int q = threadIdx.x; q < nlev * kComVals; q += blockDim.x) {
    const int mm = q / kComVals, r = q % kComVals;
    double t = 0.0;
    for (int ww = 0; ww < kComWarps; ww++) t += acc[ww][mm][r];
    part[(size_t)blockIdx.x * kComMaxLev * kComVals + q] = t;
  }
}

__global__ void com_final_kernel(const double *__restrict__ part, int nblocks, int nlev, double *__restrict__ out)
{
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nlev * kComVals) return;
  double t = 0.0;
  for (int b = 0; b < nblocks; b++) t += part[(size_t)b * kComMaxLev * kComVals + q];
  out[q] = t;
}


__global__ void __launch_bounds__(256)
level_count_kernel(const unsigned *__restrict__ level, size_t n, int nlev,
                   unsigned long long *__restrict__ counts)
{
  __shared__ unsigned int h[64];
  if (threadIdx.x < 64) h[threadIdx.x] = 0;
  __syncthreads();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    unsigned l = level[i];
    atomicAdd(&h[l < (unsigned)nlev ? l : nlev - 1], 1u);
  }
  __syncthreads();
  if (threadIdx.x < nlev && h[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)h[threadIdx.x]);
}

} // namespace

// Applies the current sort permutation to the particle store (see common.cuh).  The
// 4-byte arrays go through the sort's rank scratch, whose buffer they then swap with.
__global__ void bounds_gather_kernel(const uint32_t *__restrict__ seg, uint32_t kpl, int nlev,
                                     uint32_t *__restrict__ out)
{
  const int L = threadIdx.x;
  if (L <= nlev) out[L] = seg[(size_t)L * kpl];
}

int component_read_bounds(expb200_component_s *c, uint32_t kpl, int nlev, int from)
{
  if (nlev + 1 > 40) { set_error("component_read_bounds: %d levels", nlev); return EXPB200_EINVAL; }
  uint32_t tmp[40];
  bounds_gather_kernel<<<1, 64, 0, c->stream>>>(c->seg, kpl, nlev, c->bounds_dev);
  CU_LAUNCH_CHECK();
  CU_TRY(cudaMemcpyAsync(tmp, c->bounds_dev, (size_t)(nlev + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  for (int L = from + 1; L <= nlev; L++) c->lev_start[L] = tmp[L];
  c->part_ok = true;
  c->part_nlev = nlev;
  c->bound_ok_upto = nlev;
  return EXPB200_OK;
}

int component_reorder(expb200_component_s *c)
{
  const size_t n = c->n;
  if (n == 0) return EXPB200_OK;
  cudaStream_t st = c->stream;
  const uint32_t *perm = c->perm[c->cur];
  const unsigned grid = grid_for(n, 256);
  ProfRegion pr(K_REORDER, st, 10);
  // positions: the sorted copy becomes the store
  { rec4 *t = c->posm; c->posm = c->posm_sorted; c->posm_sorted = t; }
  gather_rec_kernel<<<grid, 256, 0, st>>>(c->spare, c->vel, perm, n);
  { rec4 *t = c->vel; c->vel = c->spare; c->spare = t; }
  gather_rec_kernel<<<grid, 256, 0, st>>>(c->spare, c->accp, perm, n);
  { rec4 *t = c->accp; c->accp = c->spare; c->spare = t; }
  auto move_u32 = [&](uint32_t **arr) {
    gather_u32_kernel<<<grid, 256, 0, st>>>(c->rank, *arr, perm, n);
    uint32_t *t = *arr; *arr = c->rank; c->rank = t;
  };
  move_u32(reinterpret_cast<uint32_t **>(&c->level));
  move_u32(reinterpret_cast<uint32_t **>(&c->prev_level));
  move_u32(reinterpret_cast<uint32_t **>(&c->dtreq));
  move_u32(reinterpret_cast<uint32_t **>(&c->scale));
  if (c->identity_order) {
    // slot_of_new[k] = perm[k]
    CU_TRY(cudaMemcpyAsync(c->slot_of, perm, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
  } else {
    move_u32(&c->slot_of);
  }
  invert_kernel<<<grid, 256, 0, st>>>(c->where, c->slot_of, n);
  iota_kernel<<<grid, 256, 0, st>>>(c->perm[c->cur], n);
  CU_LAUNCH_CHECK();
  c->identity_order = false;
  c->never_reordered = false;
  c->sorts_since_reorder = 0;
  c->psort = c->posm;
  return EXPB200_OK;
}

} // namespace expb200

using namespace expb200;

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
extern "C" {

const char *expb200_last_error(void) { return g_err.c_str(); }

int expb200_version(void) { return 100; }

int expb200_device_count(int *count)
{
  REQUIRE(count, "count is NULL");
  CU_TRY(cudaGetDeviceCount(count));
  return EXPB200_OK;
}

int expb200_device_info(int device, char *name, int name_len, int *sm_count, size_t *hbm_bytes)
{
  cudaDeviceProp p;
  CU_TRY(cudaGetDeviceProperties(&p, device));
  if (name && name_len > 0) { strncpy(name, p.name, name_len - 1); name[name_len - 1] = 0; }
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (hbm_bytes) *hbm_bytes = p.totalGlobalMem;
  return EXPB200_OK;
}

int expb200_host_alloc(void **ptr, size_t bytes)
{
  REQUIRE(ptr, "ptr is NULL");
  CU_TRY(cudaHostAlloc(ptr, bytes, cudaHostAllocDefault));
  return EXPB200_OK;
}

int expb200_host_free(void *ptr)
{
  CU_TRY(cudaFreeHost(ptr));
  return EXPB200_OK;
}

int expb200_host_register(void *ptr, size_t bytes)
{
  CU_TRY(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
  return EXPB200_OK;
}

int expb200_host_unregister(void *ptr)
{
  CU_TRY(cudaHostUnregister(ptr));
  return EXPB200_OK;
}

int expb200_component_create(int device, size_t capacity, expb200_component *out)
{
  REQUIRE(out, "out is NULL");
  REQUIRE(capacity > 0 && capacity < 0xfffffff0ull, "capacity %zu out of range", capacity);
  int ndev = 0;
  CU_TRY(cudaGetDeviceCount(&ndev));
  REQUIRE(device >= 0 && device < ndev, "device %d not present (%d visible)", device, ndev);
  DeviceGuard g(device);
  auto *c = new expb200_component_s;
  c->device = device;
  c->cap = capacity;
  CU_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CU_TRY(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  stream_register(c->stream, true);
  CU_TRY(cudaEventCreateWithFlags(&c->copy_done, cudaEventDisableTiming));
  // every per-particle array is allocated in whole tiles of 512 entries: the pipelined force
  // kernel copies sorted positions and the permutation a tile at a time (TMA bulk copies)
  const size_t padcap = (capacity + 511) / 512 * 512;
  CU_TRY(cudaMalloc(&c->posm, padcap * sizeof(rec4)));
  CU_TRY(cudaMalloc(&c->vel, padcap * sizeof(rec4)));
  CU_TRY(cudaMalloc(&c->accp, padcap * sizeof(rec4)));
  CU_TRY(cudaMalloc(&c->posm_sorted, padcap * sizeof(rec4)));
  CU_TRY(cudaMalloc(&c->spare, padcap * sizeof(rec4)));
  CU_TRY(cudaMalloc(&c->slot_of, padcap * sizeof(uint32_t)));
  CU_TRY(cudaMalloc(&c->where, padcap * sizeof(uint32_t)));
  c->psort = c->posm_sorted;
  CU_TRY(cudaMalloc(&c->level, padcap * sizeof(unsigned)));
  CU_TRY(cudaMalloc(&c->prev_level, padcap * sizeof(unsigned)));
  CU_TRY(cudaMalloc(&c->dtreq, padcap * sizeof(float)));
  CU_TRY(cudaMalloc(&c->scale, padcap * sizeof(float)));
  for (int k = 0; k < 2; k++) {
    CU_TRY(cudaMalloc(&c->keys[k], padcap * sizeof(uint32_t)));
    CU_TRY(cudaMalloc(&c->perm[k], padcap * sizeof(uint32_t)));
  }
  CU_TRY(cudaMalloc(&c->rank, padcap * sizeof(uint32_t)));
  c->hist_cap = sort_hist_words(capacity);
  CU_TRY(cudaMalloc(&c->hist, c->hist_cap * sizeof(uint32_t)));
  c->stage_chunk = capacity < kStageChunk ? capacity : kStageChunk;
  CU_TRY(cudaMalloc(&c->stage, 2 * c->stage_chunk * kStageDoubles * sizeof(double)));
  CU_TRY(cudaMalloc(&c->stats, 8 * sizeof(unsigned long long)));
  CU_TRY(cudaMalloc(&c->bounds_dev, 40 * sizeof(uint32_t)));
  CU_TRY(cudaMemsetAsync(c->accp, 0, capacity * sizeof(rec4), c->stream));
  CU_TRY(cudaMemsetAsync(c->level, 0, capacity * sizeof(unsigned), c->stream));
  CU_TRY(cudaMemsetAsync(c->prev_level, 0, capacity * sizeof(unsigned), c->stream));
  CU_TRY(cudaMemsetAsync(c->dtreq, 0, capacity * sizeof(float), c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  *out = c;
  return EXPB200_OK;
}

int expb200_component_destroy(expb200_component c)
{
  if (!c) return EXPB200_OK;
  DeviceGuard g(c->device);
  cudaStreamSynchronize(c->stream);
  cudaStreamSynchronize(c->copy_stream);
  cudaFree(c->posm); cudaFree(c->vel); cudaFree(c->accp); cudaFree(c->posm_sorted);
  cudaFree(c->spare); cudaFree(c->slot_of); cudaFree(c->where);
  cudaFree(c->level); cudaFree(c->prev_level); cudaFree(c->dtreq); cudaFree(c->scale);
  for (int k = 0; k < 2; k++) { cudaFree(c->keys[k]); cudaFree(c->perm[k]); }
  cudaFree(c->hist); cudaFree(c->rank); cudaFree(c->seg); cudaFree(c->stage); cudaFree(c->stats); cudaFree(c->com_part); cudaFree(c->bounds_dev); cudaFree(c->sw_perm);
  cudaEventDestroy(c->copy_done);
  if (c->hstage) cudaFreeHost(c->hstage);
  for (int k = 0; k < 2; k++) if (c->hstage_free[k]) cudaEventDestroy(c->hstage_free[k]);
  stream_register(c->stream, false);
  cudaStreamDestroy(c->stream);
  cudaStreamDestroy(c->copy_stream);
  delete c;
  return EXPB200_OK;
}

size_t expb200_component_count(expb200_component c) { return c ? c->n : 0; }

int expb200_component_set_count(expb200_component c, size_t n)
{
  REQUIRE(c, "component is NULL");
  REQUIRE(n <= c->cap, "n=%zu exceeds capacity %zu", n, c->cap);
  if (n != c->n) {            // a different particle set: storage order = host order again
    c->identity_order = true;
    c->never_reordered = true;
    c->sorts_since_reorder = 0;
    c->sort_kind = 0; c->part_ok = false;
  }
  c->n = n;
  c->pos_ver++; c->lev_ver++;
  return EXPB200_OK;
}

// Upload [first, first+count): H2D on the copy stream into one of two staging
// buffers, pack on the compute stream; the two alternate so that the copy of
// chunk k+1 overlaps the pack of chunk k.
namespace {
// true for ordinary (pageable) host memory: neither cudaHostAlloc'ed nor cudaHostRegister'ed
bool is_pageable(const void *p)
{
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
  return a.type == cudaMemoryTypeUnregistered;
}

// memcpy split over a few host threads: one thread moves ~10 GB/s, the PCIe link 50
void parallel_copy(const std::vector<std::tuple<double *, const double *, size_t>> &jobs)
{
  static const unsigned nthr = [] {
    unsigned h = std::thread::hardware_concurrency();
    return h >= 8 ? 8u : h >= 2 ? h : 1u;
  }();
  size_t total = 0;
  for (auto &j : jobs) total += std::get<2>(j);
  if (nthr == 1 || total < (1u << 16)) {
    for (auto &j : jobs) memcpy(std::get<0>(j), std::get<1>(j), std::get<2>(j) * sizeof(double));
    return;
  }
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nthr; t++)
    th.emplace_back([&, t] {
      for (auto &j : jobs) {
        const size_t n = std::get<2>(j), lo = n * t / nthr, hi = n * (t + 1) / nthr;
        if (hi > lo) memcpy(std::get<0>(j) + lo, std::get<1>(j) + lo, (hi - lo) * sizeof(double));
      }
    });
  for (auto &x : th) x.join();
}
} // namespace

namespace {
int ensure_bounce(expb200_component_s *c, size_t hs)
{
  if (c->hstage) return EXPB200_OK;
  CU_TRY(cudaHostAlloc(&c->hstage, 2 * hs * kStageDoubles * sizeof(double), cudaHostAllocDefault));
  for (int k = 0; k < 2; k++) CU_TRY(cudaEventCreateWithFlags(&c->hstage_free[k], cudaEventDisableTiming));
  return EXPB200_OK;
}
} // namespace

int expb200_component_upload_range(expb200_component c, size_t first, size_t count,
                                   const double *pos, const double *vel, const double *mass)
{
  REQUIRE(c && pos && mass, "NULL argument");
  REQUIRE(first + count <= c->cap, "range [%zu,%zu) exceeds capacity %zu", first, first + count, c->cap);
  DeviceGuard g(c->device);
  const size_t ch = c->stage_chunk;
  // Pageable host arrays (numpy arrays, the std::vectors of the host driver): a cudaMemcpy from
  // them is staged by the driver at ~10 GB/s.  Bounce them through our own pinned buffers,
  // filled by several host threads while the previous chunk is on the wire.
  const bool bounce = count >= 4096 && (is_pageable(pos) || is_pageable(mass) || (vel && is_pageable(vel)));
  const size_t hs = ch < kBounceChunk ? ch : kBounceChunk;      // particles per pinned buffer
  if (bounce) { int rc = ensure_bounce(c, hs); if (rc) return rc; }
  int b = 0, hb = 0;
  for (size_t off = 0; off < count; off += ch, b ^= 1) {
    size_t m = count - off < ch ? count - off : ch;
    double *s = c->stage + (size_t)b * ch * kStageDoubles;
    double *spos = s, *svel = s + 3 * ch, *smass = s + 6 * ch;
    // the pack kernel that last read this staging buffer must be done
    CU_TRY(cudaStreamWaitEvent(c->copy_stream, c->copy_done, 0));
    if (bounce) {
      for (size_t so = 0; so < m; so += hs, hb ^= 1) {
        const size_t sm = m - so < hs ? m - so : hs;
        double *h = c->hstage + (size_t)hb * hs * kStageDoubles;
        CU_TRY(cudaEventSynchronize(c->hstage_free[hb]));     // its last copy has left the host
        std::vector<std::tuple<double *, const double *, size_t>> jobs;
        jobs.emplace_back(h, pos + 3 * (off + so), 3 * sm);
        if (vel) jobs.emplace_back(h + 3 * hs, vel + 3 * (off + so), 3 * sm);
        jobs.emplace_back(h + 6 * hs, mass + off + so, sm);
        parallel_copy(jobs);
        CU_TRY(cudaMemcpyAsync(spos + 3 * so, h, 3 * sm * sizeof(double), cudaMemcpyHostToDevice, c->copy_stream));
        if (vel) CU_TRY(cudaMemcpyAsync(svel + 3 * so, h + 3 * hs, 3 * sm * sizeof(double), cudaMemcpyHostToDevice, c->copy_stream));
        CU_TRY(cudaMemcpyAsync(smass + so, h + 6 * hs, sm * sizeof(double), cudaMemcpyHostToDevice, c->copy_stream));
        CU_TRY(cudaEventRecord(c->hstage_free[hb], c->copy_stream));
      }
    } else {
      CU_TRY(cudaMemcpyAsync(spos, pos + 3 * off, 3 * m * sizeof(double), cudaMemcpyHostToDevice, c->copy_stream));
      if (vel) CU_TRY(cudaMemcpyAsync(svel, vel + 3 * off, 3 * m * sizeof(double), cudaMemcpyHostToDevice, c->copy_stream));
      CU_TRY(cudaMemcpyAsync(smass, mass + off, m * sizeof(double), cudaMemcpyHostToDevice, c->copy_stream));
    }
    cudaEvent_t ev;
    CU_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CU_TRY(cudaEventRecord(ev, c->copy_stream));
    CU_TRY(cudaStreamWaitEvent(c->stream, ev, 0));
    CU_TRY(cudaEventDestroy(ev));
    ProfRegion pr(K_PACK, c->stream);
    pack_kernel<<<grid_for(m, 256), 256, 0, c->stream>>>(c->posm, c->vel, c->accp, spos,
                                                         vel ? svel : nullptr, smass, nullptr,
                                                         first + off, m, 0,
                                                         c->identity_order ? nullptr : c->where);
    CU_LAUNCH_CHECK();
    CU_TRY(cudaEventRecord(c->copy_done, c->stream));
  }
  c->pos_ver++;
  return EXPB200_OK;
}

int expb200_component_upload(expb200_component c, size_t n, const double *pos, const double *vel,
                             const double *mass, const unsigned *level, const float *scale,
                             const double *potext)
{
  REQUIRE(c && pos && mass, "NULL argument");
  REQUIRE(n <= c->cap, "n=%zu exceeds capacity %zu", n, c->cap);
  DeviceGuard g(c->device);
  c->n = n;
  // a full upload defines a new particle set: storage order = host order again
  c->identity_order = true;
  c->never_reordered = true;
  c->sorts_since_reorder = 0;
  c->sort_kind = 0; c->part_ok = false;
  if (n == 0) { c->pos_ver++; c->lev_ver++; return EXPB200_OK; }
  int rc = expb200_component_upload_range(c, 0, n, pos, vel, mass);
  if (rc) return rc;
  CU_TRY(cudaMemsetAsync(c->accp, 0, n * sizeof(rec4), c->stream));
  if (level) CU_TRY(cudaMemcpyAsync(c->level, level, n * sizeof(unsigned), cudaMemcpyHostToDevice, c->stream));
  else fill_u32_kernel<<<grid_for(n, 256), 256, 0, c->stream>>>(c->level, n, 0u);
  if (scale) CU_TRY(cudaMemcpyAsync(c->scale, scale, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  else fill_f32_kernel<<<grid_for(n, 256), 256, 0, c->stream>>>(c->scale, n, -1.0f);
  CU_TRY(cudaMemsetAsync(c->dtreq, 0, n * sizeof(float), c->stream));
  if (potext) {
    // potext rides in vel.w: re-pack chunk-wise with it
    const size_t ch = c->stage_chunk;
    for (size_t off = 0; off < n; off += ch) {
      size_t m = n - off < ch ? n - off : ch;
      double *s = c->stage;
      CU_TRY(cudaStreamSynchronize(c->stream));
      CU_TRY(cudaMemcpy(s, pos + 3 * off, 3 * m * sizeof(double), cudaMemcpyHostToDevice));
      if (vel) CU_TRY(cudaMemcpy(s + 3 * ch, vel + 3 * off, 3 * m * sizeof(double), cudaMemcpyHostToDevice));
      CU_TRY(cudaMemcpy(s + 6 * ch, mass + off, m * sizeof(double), cudaMemcpyHostToDevice));
      CU_TRY(cudaMemcpy(s + 7 * ch, potext + off, m * sizeof(double), cudaMemcpyHostToDevice));
      pack_kernel<<<grid_for(m, 256), 256, 0, c->stream>>>(c->posm, c->vel, c->accp, s,
                                                           vel ? s + 3 * ch : nullptr, s + 6 * ch,
                                                           s + 7 * ch, off, m, 0, nullptr);
      CU_LAUNCH_CHECK();
    }
  }
  CU_TRY(cudaStreamSynchronize(c->stream));
  c->pos_ver++; c->lev_ver++;
  return EXPB200_OK;
}

int expb200_component_sync(expb200_component c)
{
  REQUIRE(c, "component is NULL");
  DeviceGuard g(c->device);
  CU_TRY(cudaStreamSynchronize(c->copy_stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  return EXPB200_OK;
}

int expb200_component_download_range(expb200_component c, size_t first, size_t count,
                                     double *pos, double *vel, double *acc, double *pot)
{
  REQUIRE(c, "component is NULL");
  REQUIRE(first + count <= c->n, "range [%zu,%zu) exceeds count %zu", first, first + count, c->n);
  DeviceGuard g(c->device);
  const size_t ch = c->stage_chunk;
  int b = 0;
  cudaEvent_t unpacked[2], copied[2];
  for (int k = 0; k < 2; k++) {
    CU_TRY(cudaEventCreateWithFlags(&unpacked[k], cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&copied[k], cudaEventDisableTiming));
  }
  // pageable destinations: bounce through the pinned buffers (see upload_range)
  const bool bounce = count >= 4096 && ((pos && is_pageable(pos)) || (vel && is_pageable(vel)) ||
                                        (acc && is_pageable(acc)) || (pot && is_pageable(pot)));
  const size_t hs = ch < kBounceChunk ? ch : kBounceChunk;
  if (bounce) { int rc = ensure_bounce(c, hs); if (rc) return rc; }
  int hb = 0;
  double *pend_h = nullptr;
  size_t pend_off = 0, pend_n = 0;
  cudaEvent_t pend_ev = nullptr;
  auto drain = [&]() -> int {          // the pending sub-chunk: wait for its copy, hand it out
    CU_TRY(cudaEventSynchronize(pend_ev));
    std::vector<std::tuple<double *, const double *, size_t>> jobs;
    if (pos) jobs.emplace_back(pos + 3 * pend_off, pend_h, 3 * pend_n);
    if (vel) jobs.emplace_back(vel + 3 * pend_off, pend_h + 3 * hs, 3 * pend_n);
    if (acc) jobs.emplace_back(acc + 3 * pend_off, pend_h + 6 * hs, 3 * pend_n);
    if (pot) jobs.emplace_back(pot + pend_off, pend_h + 9 * hs, pend_n);
    parallel_copy(jobs);
    pend_n = 0;
    return EXPB200_OK;
  };
  // order after any upload staging still in flight
  CU_TRY(cudaStreamWaitEvent(c->stream, c->copy_done, 0));
  size_t round = 0;
  for (size_t off = 0; off < count; off += ch, b ^= 1, round++) {
    size_t m = count - off < ch ? count - off : ch;
    double *s = c->stage + (size_t)b * ch * kStageDoubles;
    double *spos = s, *svel = s + 3 * ch, *sacc = s + 6 * ch, *spot = s + 9 * ch;
    if (round >= 2) CU_TRY(cudaStreamWaitEvent(c->stream, copied[b], 0));
    ProfRegion pr(K_PACK, c->stream);
    unpack_kernel<<<grid_for(m, 256), 256, 0, c->stream>>>(c->posm, c->vel, c->accp,
                                                           pos ? spos : nullptr, vel ? svel : nullptr,
                                                           acc ? sacc : nullptr, pot ? spot : nullptr,
                                                           first + off, m,
                                                           c->identity_order ? nullptr : c->where);
    CU_LAUNCH_CHECK();
    CU_TRY(cudaEventRecord(unpacked[b], c->stream));
    CU_TRY(cudaStreamWaitEvent(c->copy_stream, unpacked[b], 0));
    if (bounce) {
      // device staging -> pinned buffer (async); the previous sub-chunk is copied out to the
      // caller's pageable arrays by the host threads meanwhile
      for (size_t so = 0; so < m; so += hs, hb ^= 1) {
        const size_t sm = m - so < hs ? m - so : hs;
        double *h = c->hstage + (size_t)hb * hs * kStageDoubles;
        if (pos) CU_TRY(cudaMemcpyAsync(h, spos + 3 * so, 3 * sm * sizeof(double), cudaMemcpyDeviceToHost, c->copy_stream));
        if (vel) CU_TRY(cudaMemcpyAsync(h + 3 * hs, svel + 3 * so, 3 * sm * sizeof(double), cudaMemcpyDeviceToHost, c->copy_stream));
        if (acc) CU_TRY(cudaMemcpyAsync(h + 6 * hs, sacc + 3 * so, 3 * sm * sizeof(double), cudaMemcpyDeviceToHost, c->copy_stream));
        if (pot) CU_TRY(cudaMemcpyAsync(h + 9 * hs, spot + so, sm * sizeof(double), cudaMemcpyDeviceToHost, c->copy_stream));
        CU_TRY(cudaEventRecord(c->hstage_free[hb], c->copy_stream));
        if (pend_n) { int rc = drain(); if (rc) return rc; }
        pend_h = h; pend_off = off + so; pend_n = sm; pend_ev = c->hstage_free[hb];
      }
    } else {
    if (pos) CU_TRY(cudaMemcpyAsync(pos + 3 * off, spos, 3 * m * sizeof(double), cudaMemcpyDeviceToHost, c->copy_stream));
    if (vel) CU_TRY(cudaMemcpyAsync(vel + 3 * off, svel, 3 * m * sizeof(double), cudaMemcpyDeviceToHost, c->copy_stream));
    if (acc) CU_TRY(cudaMemcpyAsync(acc + 3 * off, sacc, 3 * m * sizeof(double), cudaMemcpyDeviceToHost, c->copy_stream));
    if (pot) CU_TRY(cudaMemcpyAsync(pot + off, spot, m * sizeof(double), cudaMemcpyDeviceToHost, c->copy_stream));
    }
    CU_TRY(cudaEventRecord(copied[b], c->copy_stream));
  }
  if (pend_n) { int rc = drain(); if (rc) return rc; }
  CU_TRY(cudaStreamSynchronize(c->copy_stream));
  // later uploads must not overwrite staging before these copies finished (they have)
  CU_TRY(cudaEventRecord(c->copy_done, c->stream));
  for (int k = 0; k < 2; k++) { cudaEventDestroy(unpacked[k]); cudaEventDestroy(copied[k]); }
  return EXPB200_OK;
}

int expb200_component_download(expb200_component c, double *pos, double *vel, double *acc,
                               double *pot, unsigned *level, float *dtreq)
{
  REQUIRE(c, "component is NULL");
  DeviceGuard g(c->device);
  if (c->n == 0) return EXPB200_OK;
  if (pos || vel || acc || pot) {
    int rc = expb200_component_download_range(c, 0, c->n, pos, vel, acc, pot);
    if (rc) return rc;
  }
  const uint32_t *map = c->identity_order ? nullptr : c->where;
  if (level) {
    gather_u32_kernel<<<grid_for(c->n, 256), 256, 0, c->stream>>>(c->rank, c->level, map, c->n);
    CU_TRY(cudaMemcpyAsync(level, c->rank, c->n * sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
  }
  if (dtreq) {
    gather_u32_kernel<<<grid_for(c->n, 256), 256, 0, c->stream>>>(
        c->rank, reinterpret_cast<const uint32_t *>(c->dtreq), map, c->n);
    CU_TRY(cudaMemcpyAsync(dtreq, c->rank, c->n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  }
  CU_TRY(cudaStreamSynchronize(c->stream));
  return EXPB200_OK;
}

int expb200_component_set_center(expb200_component c, const double center[3], double rtrunc)
{
  REQUIRE(c && center, "NULL argument");
  for (int k = 0; k < 3; k++) c->center[k] = center[k];
  c->rtrunc = rtrunc;
  c->pos_ver++;              // radial cells depend on the centre
  return EXPB200_OK;
}

int expb200_component_zero_pot_acc(expb200_component c, int minlev)
{
  REQUIRE(c, "component is NULL");
  if (c->n == 0) return EXPB200_OK;
  DeviceGuard g(c->device);
  ProfRegion pr(K_ZERO, c->stream);
  zero_kernel<<<grid_for(c->n, 256), 256, 0, c->stream>>>(c->accp, c->level, c->n, minlev);
  CU_LAUNCH_CHECK();
  return EXPB200_OK;
}

// ~ the per-level incr_velocity / incr_position calls of one sub-step (src/step.cc:215-237 and
// :253-262) in one pass: every particle of level M in [mlevel_lo, nlev) gets vel += acc dt_kick[M]
// (if dt_kick != NULL), then pos += vel dt_drift[M] (if dt_drift != NULL).
int expb200_move_levels(expb200_component c, const double *dt_kick, const double *dt_drift, int nlev,
                        int mlevel_lo)
{
  REQUIRE(c, "component is NULL");
  REQUIRE(nlev >= 1 && nlev <= 32 && mlevel_lo >= 0, "nlev=%d mlevel_lo=%d", nlev, mlevel_lo);
  if (c->n == 0 || (!dt_kick && !dt_drift)) return EXPB200_OK;
  DeviceGuard g(c->device);
  LevelDt D;
  D.kick = dt_kick != nullptr; D.drift = dt_drift != nullptr;
  for (int M = 0; M < 32; M++) {
    D.k[M] = dt_kick && M < nlev ? dt_kick[M] : 0.0;
    D.d[M] = dt_drift && M < nlev ? dt_drift[M] : 0.0;
  }
  {
    // a pass over the store in its own order: walking only the active tail of the sorted
    // permutation was measured slower (disk+halo, 1e8: the permutation is in the cell order of
    // whichever force binned last, so every record access becomes a random 32-byte sector)
    ProfRegion pr(dt_drift ? K_DRIFT : K_KICK, c->stream);
    move_levels_kernel<<<grid_for(c->n, 256), 256, 0, c->stream>>>(c->posm, c->vel, c->accp, c->level, c->n, D,
                                                                   mlevel_lo, nlev);
    CU_LAUNCH_CHECK();
  }
  if (dt_drift) c->pos_ver++;
  return EXPB200_OK;
}

int expb200_incr_position(expb200_component c, double dt, int mlevel)
{
  REQUIRE(c, "component is NULL");
  if (c->n == 0) return EXPB200_OK;
  DeviceGuard g(c->device);
  ProfRegion pr(K_DRIFT, c->stream);
  drift_kernel<<<grid_for(c->n, 256), 256, 0, c->stream>>>(c->posm, c->vel, c->level, c->n, dt, mlevel, 3);
  CU_LAUNCH_CHECK();
  c->pos_ver++;
  return EXPB200_OK;
}

int expb200_incr_velocity(expb200_component c, double dt, int mlevel)
{
  REQUIRE(c, "component is NULL");
  if (c->n == 0) return EXPB200_OK;
  DeviceGuard g(c->device);
  ProfRegion pr(K_KICK, c->stream);
  kick_kernel<<<grid_for(c->n, 256), 256, 0, c->stream>>>(c->vel, c->accp, c->level, c->n, dt, mlevel, 3);
  CU_LAUNCH_CHECK();
  return EXPB200_OK;
}

int expb200_adjust_multistep_level(expb200_component c, const expb200_level_params *p,
                                   expb200_level_stats *stats)
{
  REQUIRE(c && p, "NULL argument");
  REQUIRE(p->multistep >= 0 && p->multistep < 32, "multistep=%d out of range", p->multistep);
  DeviceGuard g(c->device);
  LevelArgs A;
  A.dtime = p->dtime; A.fS = p->dynfracS; A.fD = p->dynfracD; A.fV = p->dynfracV;
  A.fA = p->dynfracA; A.fP = p->dynfracP;
  A.multistep = p->multistep; A.shiftlevl = p->shiftlevl; A.mfirst = p->mfirst_mdrft;
  A.minlev = p->minlev; A.noswitch = p->noswitch; A.reset_dtreq = p->reset_dtreq;
  A.apply = p->apply; A.dim = p->dim > 0 ? p->dim : 3;
  // The host libm's answer for exact powers of two (what the CPU path computes)
  for (int k = 0; k < 64; k++) {
    volatile double x = std::ldexp(1.0, k);
    A.pow2lev[k] = (int)std::floor(std::log(x) / std::log(2.0));
  }
  unsigned long long init[8] = {0, 0, 0, 0, 0x7fefffffffffffffULL, 0, 0, 0};
  CU_TRY(cudaMemcpyAsync(c->stats, init, sizeof init, cudaMemcpyHostToDevice, c->stream));
  if (c->n) {
    // remember the levels before the update: multistep_update moves each switching
    // particle's contribution from prev_level to level (src/multistep.cc:198-206)
    CU_TRY(cudaMemcpyAsync(c->prev_level, c->level, c->n * sizeof(unsigned), cudaMemcpyDeviceToDevice, c->stream));
    ProfRegion pr(K_LEVELS, c->stream);
    levels_kernel<<<grid_for(c->n, 256), 256, 0, c->stream>>>(c->vel, c->accp, c->scale, c->level,
                                                              c->dtreq, c->n, A, c->stats);
    CU_LAUNCH_CHECK();
  }
  unsigned long long out[8];
  CU_TRY(cudaMemcpyAsync(out, c->stats, sizeof out, cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  if (out[1]) {
    c->lev_ver++;
    // only particles of level >= minlev were examined and none went below mfirst[mdrft]: the
    // level boundaries of the sorted permutation up to the lower of the two still hold
    const int lo = A.minlev < A.mfirst ? A.minlev : A.mfirst;
    if (lo < c->bound_ok_upto) c->bound_ok_upto = lo < 0 ? 0 : lo;
  }
  if (stats) {
    stats->examined = out[0]; stats->switched = out[1]; stats->offlo = out[2]; stats->offhi = out[3];
    double mn, mx;
    memcpy(&mn, &out[4], 8); memcpy(&mx, &out[5], 8);
    stats->mindt = mn; stats->maxdt = mx;
  }
  return EXPB200_OK;
}

int expb200_component_set_levels(expb200_component c, const unsigned *level)
{
  REQUIRE(c && level, "NULL argument");
  DeviceGuard g(c->device);
  if (c->identity_order) {
    CU_TRY(cudaMemcpyAsync(c->level, level, c->n * sizeof(unsigned), cudaMemcpyHostToDevice, c->stream));
  } else {
    CU_TRY(cudaMemcpyAsync(c->rank, level, c->n * sizeof(unsigned), cudaMemcpyHostToDevice, c->stream));
    scatter_u32_kernel<<<grid_for(c->n, 256), 256, 0, c->stream>>>(c->level, c->rank, c->where, c->n);
  }
  CU_TRY(cudaMemcpyAsync(c->prev_level, c->level, c->n * sizeof(unsigned), cudaMemcpyDeviceToDevice, c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  c->lev_ver++;
  c->part_ok = false;
  return EXPB200_OK;
}

int expb200_component_set_acc_pot(expb200_component c, const double *acc, const double *pot)
{
  // test / restart hook: install acc[n][3], pot[n] (e.g. to drive level assignment)
  REQUIRE(c && acc && pot, "NULL argument");
  DeviceGuard g(c->device);
  std::string tmp;
  tmp.resize(c->n * sizeof(rec4));
  rec4 *h = reinterpret_cast<rec4 *>(&tmp[0]);
  for (size_t i = 0; i < c->n; i++) h[i] = rec4{acc[3 * i], acc[3 * i + 1], acc[3 * i + 2], pot[i]};
  if (c->identity_order) {
    CU_TRY(cudaMemcpyAsync(c->accp, h, c->n * sizeof(rec4), cudaMemcpyHostToDevice, c->stream));
  } else {
    CU_TRY(cudaMemcpyAsync(c->spare, h, c->n * sizeof(rec4), cudaMemcpyHostToDevice, c->stream));
    scatter_rec_kernel<<<grid_for(c->n, 256), 256, 0, c->stream>>>(c->accp, c->spare, c->where, c->n);
  }
  CU_TRY(cudaStreamSynchronize(c->stream));
  return EXPB200_OK;
}

// Device-side timing of a region that may span both streams of the component:
// slot in [0,8).  record joins the copy stream into the compute stream, records there
// and makes the copy stream wait for it, so everything issued before/after on either
// stream falls on the right side of the event.
static cudaEvent_t g_tev[8] = {};

int expb200_component_event_record(expb200_component c, int slot)
{
  REQUIRE(c && slot >= 0 && slot < 8, "bad argument");
  DeviceGuard g(c->device);
  if (!g_tev[slot]) CU_TRY(cudaEventCreate(&g_tev[slot]));
  cudaEvent_t j;
  CU_TRY(cudaEventCreateWithFlags(&j, cudaEventDisableTiming));
  CU_TRY(cudaEventRecord(j, c->copy_stream));
  CU_TRY(cudaStreamWaitEvent(c->stream, j, 0));
  CU_TRY(cudaEventDestroy(j));
  CU_TRY(cudaEventRecord(g_tev[slot], c->stream));
  CU_TRY(cudaStreamWaitEvent(c->copy_stream, g_tev[slot], 0));
  return EXPB200_OK;
}

int expb200_component_event_elapsed(expb200_component c, int slot_a, int slot_b, double *ms)
{
  REQUIRE(c && ms && slot_a >= 0 && slot_a < 8 && slot_b >= 0 && slot_b < 8, "bad argument");
  REQUIRE(g_tev[slot_a] && g_tev[slot_b], "event slot never recorded");
  DeviceGuard g(c->device);
  CU_TRY(cudaEventSynchronize(g_tev[slot_b]));
  float f = 0.f;
  CU_TRY(cudaEventElapsedTime(&f, g_tev[slot_a], g_tev[slot_b]));
  *ms = f;
  return EXPB200_OK;
}

// ~ Component::fix_positions_cuda (src/cudaComponent.cu:800-960); arithmetic of
// fix_positions_thread (src/Component.cc:3280-3352) without the consp/tidal escape branch.
int expb200_component_fix_positions(expb200_component c, int mlevel, int multistep, int dim, double *sums,
                                    void *nccl_comm)
{
  REQUIRE(c && sums, "NULL argument");
  REQUIRE(multistep >= 0 && multistep < kComMaxLev, "multistep=%d outside [0,%d)", multistep, kComMaxLev);
  REQUIRE(mlevel >= 0 && mlevel <= multistep, "mlevel=%d outside [0,%d]", mlevel, multistep);
  REQUIRE(dim >= 1 && dim <= 3, "dim=%d", dim);
  DeviceGuard g(c->device);
  const int nlev = multistep + 1;
  int nblocks = (int)std::min<size_t>((c->n + kComWarps * 32 - 1) / (kComWarps * 32), (size_t)kSMsB200 * 8);
  if (nblocks < 1) nblocks = 1;
  const size_t need = ((size_t)nblocks + 1) * kComMaxLev * kComVals;
  if (c->com_part_cap < need) {
    if (c->com_part) cudaFree(c->com_part);
    c->com_part = nullptr; c->com_part_cap = 0;
    CU_TRY(cudaMalloc(&c->com_part, need * sizeof(double)));
    c->com_part_cap = need;
  }
  double *out = c->com_part + (size_t)nblocks * kComMaxLev * kComVals;
  {
    ProfRegion pr(K_OTHER, c->stream, 2);
    com_sum_kernel<<<nblocks, kComWarps * 32, 0, c->stream>>>(c->posm, c->vel, c->accp, c->level, c->n, mlevel,
                                                             nlev, dim, c->center[0], c->center[1], c->center[2],
                                                             c->rtrunc * c->rtrunc, c->com_part);
    com_final_kernel<<<div_up(nlev * kComVals, 64), 64, 0, c->stream>>>(c->com_part, nblocks, nlev, out);
    CU_LAUNCH_CHECK();
  }
  if (nccl_comm) {
    int rc = nccl_allreduce_sum_f64(nccl_comm, out, (size_t)nlev * kComVals, c->stream);
    if (rc) return rc;
  }
  double h[kComMaxLev * kComVals];
  CU_TRY(cudaMemcpyAsync(h, out, (size_t)nlev * kComVals * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  for (int mm = mlevel; mm < nlev; mm++)
    for (int r = 0; r < kComVals; r++) sums[mm * kComVals + r] = h[mm * kComVals + r];
  return EXPB200_OK;
}

int expb200_component_sort_counters(expb200_component c, uint64_t counters[4])
{
  REQUIRE(c && counters, "NULL argument");
  for (int k = 0; k < 4; k++) counters[k] = c->sort_counters[k];
  return EXPB200_OK;
}

int expb200_component_level_counts(expb200_component c, int multistep, uint64_t *counts)
{
  REQUIRE(c && counts, "NULL argument");
  REQUIRE(multistep >= 0 && multistep < 63, "multistep out of range");
  DeviceGuard g(c->device);
  unsigned long long *d = nullptr;
  CU_TRY(cudaMalloc(&d, 64 * sizeof(unsigned long long)));
  CU_TRY(cudaMemsetAsync(d, 0, 64 * sizeof(unsigned long long), c->stream));
  if (c->n) level_count_kernel<<<grid_for(c->n, 256), 256, 0, c->stream>>>(c->level, c->n, multistep + 1, d);
  unsigned long long h[64];
  CU_TRY(cudaMemcpyAsync(h, d, sizeof h, cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  cudaFree(d);
  for (int k = 0; k <= multistep; k++) counts[k] = h[k];
  return EXPB200_OK;
}

} // extern "C"
