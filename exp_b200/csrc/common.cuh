// common.cuh -- shared host/device utilities of libexpb200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/expb200.h"

namespace expb200 {

constexpr int kSMsB200 = 148;          // B200: 148 SMs on two dies
constexpr int kMaxL    = 12;           // template instantiations cover lmax <= kMaxL
constexpr int kWarp    = 32;

// src/expand.H:130 -- added to every radius by the reference
#define EXPB200_DSMALL 1.0e-16
// src/Basis.cc:7 -- clamp for the Legendre derivative
#define EXPB200_MINEPS (3.0 * 2.220446049250313e-16)

// ---- error plumbing: no exceptions across the C ABI ------------------------
void set_error(const char *fmt, ...);
int  cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define CU_TRY(expr)                                                           \
  do {                                                                         \
    cudaError_t _e = (expr);                                                   \
    if (_e != cudaSuccess) return ::expb200::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define CU_LAUNCH_CHECK() CU_TRY(cudaGetLastError())

#define REQUIRE(cond, ...)                                                     \
  do {                                                                         \
    if (!(cond)) { ::expb200::set_error(__VA_ARGS__); return EXPB200_EINVAL; } \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// Registry of live component streams: a Sphere remembers the stream it last worked on and
// must not touch it after that component was destroyed.
void stream_register(cudaStream_t st, bool live);
bool stream_is_live(cudaStream_t st);

inline unsigned div_up(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

// Grid size for a grid-stride kernel: enough CTAs to fill the 148 SMs a few
// times over, never more than the work needs.
inline unsigned grid_for(size_t n, int threads, int ctas_per_sm = 8)
{
  size_t need = (n + threads - 1) / threads;
  size_t cap = (size_t)kSMsB200 * ctas_per_sm;
  if (need < 1) need = 1;
  return (unsigned)(need < cap ? need : cap);
}

// ---- 32-byte particle records ---------------------------------------------
// One particle attribute group = one 32-byte sector: {x,y,z,m}, {vx,vy,vz,potext},
// {ax,ay,az,pot}.  A gather through the sort permutation then costs exactly one
// sector per group and a streaming pass is a fully coalesced 2x LDG.128.
struct __align__(32) rec4 { double x, y, z, w; };

#if defined(__CUDACC__)
__device__ __forceinline__ rec4 ld_rec(const rec4 *p)
{
  const double2 *q = reinterpret_cast<const double2 *>(p);
  double2 a = __ldg(q), b = __ldg(q + 1);
  return rec4{a.x, a.y, b.x, b.y};
}
__device__ __forceinline__ rec4 ld_rec_rw(const rec4 *p)   // plain load (data is written by this kernel family)
{
  const double2 *q = reinterpret_cast<const double2 *>(p);
  double2 a = q[0], b = q[1];
  return rec4{a.x, a.y, b.x, b.y};
}
__device__ __forceinline__ void st_rec(rec4 *p, const rec4 &v)
{
  double2 *q = reinterpret_cast<double2 *>(p);
  q[0] = make_double2(v.x, v.y);
  q[1] = make_double2(v.z, v.w);
}
// FP64 tensor-core tile product D(8x8) += A(8x4) B(4x8): lane (g = lane/4, t = lane%4) holds
// a = A[g][t], b = B[t][g], d0 = D[g][2t], d1 = D[g][2t+1].
__device__ __forceinline__ void dmma_m8n8k4(double &d0, double &d1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
#endif

#if defined(__CUDACC__)
// ---------------------------------------------------------------------------
// Warp reduce-scatter of N per-lane values: five butterfly steps, each lane keeps one
// half of its array and sends the other, so that afterwards lane L holds the full
// 32-lane sums of RS<N>::FIN rows (rs_row tells which).  N + ... shuffles instead of
// 5 N, and the summation tree is fixed.
// ---------------------------------------------------------------------------
template <int N, int S> struct RS {
  static constexpr int H = (N + 1) / 2;
  static constexpr int FIN = RS<H, S + 1>::FIN;
  static __device__ __forceinline__ void run(double *v, int lane)
  {
    const bool up = (lane >> (4 - S)) & 1;
#pragma unroll
    for (int i = 0; i < H; i++) {
      const double lo = v[i];
      const double hi = (i + H < N) ? v[i + H] : 0.0;
      const double send = up ? lo : hi;
      const double keep = up ? hi : lo;
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16 >> S);
    }
    RS<H, S + 1>::run(v, lane);
  }
  // original index of entry i held by `lane` after the remaining steps S..4 (or -1)
  static __device__ __forceinline__ int row(int lane, int i)
  {
    int idx = RS<H, S + 1>::row(lane, i);
    if (idx < 0) return -1;
    if ((lane >> (4 - S)) & 1) idx += H;
    return idx < N ? idx : -1;
  }
};
template <int N> struct RS<N, 5> {
  static constexpr int FIN = N;
  static __device__ __forceinline__ void run(double *, int) {}
  static __device__ __forceinline__ int row(int, int i) { return i < N ? i : -1; }
};

#endif

} // namespace expb200

struct expb200_component_s;
namespace expb200 {
// Physically re-orders the particle store into the order of the component's current sort
// (perm[cur]); afterwards the permutation is the identity and psort == posm.
// lev_start[L] = seg[L * kpl] for L in (from, nlev] -> host (synchronises the component's stream);
// the level partition of perm[cur] is then known exactly.
int component_read_bounds(expb200_component_s *c, uint32_t kpl, int nlev, int from);
int component_reorder(expb200_component_s *c);
constexpr int kReorderEvery = 24;      // sorts between physical re-orderings: the force pass scatters its
                                       // results through the permutation and slows as the store's order
                                       // ages (7.3 -> 8.9 ms per step over 40 steps at 1e8 bodies); a
                                       // re-ordering costs 8.5 ms; 96-step averages: every 64 sorts
                                       // 13.3 ms/step, 32: 12.2, 24: 12.0, 16: 11.9-12.0, 8: 12.4
constexpr int kReorderEveryMulti = 16; // the same with multistep levels (tail sorts count): the level-filtered
                                       // passes (kick, drift, zero) stream best over a level-major store
}

// ---- the device-resident particle store ------------------------------------
struct expb200_component_s {
  int device = 0;
  size_t cap = 0, n = 0;
  cudaStream_t stream = nullptr;        // compute stream
  cudaStream_t copy_stream = nullptr;   // H2D/D2H pipelining
  cudaEvent_t  copy_done = nullptr;

  expb200::rec4 *posm = nullptr;        // {x,y,z,mass}
  expb200::rec4 *vel = nullptr;         // {vx,vy,vz,potext}
  expb200::rec4 *accp = nullptr;        // {ax,ay,az,pot}
  expb200::rec4 *posm_sorted = nullptr; // posm in the order of the current sort (scratch copy)
  expb200::rec4 *spare = nullptr;       // scratch for re-ordering vel/accp
  // Storage order.  Particles are stored in "epoch order": the radial-cell order of the
  // last physical re-ordering (component_reorder), so that the per-step binning and every
  // access through the sort permutation is nearly sequential.  Host slot i lives at
  // storage index where[i]; slot_of is the inverse.  identity_order: storage == host order.
  uint32_t *slot_of = nullptr, *where = nullptr;
  bool identity_order = true;
  bool never_reordered = true;
  int sorts_since_reorder = 0;
  const expb200::rec4 *psort = nullptr; // positions in sorted order: posm_sorted, or posm itself
  bool psort_pending = false;           // posm_sorted is still to be written by the first deposit sweep
                                        // right after a re-ordering
  unsigned *level = nullptr;
  unsigned *prev_level = nullptr;      // level before the last adjust_multistep_level
  float *dtreq = nullptr;
  float *scale = nullptr;

  double center[3] = {0, 0, 0};
  double rtrunc = 1.0e20;

  // staging for (un)packing host AoS arrays
  double *stage = nullptr;              // 7*chunk doubles
  size_t stage_chunk = 0;
  // pinned bounce buffers (2 x 7 x 256K doubles) for uploads from pageable host arrays, and the
  // event that says buffer b's host->device copy has completed (allocated on first use)
  double *hstage = nullptr;
  cudaEvent_t hstage_free[2] = {nullptr, nullptr};

  // sort state: perm[cur] = particle slots ordered by key, keys[cur] sorted keys
  uint32_t *keys[2] = {nullptr, nullptr};
  uint32_t *perm[2] = {nullptr, nullptr};
  uint32_t *hist = nullptr;             // radix histograms
  uint32_t *rank = nullptr;             // sorted position of every item (sort scratch)
  size_t hist_cap = 0;
  uint32_t *seg = nullptr;              // seg[k] = first sorted position with key >= k
  size_t seg_cap = 0;
  int cur = 0;
  int sort_kind = 0;                    // 0 none, 1 level only, 2 (level, radial cell) of sort_owner
  const void *sort_owner = nullptr;
  uint32_t keys_per_level = 1;
  uint64_t pos_ver = 1, lev_ver = 1, sort_pos_ver = 0, sort_lev_ver = 0;
  // what else the cells of the current sort depend on: the expansion origin (the centre of the
  // force's owning component, which need not be this one) and the owner's orientation version
  double sort_origin[3] = {0, 0, 0};
  uint64_t sort_aux_ver = 0;
  // Level partition of perm[cur] (multistep): positions [lev_start[L], lev_start[L+1]) hold the
  // particles of level L.  Only the levels that are active on a sub-step move and change level
  // (src/multistep.cc:194, src/step.cc:210-262), so a sub-step re-bins the suffix of its active
  // levels and leaves the rest of the permutation alone (sphere.cu / cylinder.cu: refresh).
  bool part_ok = false;                 // perm[cur] is partitioned by level
  int part_nlev = 0;
  int bound_ok_upto = 0;                // lev_start[L] is exact for L <= bound_ok_upto
  int sort_from = 0;                    // the cell order of the current sort holds for levels >= sort_from
  uint32_t lev_start[40] = {};          // host copy of the boundaries
  uint32_t *bounds_dev = nullptr;       // 40 words
  uint64_t sort_counters[4] = {0, 0, 0, 0};   // full sorts, tail sorts, items of each (diagnostics)
  uint32_t *sw_perm = nullptr;          // permutation scratch of the level-switch update
  size_t sw_cap = 0;

  // multistep statistics scratch
  unsigned long long *stats = nullptr;  // device, 8 words
  // centre-of-mass sums scratch (expb200_component_fix_positions)
  double *com_part = nullptr;
  size_t com_part_cap = 0;
};
