// sort.cu -- stable LSD radix sort of (key -> particle slot) pairs and segment
// offsets.  Replaces the reference's thrust::sort_by_key / lower_bound level
// sort (src/cudaComponent.cu:383-420, 490-526) and additionally bins particles
// by radial cell, which is what makes the deterministic segmented coefficient
// reduction and the shared-row table gathers possible.
//
// Stability matters: the summation order of the coefficient reduction is the
// sorted order, so a stable sort of keys computed from the particle state makes
// every result a deterministic function of that state.
//
// Shape of one pass (digits up to 11 bits, so the (level, radial cell) keys of a
// single-level sphere are ONE pass):
//   hist    : per-CTA digit histogram of its 8192-key tile          [n*4 B read]
//   scan1/2 : two-level exclusive scan over hist[cta][digit]        [tiny, parallel]
//   rank    : each warp ranks its 512 consecutive keys with match_any (stable) against
//             per-warp digit counters in shared memory               [n*(4 R + 4 W) B]
//   apply   : key, slot and -- in the last pass -- the particle's 32-byte position
//             record go to their sorted position                    [n*(40 R + 40 W) B]
#include "sort.cuh"

namespace expb200 {

namespace {

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kItemsPerLane = 32;
constexpr int kItemsPerWarp = 32 * kItemsPerLane;        // 512
constexpr int kSortTile = kSortWarps * kItemsPerWarp;    // 4096 keys per CTA
constexpr int kMaxDigitBits = 11;
constexpr int kMaxRadix = 1 << kMaxDigitBits;
constexpr int kScanGroups = 32;

__device__ __forceinline__ uint32_t digit_of(uint32_t key, int shift, uint32_t mask)
{
  return (key >> shift) & mask;
}

// Per-CTA digit histogram, hist[cta][digit].
__global__ void __launch_bounds__(kSortThreads)
sort_hist_kernel(const uint32_t *__restrict__ keys, size_t n, int shift, uint32_t mask,
                 uint32_t *__restrict__ hist)
{
  __shared__ uint32_t h[kMaxRadix];
  const uint32_t R = mask + 1;
  for (uint32_t d = threadIdx.x; d < R; d += kSortThreads) h[d] = 0;
  __syncthreads();
  const size_t base = (size_t)blockIdx.x * kSortTile;
#pragma unroll 8
  for (int it = 0; it < kSortTile / kSortThreads; it++) {
    size_t i = base + (size_t)it * kSortThreads + threadIdx.x;
    if (i < n) atomicAdd(&h[digit_of(keys[i], shift, mask)], 1u);   // integer: order-independent
  }
  __syncthreads();
  uint32_t *out = hist + (size_t)blockIdx.x * R;
  for (uint32_t d = threadIdx.x; d < R; d += kSortThreads) out[d] = h[d];
}

// Level 1: within each group of `gs` consecutive CTAs, exclusive prefix over the CTAs for
// every digit (in place) and the group's total per digit.
__global__ void __launch_bounds__(256)
sort_scan1_kernel(uint32_t *__restrict__ hist, uint32_t *__restrict__ gtot, uint32_t R,
                  unsigned nctas, unsigned gs)
{
  const uint32_t d = blockIdx.x * 256 + threadIdx.x;
  if (d >= R) return;
  const unsigned g = blockIdx.y;
  unsigned c0 = g * gs, c1 = c0 + gs < nctas ? c0 + gs : nctas;
  uint32_t run = 0;
  for (unsigned c = c0; c < c1; c += 8) {          // loads batched: they do not depend on the stores
    uint32_t v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = c + k < c1 ? hist[(size_t)(c + k) * R + d] : 0u;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      if (c + k < c1) hist[(size_t)(c + k) * R + d] = run;
      run += v[k];
    }
  }
  gtot[(size_t)g * R + d] = run;
}

// Level 2a: exclusive prefix over the groups for every digit (in place) and the digit totals.
__global__ void __launch_bounds__(256)
sort_scan2a_kernel(uint32_t *__restrict__ gtot, uint32_t *__restrict__ dtot, uint32_t R)
{
  const uint32_t d = blockIdx.x * 256 + threadIdx.x;
  if (d >= R) return;
  uint32_t v[kScanGroups];
#pragma unroll
  for (int g = 0; g < kScanGroups; g++) v[g] = gtot[(size_t)g * R + d];
  uint32_t run = 0;
#pragma unroll
  for (int g = 0; g < kScanGroups; g++) {
    gtot[(size_t)g * R + d] = run;
    run += v[g];
  }
  dtot[d] = run;
}

// Level 2b (one CTA): exclusive scan of the digit totals across digits -> dbase[d], the
// first output position of digit d.  With seg != NULL (single-pass sort: digit == key)
// also seg[k] = first position of key k.
__global__ void __launch_bounds__(1024)
sort_scan2b_kernel(const uint32_t *__restrict__ dtot, uint32_t *__restrict__ dbase, uint32_t R,
                   uint32_t *__restrict__ seg, uint32_t nkeys, uint32_t n)
{
  __shared__ uint32_t warp_sums[32];
  const int t = threadIdx.x;
  const uint32_t dpt = (R + 1023) / 1024;             // digits per thread (contiguous)
  uint32_t tot[2] = {0, 0};
  for (uint32_t j = 0; j < dpt; j++) {
    uint32_t d = t * dpt + j;
    if (d < R) tot[j] = dtot[d];
  }
  const uint32_t s = tot[0] + tot[1];
  uint32_t v = s;
  const int lane = t & 31, w = t >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t u = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += u;
  }
  if (lane == 31) warp_sums[w] = v;
  __syncthreads();
  if (w == 0) {
    uint32_t x = warp_sums[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t u = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += u;
    }
    warp_sums[lane] = x;
  }
  __syncthreads();
  uint32_t base = v - s + (w ? warp_sums[w - 1] : 0u);
  for (uint32_t j = 0; j < dpt; j++) {
    uint32_t d = t * dpt + j;
    if (d >= R) break;
    dbase[d] = base;
    if (seg && d < nkeys) seg[d] = base;
    base += tot[j];
  }
  if (seg && t == 0) seg[nkeys] = n;
}

// Stable ranking.  Each warp owns 512 consecutive keys of the CTA tile (kept in registers)
// and walks them in order, 32 at a time; __match_any_sync groups equal digits, the lowest
// lane of a group is its leader and advances the digit's counter with one shared-memory
// atomic whose old value is the group's base.  Output: rank[i] = sorted position of item i
// (a coalesced write); the data movement is sort_apply_kernel's.
// DIRECT: no payload rides on this pass, so the key and the slot go straight to their sorted
// positions from here (4-byte scatters into runs of equal digits) and sort_apply_kernel, with its
// write and re-read of the ranks and second read of the keys, is skipped.
template <bool DIRECT>
__global__ void __launch_bounds__(kSortThreads, 3)
sort_rank_kernel(const uint32_t *__restrict__ keys_in, uint32_t *__restrict__ rank, size_t n,
                 int shift, uint32_t mask, const uint32_t *__restrict__ hist,
                 const uint32_t *__restrict__ gtot, const uint32_t *__restrict__ dbase, unsigned gs,
                 const uint32_t *__restrict__ perm_in, uint32_t *__restrict__ keys_out,
                 uint32_t *__restrict__ perm_out)
{
  extern __shared__ __align__(16) uint32_t wcount_all[];             // [kSortWarps][R]
  const uint32_t R = mask + 1;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  uint32_t *wcount = wcount_all + (size_t)w * R;

  for (uint32_t d = lane; d < R; d += 32) wcount[d] = 0;
  __syncwarp();

  // 32-bit indexing (capacity < 2^32): one base per array, compile-time offsets per item
  const uint32_t wbase = blockIdx.x * (uint32_t)kSortTile + (uint32_t)w * kItemsPerWarp + lane;
  const uint32_t nn = (uint32_t)n;
  const uint32_t nleft = wbase < nn ? nn - wbase : 0u;      // item `it` is valid iff it*32 < nleft
  const uint32_t *kin = keys_in + wbase;

  // pass A: load the warp's keys, per-warp digit counts (fire-and-forget atomics)
  uint32_t dig[kItemsPerLane];
#pragma unroll
  for (int it = 0; it < kItemsPerLane; it++)
    dig[it] = (uint32_t)(it * 32) < nleft ? digit_of(kin[it * 32], shift, mask) : 0xffffffffu;
#pragma unroll
  for (int it = 0; it < kItemsPerLane; it++) {
    uint32_t grp = __match_any_sync(0xffffffffu, dig[it]);
    if (dig[it] != 0xffffffffu && (grp & lt) == 0) atomicAdd(&wcount[dig[it]], (uint32_t)__popc(grp));
  }
  __syncthreads();

  // pass B: turn counts into output bases, warps in order within the CTA
  {
    const uint32_t *hrow = hist + (size_t)blockIdx.x * R;
    const uint32_t *grow = gtot + (size_t)(blockIdx.x / gs) * R;
    for (uint32_t d = threadIdx.x; d < R; d += kSortThreads) {
      uint32_t c[kSortWarps];
#pragma unroll
      for (int ww = 0; ww < kSortWarps; ww++) c[ww] = wcount_all[(size_t)ww * R + d];
      uint32_t run = hrow[d] + grow[d] + dbase[d];
#pragma unroll
      for (int ww = 0; ww < kSortWarps; ww++) {
        wcount_all[(size_t)ww * R + d] = run;
        run += c[ww];
      }
    }
  }
  __syncthreads();

  // pass C: the leader takes the group's base with an atomic and broadcasts it; the rank
  // within the group gives the position
  uint32_t *rout = rank + wbase;
#pragma unroll
  for (int it = 0; it < kItemsPerLane; it++) {
    const bool ok = dig[it] != 0xffffffffu;
    uint32_t grp = __match_any_sync(0xffffffffu, dig[it]);
    uint32_t base = 0;
    if (ok && (grp & lt) == 0) base = atomicAdd(&wcount[dig[it]], (uint32_t)__popc(grp));
    __syncwarp();                   // orders this iteration's atomics before the next one's
    base = __shfl_sync(0xffffffffu, base, __ffs(grp) - 1);
    if (ok) {
      const uint32_t pos = base + __popc(grp & lt);
      if (DIRECT) {
        keys_out[pos] = kin[it * 32];
        perm_out[pos] = perm_in ? perm_in[wbase + it * 32] : wbase + it * 32;
      } else {
        rout[it * 32] = pos;
      }
    }
  }
}

// Data movement of one pass: every item goes to its sorted position.
template <bool PAYLOAD>
__global__ void __launch_bounds__(256)
sort_apply_kernel(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ perm_in,
                  const uint32_t *__restrict__ rank, uint32_t *__restrict__ keys_out,
                  uint32_t *__restrict__ perm_out, uint32_t n, const rec4 *__restrict__ psrc,
                  rec4 *__restrict__ pdst)
{
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t pos = rank[i];
    const uint32_t slot = perm_in ? perm_in[i] : i;
    keys_out[pos] = keys_in[i];
    perm_out[pos] = slot;
    if (PAYLOAD) st_rec(&pdst[pos], ld_rec(&psrc[slot]));
  }
}

// seg[k] = number of sorted keys < k  (first position of key k), k in [0, nkeys]
__global__ void seg_offsets_kernel(const uint32_t *__restrict__ sorted, size_t n,
                                   uint32_t *__restrict__ seg, uint32_t nkeys, uint32_t base)
{
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > nkeys) return;
  size_t lo = 0, hi = n;
  while (lo < hi) {
    size_t mid = (lo + hi) >> 1;
    if (sorted[mid] < k) lo = mid + 1; else hi = mid;
  }
  seg[k] = base + (uint32_t)lo;
}

// ---- compaction of the indices where two arrays differ (the level switches of a substep) ----
constexpr int kCompactChunk = 4096;      // items per CTA: 256 threads x 16 consecutive items

__global__ void __launch_bounds__(256)
compact_count_kernel(const unsigned *__restrict__ a, const unsigned *__restrict__ b, size_t n,
                     uint32_t *__restrict__ counts)
{
  const size_t base = (size_t)blockIdx.x * kCompactChunk;
  unsigned cnt = 0;
  for (int k = threadIdx.x; k < kCompactChunk; k += 256) {
    const size_t i = base + k;
    if (i < n && a[i] != b[i]) cnt++;
  }
  __shared__ unsigned ws[8];
  for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = 0;
    for (int w = 0; w < 8; w++) t += ws[w];
    counts[blockIdx.x] = t;
  }
}

// exclusive scan of counts[0..nb) into bases[0..nb], bases[nb] = total (one CTA)
__global__ void __launch_bounds__(1024)
compact_scan_kernel(const uint32_t *__restrict__ counts, uint32_t *__restrict__ bases, unsigned nb)
{
  __shared__ uint32_t ws[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (unsigned b0 = 0; b0 < nb; b0 += 1024) {
    const unsigned i = b0 + threadIdx.x;
    const uint32_t v = i < nb ? counts[i] : 0u;
    uint32_t x = v;
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t u = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += u;
    }
    if (lane == 31) ws[w] = x;
    __syncthreads();
    if (w == 0) {
      uint32_t y = ws[lane];
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t u = __shfl_up_sync(0xffffffffu, y, o);
        if (lane >= o) y += u;
      }
      ws[lane] = y;
    }
    __syncthreads();
    const uint32_t incl = x + (w ? ws[w - 1] : 0u) + carry;
    if (i < nb) bases[i] = incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) bases[nb] = carry;
}

__global__ void __launch_bounds__(256)
compact_write_kernel(const unsigned *__restrict__ a, const unsigned *__restrict__ b, size_t n,
                     const uint32_t *__restrict__ bases, uint32_t *__restrict__ list)
{
  // thread t owns the 16 consecutive items [base + 16 t, base + 16 t + 16): ranks in index order
  const size_t base = (size_t)blockIdx.x * kCompactChunk + (size_t)threadIdx.x * 16;
  unsigned flags = 0;
#pragma unroll
  for (int k = 0; k < 16; k++) {
    const size_t i = base + k;
    if (i < n && a[i] != b[i]) flags |= 1u << k;
  }
  const unsigned cnt = __popc(flags);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned x = cnt;
  for (int o = 1; o < 32; o <<= 1) {
    unsigned u = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += u;
  }
  __shared__ unsigned ws[8];
  if (lane == 31) ws[w] = x;
  __syncthreads();
  unsigned off = bases[blockIdx.x] + x - cnt;
  for (int ww = 0; ww < w; ww++) off += ws[ww];
#pragma unroll
  for (int k = 0; k < 16; k++)
    if (flags & (1u << k)) list[off++] = (uint32_t)(base + k);
}

struct Plan { int npass, bits; unsigned nctas, gs; };

Plan plan_for(size_t n, uint32_t nkeys)
{
  int nbits = 1;
  while ((1ull << nbits) < nkeys) nbits++;
  Plan p;
  p.npass = (nbits + kMaxDigitBits - 1) / kMaxDigitBits;
  p.bits = (nbits + p.npass - 1) / p.npass;
  p.nctas = div_up(n ? n : 1, kSortTile);
  p.gs = (p.nctas + kScanGroups - 1) / kScanGroups;
  return p;
}

} // namespace

size_t sort_hist_words(size_t n)
{
  // hist[cta][digit] + gtot[group][digit] + dtot[digit] + dbase[digit]
  return (size_t)kMaxRadix * ((size_t)div_up(n ? n : 1, kSortTile) + kScanGroups + 2);
}

int radix_sort_launches(size_t n, uint32_t nkeys, bool want_seg, bool payload)
{
  // per pass: histogram, three scan kernels, rank (which also moves keys and slots); the last pass of
  // a sort with a payload ranks and applies separately
  Plan p = plan_for(n, nkeys);
  return 5 * p.npass + (payload ? 1 : 0) + ((want_seg && p.npass > 1) ? 1 : 0);
}

int radix_sort_pairs(cudaStream_t st, uint32_t *keys[2], uint32_t *perm[2], uint32_t *hist,
                     uint32_t *rank, size_t n, uint32_t nkeys, int *out_cur, uint32_t *seg,
                     SortPayload payload, const uint32_t *perm_init)
{
  // keys[0] holds the unsorted keys; values start as the identity.
  const Plan p = plan_for(n, nkeys);
  const uint32_t R = 1u << p.bits, mask = R - 1u;
  uint32_t *gtot = hist + (size_t)p.nctas * R;
  uint32_t *dtot = gtot + (size_t)kScanGroups * R;
  uint32_t *dbase = dtot + R;
  const size_t smem = (size_t)kSortWarps * R * sizeof(uint32_t);
  static bool attr_done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev & 63]) {
    CU_TRY(cudaFuncSetAttribute(sort_rank_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(kSortWarps * kMaxRadix * sizeof(uint32_t))));
    CU_TRY(cudaFuncSetAttribute(sort_rank_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(kSortWarps * kMaxRadix * sizeof(uint32_t))));
    attr_done[dev & 63] = true;
  }
  int cur = 0;
  for (int q = 0; q < p.npass; q++) {
    const int shift = q * p.bits;
    const bool last = q == p.npass - 1;
    sort_hist_kernel<<<p.nctas, kSortThreads, 0, st>>>(keys[cur], n, shift, mask, hist);
    sort_scan1_kernel<<<dim3(div_up(R, 256), kScanGroups), 256, 0, st>>>(hist, gtot, R, p.nctas, p.gs);
    sort_scan2a_kernel<<<div_up(R, 256), 256, 0, st>>>(gtot, dtot, R);
    sort_scan2b_kernel<<<1, 1024, 0, st>>>(dtot, dbase, R, (seg && p.npass == 1) ? seg : nullptr, nkeys,
                                           (uint32_t)n);
    const uint32_t *pin = q == 0 ? perm_init : perm[cur];
    if (last && payload.src) {
      sort_rank_kernel<false><<<p.nctas, kSortThreads, smem, st>>>(keys[cur], rank, n, shift, mask, hist, gtot,
                                                                   dbase, p.gs, nullptr, nullptr, nullptr);
      sort_apply_kernel<true><<<grid_for(n, 256), 256, 0, st>>>(keys[cur], pin, rank, keys[cur ^ 1],
                                                                perm[cur ^ 1], (uint32_t)n, payload.src,
                                                                payload.dst);
    } else {
      sort_rank_kernel<true><<<p.nctas, kSortThreads, smem, st>>>(keys[cur], rank, n, shift, mask, hist, gtot,
                                                                  dbase, p.gs, pin, keys[cur ^ 1], perm[cur ^ 1]);
    }
    cur ^= 1;
  }
  CU_LAUNCH_CHECK();
  *out_cur = cur;
  if (seg && p.npass > 1) return segment_offsets(st, keys[cur], n, seg, nkeys);
  return EXPB200_OK;
}

int compact_differing(cudaStream_t st, const unsigned *a, const unsigned *b, size_t n, uint32_t *scratch,
                      uint32_t *list, size_t *count_host)
{
  *count_host = 0;
  if (n == 0) return EXPB200_OK;
  const unsigned nb = div_up(n, kCompactChunk);
  uint32_t *counts = scratch, *bases = scratch + nb;
  compact_count_kernel<<<nb, 256, 0, st>>>(a, b, n, counts);
  compact_scan_kernel<<<1, 1024, 0, st>>>(counts, bases, nb);
  CU_LAUNCH_CHECK();
  uint32_t total = 0;
  CU_TRY(cudaMemcpyAsync(&total, bases + nb, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  *count_host = total;
  if (total == 0) return EXPB200_OK;
  compact_write_kernel<<<nb, 256, 0, st>>>(a, b, n, bases, list);
  CU_LAUNCH_CHECK();
  return EXPB200_OK;
}

int segment_offsets(cudaStream_t st, const uint32_t *sorted, size_t n, uint32_t *seg, uint32_t nkeys,
                    uint32_t base)
{
  seg_offsets_kernel<<<div_up((size_t)nkeys + 1, 256), 256, 0, st>>>(sorted, n, seg, nkeys, base);
  CU_LAUNCH_CHECK();
  return EXPB200_OK;
}

__global__ void add_u32_kernel(uint32_t *__restrict__ a, size_t n, uint32_t add)
{
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    a[i] += add;
}

int radix_sort_passes(uint32_t nkeys) { return plan_for(1, nkeys).npass; }

// the result of p passes started in buffer W lands in W ^ (p & 1): start where it ends in cur
int suffix_key_buffer(int cur, uint32_t nkeys_sub) { return cur ^ (radix_sort_passes(nkeys_sub) & 1); }

int radix_sort_suffix(cudaStream_t st, uint32_t *keys[2], uint32_t *perm[2], int cur, uint32_t *hist,
                      uint32_t *rank, size_t off, size_t n, uint32_t nkeys_sub, uint32_t *seg_sub,
                      SortPayload payload, uint32_t key_base)
{
  const size_t nsub = n - off;
  const int W = suffix_key_buffer(cur, nkeys_sub);
  uint32_t *kl[2] = {keys[W] + off, keys[W ^ 1] + off};
  uint32_t *pl[2] = {perm[W] + off, perm[W ^ 1] + off};
  if (W != cur)        // odd number of passes: the first pass writes perm[cur], so it reads a copy
    CU_TRY(cudaMemcpyAsync(pl[0], perm[cur] + off, nsub * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
  if (payload.dst) payload.dst += off;
  int out = 0;
  int rc = radix_sort_pairs(st, kl, pl, hist, rank, nsub, nkeys_sub, &out, nullptr, payload, pl[0]);
  if (rc) return rc;
  if ((W ^ out) != cur) { set_error("radix_sort_suffix: result in the wrong buffer"); return EXPB200_ESTATE; }
  rc = segment_offsets(st, keys[cur] + off, nsub, seg_sub, nkeys_sub, (uint32_t)off);
  if (rc) return rc;
  if (key_base) {        // the sorted keys as the consumers index them (absolute)
    add_u32_kernel<<<grid_for(nsub, 256), 256, 0, st>>>(keys[cur] + off, nsub, key_base);
    CU_LAUNCH_CHECK();
  }
  return EXPB200_OK;
}

} // namespace expb200
