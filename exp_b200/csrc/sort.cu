// sort.cu -- stable LSD radix sort of (key -> particle slot) pairs and segment
// offsets.  Replaces the reference's thrust::sort_by_key / lower_bound level
// sort (src/cudaComponent.cu:383-420, 490-526) and additionally bins particles
// by radial cell, which is what makes the deterministic segmented coefficient
// reduction and the shared-row table gathers possible.
//
// Stability matters: the summation order of the coefficient reduction is the
// sorted order, so a stable sort of keys computed from the particle state makes
// every result a deterministic function of that state.
#include "sort.cuh"

namespace expb200 {

namespace {

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kItemsPerWarp = 512;                       // 16 rounds of 32
constexpr int kSortTile = kSortWarps * kItemsPerWarp;    // 4096 keys per CTA
constexpr int kMaxRadix = 256;

__device__ __forceinline__ uint32_t digit_of(uint32_t key, int shift, uint32_t mask)
{
  return (key >> shift) & mask;
}

// Per-CTA digit histogram, written digit-major so that one exclusive scan over
// hist[digit][cta] yields every CTA's output base for every digit.
__global__ void __launch_bounds__(kSortThreads)
sort_hist_kernel(const uint32_t *__restrict__ keys, size_t n, int shift, uint32_t mask,
                 uint32_t *__restrict__ hist, unsigned nctas)
{
  __shared__ uint32_t h[kMaxRadix];
  for (int d = threadIdx.x; d < kMaxRadix; d += kSortThreads) h[d] = 0;
  __syncthreads();
  size_t base = (size_t)blockIdx.x * kSortTile;
#pragma unroll 4
  for (int it = 0; it < kSortTile / kSortThreads; it++) {
    size_t i = base + (size_t)it * kSortThreads + threadIdx.x;
    if (i < n) atomicAdd(&h[digit_of(keys[i], shift, mask)], 1u);   // integer: order-independent
  }
  __syncthreads();
  for (uint32_t d = threadIdx.x; d <= mask; d += kSortThreads)
    hist[(size_t)d * nctas + blockIdx.x] = h[d];
}

// Exclusive scan of `total` counters by one CTA of 1024 threads: each thread owns
// a contiguous slice (serial sum, block scan of the slice sums, serial rewrite).
__global__ void __launch_bounds__(1024)
sort_scan_kernel(uint32_t *__restrict__ a, size_t total)
{
  __shared__ uint32_t warp_sums[32];
  const int t = threadIdx.x;
  size_t per = (total + 1023) / 1024;
  size_t lo = (size_t)t * per, hi = lo + per < total ? lo + per : total;
  uint32_t s = 0;
  for (size_t i = lo; i < hi; i++) s += a[i];
  // block exclusive scan of s
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
  uint32_t run = v - s + (w ? warp_sums[w - 1] : 0u);
  for (size_t i = lo; i < hi; i++) {
    uint32_t c = a[i];
    a[i] = run;
    run += c;
  }
}

// Stable scatter.  Each warp owns 512 consecutive keys of the CTA tile and walks
// them in order, 32 at a time; __match_any_sync groups equal digits, the lowest
// lane of a group is its leader.
__global__ void __launch_bounds__(kSortThreads)
sort_scatter_kernel(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ perm_in,
                    uint32_t *__restrict__ keys_out, uint32_t *__restrict__ perm_out,
                    size_t n, int shift, uint32_t mask,
                    const uint32_t *__restrict__ offs, unsigned nctas)
{
  __shared__ uint32_t wcount[kSortWarps][kMaxRadix];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t lt = (1u << lane) - 1u;

  for (int d = lane; d < kMaxRadix; d += 32) wcount[w][d] = 0;
  __syncwarp();

  const size_t wbase = (size_t)blockIdx.x * kSortTile + (size_t)w * kItemsPerWarp;

  // pass A: per-warp digit counts
  for (int it = 0; it < kItemsPerWarp / 32; it++) {
    size_t i = wbase + it * 32 + lane;
    uint32_t d = i < n ? digit_of(keys_in[i], shift, mask) : 0xffffffffu;
    uint32_t grp = __match_any_sync(0xffffffffu, d);
    if (i < n && (grp & lt) == 0) wcount[w][d] += __popc(grp);
    __syncwarp();
  }
  __syncthreads();

  // pass B: turn counts into output bases, warps in order within the CTA
  for (uint32_t d = threadIdx.x; d <= mask; d += kSortThreads) {
    uint32_t run = offs[(size_t)d * nctas + blockIdx.x];
#pragma unroll
    for (int ww = 0; ww < kSortWarps; ww++) {
      uint32_t c = wcount[ww][d];
      wcount[ww][d] = run;
      run += c;
    }
  }
  __syncthreads();

  // pass C: ranks within the group give the position; the leader advances the base
  for (int it = 0; it < kItemsPerWarp / 32; it++) {
    size_t i = wbase + it * 32 + lane;
    bool ok = i < n;
    uint32_t key = ok ? keys_in[i] : 0u;
    uint32_t d = ok ? digit_of(key, shift, mask) : 0xffffffffu;
    uint32_t grp = __match_any_sync(0xffffffffu, d);
    uint32_t pos = 0;
    if (ok) pos = wcount[w][d] + __popc(grp & lt);
    __syncwarp();
    if (ok && (grp & lt) == 0) wcount[w][d] += __popc(grp);
    __syncwarp();
    if (ok) {
      keys_out[pos] = key;
      perm_out[pos] = perm_in ? perm_in[i] : (uint32_t)i;
    }
  }
}

// seg[k] = number of sorted keys < k  (first position of key k), k in [0, nkeys]
__global__ void seg_offsets_kernel(const uint32_t *__restrict__ sorted, size_t n,
                                   uint32_t *__restrict__ seg, uint32_t nkeys)
{
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > nkeys) return;
  size_t lo = 0, hi = n;
  while (lo < hi) {
    size_t mid = (lo + hi) >> 1;
    if (sorted[mid] < k) lo = mid + 1; else hi = mid;
  }
  seg[k] = (uint32_t)lo;
}

} // namespace

size_t sort_hist_words(size_t n)
{
  return (size_t)kMaxRadix * div_up(n ? n : 1, kSortTile);
}

int radix_sort_pairs(cudaStream_t st, uint32_t *keys[2], uint32_t *perm[2], uint32_t *hist,
                     size_t n, uint32_t nkeys, int *out_cur)
{
  // keys[0] holds the unsorted keys; values start as the identity.
  int nbits = 1;
  while ((1ull << nbits) < nkeys) nbits++;
  int npass = (nbits + 7) / 8;
  int bits = (nbits + npass - 1) / npass;
  unsigned nctas = div_up(n ? n : 1, kSortTile);
  int cur = 0;
  for (int p = 0; p < npass; p++) {
    int shift = p * bits;
    uint32_t mask = (1u << bits) - 1u;
    sort_hist_kernel<<<nctas, kSortThreads, 0, st>>>(keys[cur], n, shift, mask, hist, nctas);
    sort_scan_kernel<<<1, 1024, 0, st>>>(hist, (size_t)(mask + 1) * nctas);
    sort_scatter_kernel<<<nctas, kSortThreads, 0, st>>>(keys[cur], p == 0 ? nullptr : perm[cur],
                                                        keys[cur ^ 1], perm[cur ^ 1], n, shift,
                                                        mask, hist, nctas);
    cur ^= 1;
  }
  CU_LAUNCH_CHECK();
  *out_cur = cur;
  return EXPB200_OK;
}

int segment_offsets(cudaStream_t st, const uint32_t *sorted, size_t n, uint32_t *seg, uint32_t nkeys)
{
  seg_offsets_kernel<<<div_up((size_t)nkeys + 1, 256), 256, 0, st>>>(sorted, n, seg, nkeys);
  CU_LAUNCH_CHECK();
  return EXPB200_OK;
}

} // namespace expb200
