// sort.cuh -- stable key/slot radix binning + segment offsets (see sort.cu)
#pragma once
#include "common.cuh"

namespace expb200 {

// Number of uint32 words of histogram scratch needed for n keys.
size_t sort_hist_words(size_t n);

// Optional payload of the last pass: one 32-byte record per particle is moved into
// sorted order (dst[k] = src[slot of sorted position k]) together with the permutation,
// so that the kernels that walk the sorted order stream it instead of gathering.
struct SortPayload {
  const rec4 *src = nullptr;
  rec4 *dst = nullptr;
};

// Sorts (keys[0][i], i) for i<n by key (keys < nkeys), stably.  On return
// keys[*out_cur] are the sorted keys and perm[*out_cur] the slots in sorted order.
// seg (nkeys+1 words, may be NULL) receives seg[k] = first sorted position whose key
// is >= k.  Digits are up to 11 bits wide: nkeys <= 2048 is a single pass.
// rank: n words of scratch (the sorted position of every item).
int radix_sort_pairs(cudaStream_t st, uint32_t *keys[2], uint32_t *perm[2], uint32_t *hist,
                     uint32_t *rank, size_t n, uint32_t nkeys, int *out_cur, uint32_t *seg = nullptr,
                     SortPayload payload = SortPayload());

// Number of kernels radix_sort_pairs launches for these sizes (for the launch counters).
int radix_sort_launches(size_t n, uint32_t nkeys, bool want_seg);

// seg[k] = first sorted position whose key is >= k, for k in [0, nkeys].
int segment_offsets(cudaStream_t st, const uint32_t *sorted, size_t n, uint32_t *seg, uint32_t nkeys);

} // namespace expb200
