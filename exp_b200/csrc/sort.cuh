// sort.cuh -- stable key/slot radix sort + segment offsets (see sort.cu)
#pragma once
#include "common.cuh"

namespace expb200 {

// Number of uint32 words of histogram scratch needed for n keys.
size_t sort_hist_words(size_t n);

// Sorts (keys[0][i], i) for i<n by key (keys < nkeys), stably.  On return
// keys[*out_cur] are the sorted keys and perm[*out_cur] the slots in sorted order.
int radix_sort_pairs(cudaStream_t st, uint32_t *keys[2], uint32_t *perm[2], uint32_t *hist,
                     size_t n, uint32_t nkeys, int *out_cur);

// seg[k] = first sorted position whose key is >= k, for k in [0, nkeys].
int segment_offsets(cudaStream_t st, const uint32_t *sorted, size_t n, uint32_t *seg, uint32_t nkeys);

} // namespace expb200
