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
// perm_init (may be NULL = identity): the slot of input item i; must not alias perm[1].
int radix_sort_pairs(cudaStream_t st, uint32_t *keys[2], uint32_t *perm[2], uint32_t *hist,
                     uint32_t *rank, size_t n, uint32_t nkeys, int *out_cur, uint32_t *seg = nullptr,
                     SortPayload payload = SortPayload(), const uint32_t *perm_init = nullptr);

// Indices i < n with a[i] != b[i], ascending, into list (capacity n words; may not alias a or b);
// *count_host receives their number (the call synchronises the stream).  scratch: at least
// 2 * ceil(n / 4096) + 2 words.  Deterministic: a counting pass, a scan, a writing pass.
int compact_differing(cudaStream_t st, const unsigned *a, const unsigned *b, size_t n, uint32_t *scratch,
                      uint32_t *list, size_t *count_host);

// Number of kernels radix_sort_pairs launches for these sizes (for the launch counters).
int radix_sort_launches(size_t n, uint32_t nkeys, bool want_seg, bool payload = false);

// seg[k] = base + first sorted position whose key is >= k, for k in [0, nkeys].
int segment_offsets(cudaStream_t st, const uint32_t *sorted, size_t n, uint32_t *seg, uint32_t nkeys,
                    uint32_t base = 0);

// Number of passes radix_sort_pairs makes for keys < nkeys.
int radix_sort_passes(uint32_t nkeys);

// Re-bins the tail [off, n) of a sorted permutation in place: on entry perm[cur][off..n) holds the
// slots to sort and keys[W][off + k] the new key (< nkeys_sub) of slot perm[cur][off + k], where
// W = suffix_key_buffer(cur, nkeys_sub); on return keys[cur][off..n), perm[cur][off..n) are sorted
// (cur unchanged, [0, off) untouched) and seg_sub[k] = off + first position of the tail whose key
// is >= k, k in [0, nkeys_sub]; key_base is then added to the sorted keys.  payload.dst is indexed
// like perm (dst[off + k]).
int suffix_key_buffer(int cur, uint32_t nkeys_sub);
int radix_sort_suffix(cudaStream_t st, uint32_t *keys[2], uint32_t *perm[2], int cur, uint32_t *hist,
                      uint32_t *rank, size_t off, size_t n, uint32_t nkeys_sub, uint32_t *seg_sub,
                      SortPayload payload = SortPayload(), uint32_t key_base = 0);

} // namespace expb200
