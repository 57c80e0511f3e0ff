// profile.cuh -- per-kernel launch counters and optional CUDA-event timing of the
// kernels on the stream they are launched on (what bench.py's roofline reads).
#pragma once
#include "common.cuh"

namespace expb200 {

enum KernelId {
  K_KEY = 0, K_SORT, K_SEG, K_DEPOSIT, K_COMBINE, K_CONTRACT, K_FIELD, K_FORCE,
  K_KICK, K_DRIFT, K_ZERO, K_LEVELS, K_ALLREDUCE, K_PACK, K_COV, K_CYL_DEPOSIT,
  K_CYL_FORCE, K_CYL_FIELD, K_KDK, K_REORDER, K_OTHER, K_COUNT
};

// Counts `launches` kernel launches for id; when timing is enabled, brackets the
// region with events on `st` (RAII).
struct ProfRegion {
  int id; cudaStream_t st; int slot;
  ProfRegion(KernelId id, cudaStream_t st, int launches = 1);
  ~ProfRegion();
};

} // namespace expb200
