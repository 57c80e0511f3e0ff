// nccl_shim.h -- NCCL reached through dlopen so that libexpb200.so has no link-time
// dependency on it (the host process decides which libnccl.so.2 is loaded: the
// reference's MPI ranks would load the system one, a torch process its bundled one).
#pragma once
#include <cuda_runtime.h>
#include <cstddef>

namespace expb200 {
// In-place sum all-reduce of n doubles on `stream` (the reference's MPI_Allreduce of
// the coefficient vector, src/SphericalBasis.cc:864-903, but device-resident and one
// call for the whole vector).
int nccl_allreduce_sum_f64(void *comm, double *buf, size_t n, cudaStream_t stream);
}
