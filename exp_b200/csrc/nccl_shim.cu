// nccl_shim.cu -- see nccl_shim.h
#include "nccl_shim.h"
#include "common.cuh"

#include <dlfcn.h>

namespace expb200 {
namespace {

// Minimal NCCL ABI (nccl.h): ncclResult_t is an int enum, ncclSuccess == 0,
// ncclFloat64 == 8, ncclSum == 0, ncclUniqueId is 128 opaque bytes.
typedef int (*allreduce_fn)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef int (*getuid_fn)(void *);
typedef int (*destroy_fn)(void *);
typedef const char *(*errstr_fn)(int);

struct Api {
  void *h = nullptr;
  allreduce_fn allreduce = nullptr;
  getuid_fn getuid = nullptr;
  void *initrank = nullptr;
  destroy_fn destroy = nullptr;
  errstr_fn errstr = nullptr;
};

Api *api()
{
  static Api a;
  if (a.h) return &a;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    a.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (a.h) break;
  }
  if (!a.h) {
    set_error("cannot dlopen libnccl.so.2: %s", dlerror());
    return nullptr;
  }
  a.allreduce = (allreduce_fn)dlsym(a.h, "ncclAllReduce");
  a.getuid = (getuid_fn)dlsym(a.h, "ncclGetUniqueId");
  a.initrank = dlsym(a.h, "ncclCommInitRank");
  a.destroy = (destroy_fn)dlsym(a.h, "ncclCommDestroy");
  a.errstr = (errstr_fn)dlsym(a.h, "ncclGetErrorString");
  if (!a.allreduce || !a.getuid || !a.initrank || !a.destroy) {
    set_error("libnccl is missing expected symbols");
    a.h = nullptr;
    return nullptr;
  }
  return &a;
}

struct uid128 { char bytes[128]; };

int fail(Api *a, int rc, const char *what)
{
  set_error("NCCL error %d (%s) in %s", rc, a && a->errstr ? a->errstr(rc) : "?", what);
  return EXPB200_ENCCL;
}

} // namespace

int nccl_allreduce_sum_f64(void *comm, double *buf, size_t n, cudaStream_t stream)
{
  Api *a = api();
  if (!a) return EXPB200_ENCCL;
  int rc = a->allreduce(buf, buf, n, /*ncclFloat64*/ 8, /*ncclSum*/ 0, comm, stream);
  if (rc) return fail(a, rc, "ncclAllReduce");
  return EXPB200_OK;
}

} // namespace expb200

using namespace expb200;

extern "C" {

int expb200_comm_unique_id(char id[128])
{
  REQUIRE(id, "id is NULL");
  Api *a = api();
  if (!a) return EXPB200_ENCCL;
  int rc = a->getuid(id);
  if (rc) return fail(a, rc, "ncclGetUniqueId");
  return EXPB200_OK;
}

int expb200_comm_create(int device, int nranks, int rank, const char id[128], void **comm)
{
  REQUIRE(id && comm, "NULL argument");
  Api *a = api();
  if (!a) return EXPB200_ENCCL;
  DeviceGuard g(device);
  uid128 u;
  memcpy(u.bytes, id, 128);
  typedef int (*fn)(void **, int, uid128, int);
  int rc = ((fn)a->initrank)(comm, nranks, u, rank);
  if (rc) return fail(a, rc, "ncclCommInitRank");
  return EXPB200_OK;
}

int expb200_comm_destroy(void *comm)
{
  Api *a = api();
  if (!a) return EXPB200_ENCCL;
  if (comm) a->destroy(comm);
  return EXPB200_OK;
}

int expb200_allreduce_sum(void *comm, double *device_buf, size_t n, void *stream)
{
  return nccl_allreduce_sum_f64(comm, device_buf, n, (cudaStream_t)stream);
}

} // extern "C"
