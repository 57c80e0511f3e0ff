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

// ---------------------------------------------------------------------------
// The communicator handed out by expb200_comm_create: the NCCL communicator plus,
// once expb200_comm_peer_connect has run, peer-mapped receive buffers for the small
// all-reduces of this path (882 doubles for sphereSL at L=6, nmax=18; 234 for the
// cylinder; 10 per level for the centring pass).  For those NCCL's launch + protocol
// latency (~50 us) is the whole cost, so they are done by ONE kernel over NVLink peer
// memory: every rank stores its vector straight into slot [rank] of every peer's buffer
// (P2P stores through NVSwitch), publishes a sequence number, waits for the others'
// numbers and adds the world's vectors in rank order -- the result is therefore bitwise
// identical on every rank and independent of arrival order, which ncclAllReduce does not
// promise.  Buffers are double-buffered by the parity of the sequence number: a rank can
// run at most one all-reduce ahead of the slowest (it needs that rank's data to finish).
// That argument needs the calls of one communicator to execute in issue order on every
// rank; callers pass different streams (one per component), so each call waits on the
// event recorded after the previous one when the stream differs.  A flag that is not
// exactly the expected sequence number, or a wait longer than the time-out
// (EXPB200_PEER_TIMEOUT_S seconds, default 600, 0 = wait for ever like MPI), poisons the
// vector with NaN instead of summing stale slots and raises the error flag, which
// expb200_comm_check() and the next all-reduce report.
// ---------------------------------------------------------------------------
constexpr int kPeerMaxWorld = 16;
constexpr size_t kPeerCap = 16384;            // doubles per rank slot
constexpr uint32_t kCommMagic = 0xb200c033u;

struct PeerPtrs {
  double *data[kPeerMaxWorld];                // rank r's receive area, [2][world][kPeerCap]
  unsigned long long *flags[kPeerMaxWorld];   // rank r's flags, [2][world]
};

struct Comm {
  uint32_t magic = kCommMagic;
  void *nccl = nullptr;
  int device = 0, world = 1, rank = 0;
  void *local = nullptr;                      // this rank's allocation (flags, then data)
  void *mapped[kPeerMaxWorld] = {};           // peers' allocations opened through CUDA IPC
  PeerPtrs P = {};
  bool connected = false;
  unsigned long long seq = 0;
  int *err_host = nullptr, *err_dev = nullptr;   // mapped pinned: set by a kernel that timed out
  cudaStream_t last_stream = nullptr;            // stream of the previous all-reduce
  cudaEvent_t last_done = nullptr;               // recorded after it
  unsigned long long timeout_ns = 600ull * 1000000000ull;
  cudaStream_t host_stream = nullptr;            // expb200_allreduce_sum_host staging
  double *host_stage = nullptr;
  size_t host_stage_cap = 0;
};

__device__ __forceinline__ unsigned long long global_ns()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

size_t peer_flag_bytes(int world) { return ((size_t)2 * world * sizeof(unsigned long long) + 255) / 256 * 256; }
size_t peer_bytes(int world) { return peer_flag_bytes(world) + (size_t)2 * world * kPeerCap * sizeof(double); }

__global__ void __launch_bounds__(256)
peer_allreduce_kernel(PeerPtrs P, int world, int rank, double *__restrict__ buf, size_t n,
                      unsigned long long seq, unsigned long long timeout_ns, int *err)
{
  __shared__ int bad;
  const int par = (int)(seq & 1ull), tid = threadIdx.x;
  if (tid == 0) bad = 0;
  // 1. my vector -> slot [par][rank] of every rank (P2P stores; the local one included)
  for (int r = 0; r < world; r++) {
    double *dst = P.data[r] + ((size_t)par * world + rank) * kPeerCap;
    for (size_t i = tid; i < n; i += blockDim.x) dst[i] = buf[i];
  }
  __threadfence_system();
  __syncthreads();
  if (tid < world) {
    volatile unsigned long long *f = P.flags[tid] + (size_t)par * world + rank;
    *f = seq;
  }
  // 2. wait until every rank's vector of this sequence number has landed here
  if (tid < world) {
    volatile unsigned long long *f = P.flags[rank] + (size_t)par * world + tid;
    const unsigned long long t0 = global_ns();
    unsigned long long v;
    while ((v = *f) < seq) {
      if (timeout_ns && global_ns() - t0 > timeout_ns) { *err = 1; bad = 1; break; }   // a peer died
    }
    if (v > seq) { *err = 2; bad = 1; }      // calls of this communicator ran out of order
  }
  __syncthreads();
  __threadfence_system();
  if (bad) {                                  // never sum stale slots: make the failure visible
    for (size_t i = tid; i < n; i += blockDim.x) buf[i] = __longlong_as_double(0x7ff8000000000000ll);
    return;
  }
  // 3. sum in rank order (bypassing L1: the data was written by other GPUs)
  const double *src = P.data[rank] + (size_t)par * world * kPeerCap;
  for (size_t i = tid; i < n; i += blockDim.x) {
    double t = 0.0;
    for (int r = 0; r < world; r++) t += __ldcg(src + (size_t)r * kPeerCap + i);
    buf[i] = t;
  }
}

Comm *as_comm(void *comm)
{
  Comm *c = static_cast<Comm *>(comm);
  return c && c->magic == kCommMagic ? c : nullptr;
}

} // namespace

int nccl_allreduce_sum_f64(void *comm, double *buf, size_t n, cudaStream_t stream)
{
  Comm *c = as_comm(comm);
  if (!c) { set_error("not an expb200 communicator"); return EXPB200_EINVAL; }
  DeviceGuard g(c->device);         // the stream belongs to the communicator's device
  // collectives on one communicator must execute in issue order on every rank (the peer
  // kernel's double buffering and NCCL both rely on it): join the previous call's stream
  if (c->last_done && c->last_stream != stream) CU_TRY(cudaStreamWaitEvent(stream, c->last_done, 0));
  if (c->connected && n <= kPeerCap) {
    if (*c->err_host) {
      set_error(*c->err_host == 2 ? "peer all-reduce: sequence numbers out of order"
                                  : "peer all-reduce timed out waiting for another rank");
      return EXPB200_ENCCL;
    }
    c->seq++;
    peer_allreduce_kernel<<<1, 256, 0, stream>>>(c->P, c->world, c->rank, buf, n, c->seq, c->timeout_ns,
                                                 c->err_dev);
    CU_LAUNCH_CHECK();
  } else {
    Api *a = api();
    if (!a) return EXPB200_ENCCL;
    int rc = a->allreduce(buf, buf, n, /*ncclFloat64*/ 8, /*ncclSum*/ 0, c->nccl, stream);
    if (rc) return fail(a, rc, "ncclAllReduce");
  }
  if (!c->last_done) CU_TRY(cudaEventCreateWithFlags(&c->last_done, cudaEventDisableTiming));
  CU_TRY(cudaEventRecord(c->last_done, stream));
  c->last_stream = stream;
  return EXPB200_OK;
}

} // namespace expb200

using namespace expb200;

extern "C" {

int expb200_comm_check(void *comm);

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
  REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "rank %d of %d", rank, nranks);
  Api *a = api();
  if (!a) return EXPB200_ENCCL;
  DeviceGuard g(device);
  uid128 u;
  memcpy(u.bytes, id, 128);
  typedef int (*fn)(void **, int, uid128, int);
  void *nc = nullptr;
  int rc = ((fn)a->initrank)(&nc, nranks, u, rank);
  if (rc) return fail(a, rc, "ncclCommInitRank");
  Comm *c = new Comm;
  c->nccl = nc; c->device = device; c->world = nranks; c->rank = rank;
  *comm = c;
  return EXPB200_OK;
}

int expb200_comm_destroy(void *comm)
{
  Comm *c = as_comm(comm);
  if (!c) return EXPB200_OK;
  DeviceGuard g(c->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < c->world; r++)
    if (c->mapped[r]) cudaIpcCloseMemHandle(c->mapped[r]);
  if (c->local) cudaFree(c->local);
  if (c->err_host) cudaFreeHost(c->err_host);
  if (c->last_done) cudaEventDestroy(c->last_done);
  if (c->host_stage) cudaFree(c->host_stage);
  if (c->host_stream) cudaStreamDestroy(c->host_stream);
  Api *a = api();
  if (a && c->nccl) a->destroy(c->nccl);
  c->magic = 0;
  delete c;
  return EXPB200_OK;
}

int expb200_allreduce_sum(void *comm, double *device_buf, size_t n, void *stream)
{
  return nccl_allreduce_sum_f64(comm, device_buf, n, (cudaStream_t)stream);
}

// In-place sum over ranks of a HOST vector (the reference's MPI_Allreduce of the covariance
// arrays, src/AxisymmetricBasis.cc:1224-1254, exputil/EmpCylSL.cc:4272-4275, 4567-4571):
// staged through a device buffer of the communicator, blocking.
int expb200_allreduce_sum_host(void *comm, double *host_buf, size_t n)
{
  Comm *c = as_comm(comm);
  REQUIRE(c && (host_buf || n == 0), "bad argument");
  if (n == 0) return EXPB200_OK;
  DeviceGuard g(c->device);
  if (!c->host_stream) CU_TRY(cudaStreamCreateWithFlags(&c->host_stream, cudaStreamNonBlocking));
  if (c->host_stage_cap < n) {
    if (c->host_stage) cudaFree(c->host_stage);
    c->host_stage = nullptr; c->host_stage_cap = 0;
    CU_TRY(cudaMalloc(&c->host_stage, n * sizeof(double)));
    c->host_stage_cap = n;
  }
  CU_TRY(cudaMemcpyAsync(c->host_stage, host_buf, n * sizeof(double), cudaMemcpyHostToDevice, c->host_stream));
  int rc = nccl_allreduce_sum_f64(comm, c->host_stage, n, c->host_stream);
  if (rc) return rc;
  CU_TRY(cudaMemcpyAsync(host_buf, c->host_stage, n * sizeof(double), cudaMemcpyDeviceToHost, c->host_stream));
  CU_TRY(cudaStreamSynchronize(c->host_stream));
  return expb200_comm_check(comm);
}

// Error state of the peer all-reduces issued so far (call after a stream/device sync):
// 0, or EXPB200_ENCCL when one timed out or saw an out-of-order sequence number (its result
// vector was filled with NaN).
int expb200_comm_check(void *comm)
{
  Comm *c = as_comm(comm);
  REQUIRE(c, "not an expb200 communicator");
  if (c->err_host && *c->err_host) {
    set_error(*c->err_host == 2 ? "peer all-reduce: sequence numbers out of order"
                                : "peer all-reduce timed out waiting for another rank");
    return EXPB200_ENCCL;
  }
  return EXPB200_OK;
}

// Peer-memory all-reduce, step 1: allocate this rank's receive buffer and export it
// (cudaIpcMemHandle_t, 64 bytes).  The host exchanges the handles of all ranks (any
// transport: the tests use torch.distributed all_gather) and calls _peer_connect.
int expb200_comm_peer_handle(void *comm, char handle[64])
{
  Comm *c = as_comm(comm);
  REQUIRE(c && handle, "bad argument");
  REQUIRE(c->world <= kPeerMaxWorld, "peer all-reduce supports up to %d ranks", kPeerMaxWorld);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  DeviceGuard g(c->device);
  if (!c->local) {
    CU_TRY(cudaMalloc(&c->local, peer_bytes(c->world)));
    CU_TRY(cudaMemset(c->local, 0, peer_bytes(c->world)));
    CU_TRY(cudaHostAlloc(&c->err_host, sizeof(int), cudaHostAllocMapped));
    *c->err_host = 0;
    CU_TRY(cudaHostGetDevicePointer(&c->err_dev, c->err_host, 0));
    if (const char *e = getenv("EXPB200_PEER_TIMEOUT_S")) c->timeout_ns = strtoull(e, nullptr, 10) * 1000000000ull;
  }
  cudaIpcMemHandle_t h;
  CU_TRY(cudaIpcGetMemHandle(&h, c->local));
  memcpy(handle, &h, 64);
  return EXPB200_OK;
}

// step 2: handles[nranks][64] in rank order.  Collective in effect: every rank must have
// zeroed its buffer (done in _peer_handle) before any rank starts an all-reduce, which
// the handle exchange itself guarantees.
int expb200_comm_peer_connect(void *comm, const char *handles)
{
  Comm *c = as_comm(comm);
  REQUIRE(c && handles && c->local, "call expb200_comm_peer_handle first");
  DeviceGuard g(c->device);
  for (int r = 0; r < c->world; r++) {
    void *base = c->local;
    if (r != c->rank) {
      cudaIpcMemHandle_t h;
      memcpy(&h, handles + (size_t)r * 64, 64);
      CU_TRY(cudaIpcOpenMemHandle(&c->mapped[r], h, cudaIpcMemLazyEnablePeerAccess));
      base = c->mapped[r];
    }
    c->P.flags[r] = static_cast<unsigned long long *>(base);
    c->P.data[r] = reinterpret_cast<double *>(static_cast<char *>(base) + peer_flag_bytes(c->world));
  }
  c->connected = true;
  return EXPB200_OK;
}

} // extern "C"
