// profile.cu -- see profile.cuh
#include "profile.cuh"

#include <vector>

namespace expb200 {
namespace {

struct Pair { cudaEvent_t a, b; int id; };
bool g_enabled = false;
unsigned long long g_launches[K_COUNT] = {};
double g_ms[K_COUNT] = {};
unsigned long long g_timed[K_COUNT] = {};
std::vector<Pair> g_pending;
std::vector<Pair> g_free;

const char *kNames[K_COUNT] = {"key", "sort", "segments", "deposit", "combine", "contract",
                               "field", "force", "kick", "drift", "zero", "levels", "allreduce",
                               "pack", "covariance", "cyl_deposit", "cyl_force", "cyl_field", "kick_drift_key", "reorder", "other"};

void resolve()
{
  for (auto &p : g_pending) {
    cudaEventSynchronize(p.b);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
      g_ms[p.id] += ms;
      g_timed[p.id]++;
    }
    g_free.push_back(p);
  }
  g_pending.clear();
}

} // namespace

ProfRegion::ProfRegion(KernelId id_, cudaStream_t st_, int launches) : id(id_), st(st_), slot(-1)
{
  g_launches[id] += launches;
  if (!g_enabled) return;
  if (g_pending.size() >= 8192) resolve();
  Pair p;
  if (!g_free.empty()) { p = g_free.back(); g_free.pop_back(); }
  else { cudaEventCreate(&p.a); cudaEventCreate(&p.b); }
  p.id = id;
  cudaEventRecord(p.a, st);
  g_pending.push_back(p);
  slot = (int)g_pending.size() - 1;
}

ProfRegion::~ProfRegion()
{
  if (slot >= 0 && slot < (int)g_pending.size()) cudaEventRecord(g_pending[slot].b, st);
}

} // namespace expb200

using namespace expb200;

extern "C" {

int expb200_profile_enable(int on)
{
  if (!on) resolve();
  g_enabled = on != 0;
  return EXPB200_OK;
}

int expb200_profile_reset(void)
{
  resolve();
  for (int k = 0; k < K_COUNT; k++) { g_launches[k] = 0; g_ms[k] = 0.0; g_timed[k] = 0; }
  return EXPB200_OK;
}

int expb200_profile_kernels(void) { return K_COUNT; }

const char *expb200_profile_name(int id) { return id >= 0 && id < K_COUNT ? kNames[id] : ""; }

/* launches[k]: kernel launches since the last reset (always counted);
 * ms[k], timed[k]: summed event time and number of timed regions (only while enabled). */
int expb200_profile_read(unsigned long long *launches, double *ms, unsigned long long *timed)
{
  resolve();
  for (int k = 0; k < K_COUNT; k++) {
    if (launches) launches[k] = g_launches[k];
    if (ms) ms[k] = g_ms[k];
    if (timed) timed[k] = g_timed[k];
  }
  return EXPB200_OK;
}

} // extern "C"

// ---------------------------------------------------------------------------
// FP64 pipe peak: dependency-free DFMA chains (MEASURED_PEAKS.json carries HBM and
// bf16 figures only; the sphereSL kernels are bound by the FP64 FMA pipe, so the
// roofline denominator is measured here, on the same device, by the same process).
// ---------------------------------------------------------------------------
namespace expb200 {
namespace {
__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, int iters, double a, double b)
{
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5,
         x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (s == 123.456) out[0] = s;     // never true; keeps the chains alive
}
} // namespace
} // namespace expb200

extern "C" int expb200_fp64_peak(int device, double *tflops, double *sm_mhz_hint)
{
  REQUIRE(tflops, "tflops is NULL");
  DeviceGuard g(device);
  double *d = nullptr;
  CU_TRY(cudaMalloc(&d, sizeof(double)));
  cudaEvent_t a, b;
  CU_TRY(cudaEventCreate(&a));
  CU_TRY(cudaEventCreate(&b));
  const int blocks = kSMsB200 * 8, threads = 256, iters = 4096;
  double best = 0.0;
  for (int rep = 0; rep < 5; rep++) {
    CU_TRY(cudaEventRecord(a, 0));
    dfma_peak_kernel<<<blocks, threads>>>(d, iters, 0.999999, 1.0e-9);
    CU_TRY(cudaEventRecord(b, 0));
    CU_TRY(cudaEventSynchronize(b));
    float ms = 0.f;
    CU_TRY(cudaEventElapsedTime(&ms, a, b));
    double flops = 2.0 * 64.0 * iters * (double)blocks * threads;
    double tf = flops / (ms * 1.0e-3) / 1.0e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(d);
  *tflops = best;
  if (sm_mhz_hint) *sm_mhz_hint = best * 1.0e12 / (2.0 * 64.0 * kSMsB200) / 1.0e6;   // if 64 DFMA/clk/SM
  return EXPB200_OK;
}

// ---------------------------------------------------------------------------
// FP64 tensor-core (DMMA) peak by instruction shape: the roofline denominator of the covariance
// kernels and the evidence for which mma.sync f64 shape to use on sm_100a.
//   shape 0: m8n8k4 (the sm_80 shape), 1: m16n8k4, 2: m16n8k8, 3: m16n8k16
// ---------------------------------------------------------------------------
namespace expb200 {
namespace {
template <int SHAPE>
__global__ void __launch_bounds__(256) dmma_peak_kernel(double *out, int iters, double a0, double b0)
{
  double c[4][4];
#pragma unroll
  for (int k = 0; k < 4; k++) { c[k][0] = c[k][1] = c[k][2] = c[k][3] = threadIdx.x * 1e-9 + k; }
  double a[8], b[4];
#pragma unroll
  for (int k = 0; k < 8; k++) a[k] = a0 + k * 1e-3;
#pragma unroll
  for (int k = 0; k < 4; k++) b[k] = b0 + k * 1e-3;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 4; k++) {       // four independent accumulator tiles
      if (SHAPE == 0)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[k][0]), "+d"(c[k][1]) : "d"(a[0]), "d"(b[0]));
      else if (SHAPE == 1)
        asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+d"(c[k][0]), "+d"(c[k][1]), "+d"(c[k][2]), "+d"(c[k][3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
      else if (SHAPE == 2)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+d"(c[k][0]), "+d"(c[k][1]), "+d"(c[k][2]), "+d"(c[k][3])
                     : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
      else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
                     "{%12,%13,%14,%15}, {%0,%1,%2,%3};"
                     : "+d"(c[k][0]), "+d"(c[k][1]), "+d"(c[k][2]), "+d"(c[k][3])
                     : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                       "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
    }
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 4; k++) s += c[k][0] + c[k][1] + c[k][2] + c[k][3];
  if (s == 123.456) out[0] = s;
}
} // namespace
} // namespace expb200

extern "C" int expb200_dmma_peak(int device, int shape, double *tflops)
{
  REQUIRE(tflops && shape >= 0 && shape <= 3, "bad argument");
  DeviceGuard g(device);
  double *d = nullptr;
  CU_TRY(cudaMalloc(&d, sizeof(double)));
  cudaEvent_t a, b;
  CU_TRY(cudaEventCreate(&a));
  CU_TRY(cudaEventCreate(&b));
  const int blocks = kSMsB200 * 8, threads = 256, iters = 2048;
  static const double fma_per_instr[4] = {8.0 * 8 * 4, 16.0 * 8 * 4, 16.0 * 8 * 8, 16.0 * 8 * 16};
  double best = 0.0;
  for (int rep = 0; rep < 4; rep++) {
    CU_TRY(cudaEventRecord(a, 0));
    switch (shape) {
      case 0: dmma_peak_kernel<0><<<blocks, threads>>>(d, iters, 1.0, 1e-9); break;
      case 1: dmma_peak_kernel<1><<<blocks, threads>>>(d, iters, 1.0, 1e-9); break;
      case 2: dmma_peak_kernel<2><<<blocks, threads>>>(d, iters, 1.0, 1e-9); break;
      default: dmma_peak_kernel<3><<<blocks, threads>>>(d, iters, 1.0, 1e-9); break;
    }
    CU_TRY(cudaEventRecord(b, 0));
    CU_TRY(cudaEventSynchronize(b));
    float ms = 0.f;
    CU_TRY(cudaEventElapsedTime(&ms, a, b));
    const double flops = 2.0 * fma_per_instr[shape] * 4.0 * iters * (double)blocks * (threads / 32);
    const double tf = flops / (ms * 1.0e-3) / 1.0e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(d);
  *tflops = best;
  return EXPB200_OK;
}
