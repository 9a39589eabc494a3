// Library-level entry points of the C-ABI (see include/rade_b200.h).
#include "common.cuh"

static thread_local int g_last_cuda_error = 0;

extern "C" void rs_set_last_cuda_error(int code) { g_last_cuda_error = code; }

// number of kernel launches issued through this library (bench.py reports it as gpu_launches)
static unsigned long long g_launches = 0;
extern "C" void rs_count_launches(int n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }
extern "C" unsigned long long rs_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

extern "C" int rs_version(void) { return 100; }  // 0.1.0

extern "C" int rs_last_cuda_error(void) { return g_last_cuda_error; }

extern "C" const char* rs_error_string(int status) {
  switch (status) {
    case RS_OK: return "ok";
    case RS_ERR_BAD_ARG: return "bad argument (null pointer, negative size or unsupported shape)";
    case RS_ERR_LAUNCH: return cudaGetErrorString((cudaError_t)g_last_cuda_error);
    case RS_ERR_UNSUPPORTED: return "unsupported configuration (see include/rade_b200.h)";
    default: return "unknown status";
  }
}

// ------------------------------------------------------------------------------------------------ in-situ timing
// Optional per-entry-point device timing: when enabled, every extern "C" launcher brackets its launches with a
// pair of CUDA events on the launching stream.  rs_timing_collect() synchronises those events and returns the
// accumulated milliseconds per entry point.  bench.py uses it to time each kernel *inside* the real step
// (same inputs, same cache state) instead of in isolation.
#include <string.h>
namespace {
constexpr int kMaxSpans = 4096;
struct Span { const char* name; cudaEvent_t e0, e1; };
Span g_spans[kMaxSpans];
int g_n_spans = 0;
int g_timing_on = 0;
cudaEvent_t g_pool[2 * kMaxSpans];
int g_pool_ready = 0;
}  // namespace

extern "C" void rs_timing_enable(int on) {
  if (on && !g_pool_ready) {
    for (int i = 0; i < 2 * kMaxSpans; ++i) cudaEventCreate(&g_pool[i]);
    g_pool_ready = 1;
  }
  g_timing_on = on;
  g_n_spans = 0;
}

extern "C" int rs_timing_begin(const char* name, void* stream) {
  if (!g_timing_on || g_n_spans >= kMaxSpans) return -1;
  const int i = g_n_spans++;
  g_spans[i].name = name;
  g_spans[i].e0 = g_pool[2 * i];
  g_spans[i].e1 = g_pool[2 * i + 1];
  cudaEventRecord(g_spans[i].e0, (cudaStream_t)stream);
  return i;
}

extern "C" void rs_timing_end(int span, void* stream) {
  if (span >= 0) cudaEventRecord(g_spans[span].e1, (cudaStream_t)stream);
}

// Writes up to `cap` (name, total ms, calls) rows; returns the number of distinct names.  Resets the span list.
extern "C" int rs_timing_collect(char* names /* cap x 48 bytes */, float* ms, int* calls, int cap) {
  int n = 0;
  for (int i = 0; i < g_n_spans; ++i) {
    cudaEventSynchronize(g_spans[i].e1);
    float t = 0.f;
    cudaEventElapsedTime(&t, g_spans[i].e0, g_spans[i].e1);
    int k = 0;
    for (; k < n; ++k)
      if (strncmp(names + 48 * k, g_spans[i].name, 47) == 0) break;
    if (k == n) {
      if (n >= cap) continue;
      strncpy(names + 48 * n, g_spans[i].name, 47);
      names[48 * n + 47] = 0;
      ms[n] = 0.f;
      calls[n] = 0;
      ++n;
    }
    ms[k] += t;
    calls[k] += 1;
  }
  g_n_spans = 0;
  return n;
}

// ------------------------------------------------------------------------------------------------ FP32 peak probe
// Sustained FP32 FMA throughput of this device (the denominator for the compositing kernels' FP32 roofline,
// BASELINE.md section 1): 16 independent FFMA chains per thread, 2 flops per FFMA.
__global__ void __launch_bounds__(256) fma_peak_kernel(int iters, float* out) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = 1.0f + 1e-3f * (threadIdx.x + i);
  const float b = 1.000001f, c = 1e-7f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], b, c);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  if (s == 123.456f) out[0] = s;  // never true: keeps the chains alive
}

// Launches the probe; flops executed = blocks * 256 * iters * 16 * 2.  The caller times it with CUDA events.
extern "C" int rs_fma_peak_probe(int blocks, int iters, float* out, void* stream) {
  if (blocks <= 0 || iters <= 0 || !out) return RS_ERR_BAD_ARG;
  fma_peak_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, out);
  RS_RETURN_LAST_ERROR();
}

// ---- small device utilities used by the host layer in place of framework kernels
namespace {
struct ScaleList {
  float* p[8];
  long long n[8];
  int count;
};

// x *= *g for every listed buffer, skipped altogether when *g == 1 (the usual upstream gradient of a loss): the test
// is made on the device, so the host never reads the value
__global__ void __launch_bounds__(256) scale_unless_one_kernel(ScaleList L, const float* __restrict__ g) {
  const float s = __ldg(g);
  if (s == 1.f) return;
  const long long stride = (long long)gridDim.x * 256 * 4;
  for (int b = 0; b < L.count; ++b) {
    float* x = L.p[b];
    const long long n = L.n[b], n4 = n & ~3ll;
    for (long long i = ((long long)blockIdx.x * 256 + threadIdx.x) * 4; i < n4; i += stride) {
      float4 v = *reinterpret_cast<float4*>(x + i);
      v.x *= s; v.y *= s; v.z *= s; v.w *= s;
      *reinterpret_cast<float4*>(x + i) = v;
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(n - n4)) x[n4 + threadIdx.x] *= s;
  }
}
}  // namespace

// In-place x_b *= *scale for up to 8 fp32 buffers (16-byte aligned), one launch; a no-op pass when *scale == 1.
extern "C" int rs_scale_unless_one(float* const* bufs, const long long* counts, int n_bufs, const float* scale,
                                   void* stream) {
  RsSpan span__("rs_scale_unless_one", stream);
  if (n_bufs < 0 || n_bufs > 8 || !scale || (n_bufs > 0 && (!bufs || !counts))) return RS_ERR_BAD_ARG;
  if (n_bufs == 0) return RS_OK;
  ScaleList L;
  L.count = n_bufs;
  for (int i = 0; i < n_bufs; ++i) {
    if (!bufs[i] || counts[i] < 0 || ((uintptr_t)bufs[i] & 15)) return RS_ERR_BAD_ARG;
    L.p[i] = bufs[i]; L.n[i] = counts[i];
  }
  scale_unless_one_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(L, scale);
  RS_RETURN_LAST_ERROR();
}

// cudaMemsetAsync(ptr, 0, bytes) on the caller's stream (the copy-engine / driver fill, faster than a framework fill
// kernel for the 64 MB gradient records)
extern "C" int rs_zero_bytes(void* ptr, long long bytes, void* stream) {
  if (bytes < 0 || (bytes > 0 && !ptr)) return RS_ERR_BAD_ARG;
  if (bytes == 0) return RS_OK;
  cudaError_t e = cudaMemsetAsync(ptr, 0, (size_t)bytes, (cudaStream_t)stream);
  if (e != cudaSuccess) { rs_set_last_cuda_error((int)e); return RS_ERR_LAUNCH; }
  return RS_OK;
}
