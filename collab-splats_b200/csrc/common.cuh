// Shared device/host helpers for the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "rade_config.h"

#define RS_FULL_MASK 0xffffffffu

// launch-status helper: every extern "C" entry returns RS_OK or a negative code and never throws
#define RS_RETURN_LAST_ERROR()                                \
  do {                                                        \
    rs_count_launches(1);                                     \
    cudaError_t e__ = cudaPeekAtLastError();                  \
    if (e__ != cudaSuccess) {                                 \
      rs_set_last_cuda_error((int)e__);                       \
      return RS_ERR_LAUNCH;                                   \
    }                                                         \
    return RS_OK;                                             \
  } while (0)

extern "C" void rs_set_last_cuda_error(int code);
extern "C" void rs_count_launches(int n);
extern "C" int rs_timing_begin(const char* name, void* stream);
extern "C" void rs_timing_end(int span, void* stream);

// RAII span for the optional in-situ timing (api.cu); no-op unless rs_timing_enable(1) was called
struct RsSpan {
  int id;
  void* st;
  RsSpan(const char* name, void* stream) : id(rs_timing_begin(name, stream)), st(stream) {}
  ~RsSpan() { rs_timing_end(id, st); }
};

static inline int rs_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

namespace rs {

// ---- cp.async (LDGSTS): 16-byte global -> shared copies that bypass registers
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// 2^x, approximate (MUFU.EX2)
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// Reduce K (power of two, <= 32) per-lane values across the warp with K-1 + log2(32/K) shuffles
// (recursive halving: at every step each lane keeps half of its slots and sends the other half).
// On return v[0] of lane L holds the warp total of slot (L >> log2(32/K)).
template <int K>
__device__ __forceinline__ void warp_reduce_scatter(float (&v)[K], int lane) {
  static_assert(K >= 1 && K <= 32 && (K & (K - 1)) == 0, "K must be a power of two <= 32");
  int dist = 16;
#pragma unroll
  for (int half = K / 2; half >= 1; half /= 2) {
    const bool upper = (lane & dist) != 0;
#pragma unroll
    for (int j = 0; j < half; ++j) {
      float send = upper ? v[j] : v[j + half];
      float keep = upper ? v[j + half] : v[j];
      v[j] = keep + __shfl_xor_sync(RS_FULL_MASK, send, dist);
    }
    dist >>= 1;
  }
#pragma unroll
  for (int d = 16 / K; d >= 1; d >>= 1) v[0] += __shfl_xor_sync(RS_FULL_MASK, v[0], d);
}

}  // namespace rs
