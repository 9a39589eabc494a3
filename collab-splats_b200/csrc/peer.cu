// Peer-visible device memory for the camera-sharded multi-GPU step (one process per GPU): allocation, CUDA-IPC
// export / import so that a kernel on rank r can read rank g's buffer over NVLink, and a flag handshake
// (signal = remote store, wait = local spin) that orders "rank g has published step s" before "rank r reads it".
// These are set-up / synchronisation entry points, not part of the single-GPU hot path: rs_peer_alloc and
// rs_peer_import are the only functions of the library that allocate or map memory.
#include "common.cuh"
#include <string.h>

namespace {

__global__ void peer_signal_kernel(unsigned long long* const* flag_arrays, int n_ranks, int my_rank,
                                   unsigned long long value) {
  const int g = threadIdx.x;
  if (g >= n_ranks) return;
  __threadfence_system();   // everything this stream wrote before the signal is visible system-wide first
  unsigned long long* dst = flag_arrays[g] + my_rank;
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(value) : "memory");
}

// spins until flags[g] >= value for every g; gives up after ~timeout_ms and raises *timed_out instead of hanging
__global__ void peer_wait_kernel(const unsigned long long* flags, int n_ranks, unsigned long long value,
                                 long long timeout_cycles, int* timed_out) {
  const int g = threadIdx.x;
  if (g >= n_ranks) return;
  const long long t0 = clock64();
  unsigned long long v;
  do {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + g) : "memory");
    if (v >= value) return;
    __nanosleep(200);
  } while (clock64() - t0 < timeout_cycles);
  *timed_out = 1;
}

struct PushDsts {
  uint4* dst[16];
};

// SM-driven push: every destination gets gridDim.x CTAs that stream `src` into it with 16-byte stores (posted writes
// over NVLink), four loads in flight per thread
__global__ void __launch_bounds__(256) peer_push_kernel(PushDsts d, const uint4* __restrict__ src, long long n16) {
  uint4* __restrict__ dst = d.dst[blockIdx.y];
  const long long stride = (long long)gridDim.x * 256;
  long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  for (; i + 3 * stride < n16; i += 4 * stride) {
    const uint4 a = __ldg(src + i), b = __ldg(src + i + stride), c = __ldg(src + i + 2 * stride),
                e = __ldg(src + i + 3 * stride);
    dst[i] = a; dst[i + stride] = b; dst[i + 2 * stride] = c; dst[i + 3 * stride] = e;
  }
  for (; i < n16; i += stride) dst[i] = __ldg(src + i);
}

struct ArBufs {
  float4* buf[16];
};

// Two-shot all-reduce (sum) over peer-mapped buffers, in place, in ONE kernel: rank r owns slice r of the vector.
// Shot 1 (reduce-scatter by pull): it reads slice r of every rank's buffer over NVLink and adds them in rank order
// 0..G-1 -- one fixed order, computed once, so every replica ends up with bit-identical sums.  Shot 2 (all-gather by
// push): it stores the sum into slice r of every rank's buffer.  Slice g of any buffer is read and written by rank g
// only, so no buffer is touched by two ranks at once; the caller brackets the kernel with the flag handshake
// (rs_peer_signal / rs_peer_wait): "all inputs are in place" before, "all sums have landed" after.
template <int G>
__global__ void __launch_bounds__(256) peer_allreduce_kernel(ArBufs b, int rank, long long slice4) {
  const long long lo = (long long)rank * slice4;
  const long long stride = (long long)gridDim.x * 256;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < slice4; i += stride) {
    float4 v[G];
#pragma unroll
    for (int g = 0; g < G; ++g) v[g] = b.buf[g][lo + i];   // G independent loads in flight (G - 1 of them over NVLink)
    float4 acc = v[0];
#pragma unroll
    for (int g = 1; g < G; ++g) { acc.x += v[g].x; acc.y += v[g].y; acc.z += v[g].z; acc.w += v[g].w; }
#pragma unroll
    for (int g = 0; g < G; ++g) b.buf[g][lo + i] = acc;     // posted stores
  }
}

}  // namespace

// In-place sum of `n_floats` fp32 values over n_ranks peer-mapped buffers (bufs: HOST array, bufs[g] = rank g's buffer
// as mapped in this process, bufs[my_rank] = the local one).  n_floats must be a multiple of 4 * n_ranks (pad with
// zeros); n_ranks in {2, 4, 8}.  The caller orders it with the flag handshake: every rank has finished writing its
// buffer before any rank launches this, and nobody reads a buffer before every rank's launch has completed.
extern "C" int rs_peer_allreduce(void* const* bufs, int n_ranks, int my_rank, long long n_floats, int ctas,
                                 void* stream) {
  RsSpan span__("rs_peer_allreduce", stream);
  if (!bufs || my_rank < 0 || my_rank >= n_ranks || n_floats < 0 || ctas <= 0) return RS_ERR_BAD_ARG;
  if (n_ranks != 2 && n_ranks != 4 && n_ranks != 8) return RS_ERR_UNSUPPORTED;
  if (n_floats % (4ll * n_ranks)) return RS_ERR_BAD_ARG;
  if (n_floats == 0) return RS_OK;
  ArBufs b;
  for (int g = 0; g < n_ranks; ++g) {
    if (!bufs[g]) return RS_ERR_BAD_ARG;
    b.buf[g] = (float4*)bufs[g];
  }
  const long long slice4 = n_floats / 4 / n_ranks;
  cudaStream_t st = (cudaStream_t)stream;
  if (n_ranks == 2) peer_allreduce_kernel<2><<<ctas, 256, 0, st>>>(b, my_rank, slice4);
  else if (n_ranks == 4) peer_allreduce_kernel<4><<<ctas, 256, 0, st>>>(b, my_rank, slice4);
  else peer_allreduce_kernel<8><<<ctas, 256, 0, st>>>(b, my_rank, slice4);
  RS_RETURN_LAST_ERROR();
}

extern "C" int rs_peer_alloc(long long bytes, void** ptr) {
  if (bytes <= 0 || !ptr) return RS_ERR_BAD_ARG;
  cudaError_t e = cudaMalloc(ptr, (size_t)bytes);
  if (e == cudaSuccess) e = cudaMemset(*ptr, 0, (size_t)bytes);
  if (e != cudaSuccess) { rs_set_last_cuda_error((int)e); return RS_ERR_LAUNCH; }
  return RS_OK;
}

extern "C" int rs_peer_free(void* ptr) {
  if (!ptr) return RS_OK;
  cudaError_t e = cudaFree(ptr);
  if (e != cudaSuccess) { rs_set_last_cuda_error((int)e); return RS_ERR_LAUNCH; }
  return RS_OK;
}

extern "C" int rs_peer_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

extern "C" int rs_peer_export(void* ptr, void* handle_out) {
  if (!ptr || !handle_out) return RS_ERR_BAD_ARG;
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
  if (e != cudaSuccess) { rs_set_last_cuda_error((int)e); return RS_ERR_LAUNCH; }
  memcpy(handle_out, &h, sizeof(h));
  return RS_OK;
}

extern "C" int rs_peer_import(const void* handle, void** ptr) {
  if (!handle || !ptr) return RS_ERR_BAD_ARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) { rs_set_last_cuda_error((int)e); return RS_ERR_LAUNCH; }
  return RS_OK;
}

extern "C" int rs_peer_unimport(void* ptr) {
  if (!ptr) return RS_OK;
  cudaError_t e = cudaIpcCloseMemHandle(ptr);
  if (e != cudaSuccess) { rs_set_last_cuda_error((int)e); return RS_ERR_LAUNCH; }
  return RS_OK;
}

// device-to-device copy between peer-visible buffers on the copy engines (no SM involvement): the "push" form of
// the gradient exchange sends this rank's region into every peer's inbox while the SMs run the projection VJP
extern "C" int rs_peer_copy(void* dst, const void* src, long long bytes, void* stream) {
  if (bytes < 0 || (bytes > 0 && (!dst || !src))) return RS_ERR_BAD_ARG;
  if (bytes == 0) return RS_OK;
  cudaError_t e = cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, (cudaStream_t)stream);
  if (e != cudaSuccess) { rs_set_last_cuda_error((int)e); return RS_ERR_LAUNCH; }
  return RS_OK;
}

// flag_arrays (DEVICE array of n_ranks pointers): rank g's flag array (u64[n_ranks]) as mapped in this process;
// writes `value` into slot my_rank of every rank's array, after all earlier work of `stream`.
extern "C" int rs_peer_signal(void* const* flag_arrays_dev, int n_ranks, int my_rank, unsigned long long value,
                              void* stream) {
  RsSpan span__("rs_peer_signal", stream);
  if (!flag_arrays_dev || n_ranks <= 0 || n_ranks > 32 || my_rank < 0 || my_rank >= n_ranks) return RS_ERR_BAD_ARG;
  peer_signal_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((unsigned long long* const*)flag_arrays_dev, n_ranks, my_rank,
                                                         value);
  RS_RETURN_LAST_ERROR();
}

// blocks `stream` (on the device) until every slot of the LOCAL flag array is >= value; *timed_out_dev (device int,
// zeroed by the caller) is raised instead of hanging if a peer does not arrive within timeout_ms.
extern "C" int rs_peer_wait(const void* local_flags, int n_ranks, unsigned long long value, int timeout_ms,
                            int* timed_out_dev, void* stream) {
  RsSpan span__("rs_peer_wait", stream);
  if (!local_flags || n_ranks <= 0 || n_ranks > 32 || !timed_out_dev || timeout_ms <= 0) return RS_ERR_BAD_ARG;
  const long long cycles = (long long)timeout_ms * 1900000ll;   // ~1.9 GHz SM clock
  peer_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((const unsigned long long*)local_flags, n_ranks, value, cycles,
                                                       timed_out_dev);
  RS_RETURN_LAST_ERROR();
}

// The same transfer as n_dst rs_peer_copy calls, driven by the SMs instead of the copy engines: one kernel,
// ctas_per_dst CTAs per destination (dsts: HOST array of n_dst <= 16 peer-mapped device pointers; bytes % 16 == 0).
// On an NVSwitch box the copy engines deliver ~380 GB/s for 7 x 16 MB, SM stores more.
extern "C" int rs_peer_push(void* const* dsts, int n_dst, const void* src, long long bytes, int ctas_per_dst,
                            void* stream) {
  RsSpan span__("rs_peer_push", stream);
  if (n_dst < 0 || n_dst > 16 || bytes < 0 || (bytes & 15) || ctas_per_dst <= 0) return RS_ERR_BAD_ARG;
  if (n_dst == 0 || bytes == 0) return RS_OK;
  if (!dsts || !src) return RS_ERR_BAD_ARG;
  PushDsts d;
  for (int i = 0; i < n_dst; ++i) {
    if (!dsts[i]) return RS_ERR_BAD_ARG;
    d.dst[i] = (uint4*)dsts[i];
  }
  peer_push_kernel<<<dim3(ctas_per_dst, n_dst), 256, 0, (cudaStream_t)stream>>>(d, (const uint4*)src, bytes / 16);
  RS_RETURN_LAST_ERROR();
}
