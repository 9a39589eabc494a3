// Micro-benchmarks behind the design choices of the compositing kernels (run on a B200):
// per-SM issue rates of mma.sync m16n8k8 TF32 (HMMA.1688), SHFL, LDS.128, STS.128, REDG.v4 and FFMA.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters) {
  __shared__ float4 sm[256 * 4];
  const int t = threadIdx.x;
  float acc[4][4] = {};
  unsigned a[4] = {(unsigned)t, 1u, 2u, 3u};
  float4 v = make_float4(t, 1, 2, 3);
  sm[t] = v; sm[t + 256] = v; sm[t + 512] = v; sm[t + 768] = v;
  __syncthreads();
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {          // 4 independent HMMA chains
#pragma unroll
      for (int j = 0; j < 4; ++j) mma(acc[j], a, a[1], a[2]);
    } else if (MODE == 1) {   // 4 independent shuffles
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j][0] += __shfl_xor_sync(0xffffffffu, acc[j][1], 1 + j);
    } else if (MODE == 2) {   // 4 LDS.128, conflict free
#pragma unroll
      for (int j = 0; j < 4; ++j) { float4 x = sm[((t + i) & 255) + 256 * j]; acc[j][0] += x.x; acc[j][1] += x.w; }
    } else if (MODE == 3) {   // 4 STS.128
#pragma unroll
      for (int j = 0; j < 4; ++j) sm[((t + i) & 255) + 256 * j] = v;
    } else if (MODE == 4) {   // 16 FFMA
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[j][q] = fmaf(acc[j][q], 1.0001f, 0.5f);
    } else if (MODE == 6) {   // 4 LDS.64, conflict free
      const float2* s2 = reinterpret_cast<const float2*>(sm);
#pragma unroll
      for (int j = 0; j < 4; ++j) { float2 x = s2[((t + i) & 255) + 256 * j]; acc[j][0] += x.x; acc[j][1] += x.y; }
    } else if (MODE == 7) {   // 4 LDS.32, conflict free
      const float* s1 = reinterpret_cast<const float*>(sm);
#pragma unroll
      for (int j = 0; j < 4; ++j) { float x = s1[((t + i) & 255) + 256 * j]; acc[j][0] += x; }
    } else if (MODE == 8) {   // 4 LDS.128, 20 of 32 lanes active
      if ((t & 31) < 20) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { float4 x = sm[((t + i) & 255) + 256 * j]; acc[j][0] += x.x; acc[j][1] += x.w; }
      }
    } else if (MODE == 9) {   // 4 ldmatrix.x4 (8 rows of 16 bytes per matrix, conflict free)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        unsigned r0, r1, r2, r3;
        const unsigned addr = (unsigned)__cvta_generic_to_shared(sm + (((t & 31) + i) & 255) + 256 * j);
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
        acc[j][0] += __uint_as_float(r0); acc[j][1] += __uint_as_float(r3);
      }
    } else if (MODE == 10) {  // 4 STS.64
      float2* s2 = reinterpret_cast<float2*>(sm);
#pragma unroll
      for (int j = 0; j < 4; ++j) s2[((t + i) & 255) + 256 * j] = make_float2(v.x, v.y);
    } else if (MODE == 5) {   // 4 broadcast LDS.128
#pragma unroll
      for (int j = 0; j < 4; ++j) { float4 x = sm[(i & 255) + 256 * j]; acc[j][0] += x.x; acc[j][1] += x.w; }
    }
  }
  float s = 0.f;
  for (int j = 0; j < 4; ++j) for (int q = 0; q < 4; ++q) s += acc[j][q];
  if (s == 12345.f) out[0] = s + sm[5].x;
}

__global__ void __launch_bounds__(256) red4(float4* dst, int iters, int spread) {
  const int t = blockIdx.x * 256 + threadIdx.x;
  for (int i = 0; i < iters; ++i) atomicAdd(dst + ((t * 4 + i * 1031) % spread), make_float4(1, 2, 3, 4));
}
__global__ void __launch_bounds__(256) red1(float* dst, int iters, int spread) {
  const int t = blockIdx.x * 256 + threadIdx.x;
  for (int i = 0; i < iters; ++i) atomicAdd(dst + ((t * 4 + i * 1031) % spread), 1.f);   // 64-byte runs per half-warp... (stride 4)
}

template <int MODE> void run(const char* name, int per_iter, float* out) {
  const int blocks = 148 * 8, iters = 4096;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, 256>>>(out, 16);
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0); k<MODE><<<blocks, 256>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
  }
  const double warp_instr = (double)blocks * 8 * iters * per_iter;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double cyc = best * 1e-3 * clk * 1e3;
  printf("%-22s %8.3f ms  %7.2f warp-instr/clk/SM (at %d MHz nominal)  -> %.2f clk per warp-instr per SMSP\n", name, best,
         warp_instr / cyc / 148, clk / 1000, cyc * 148 * 4 / warp_instr);
}

int main() {
  float* out; cudaMalloc(&out, 1 << 28);
  cudaMemset(out, 0, 1 << 28);
  run<0>("HMMA.1688 tf32", 4, out);
  run<1>("SHFL", 4, out);
  run<2>("LDS.128", 4, out);
  run<5>("LDS.128 broadcast", 4, out);
  run<6>("LDS.64", 4, out);
  run<7>("LDS.32", 4, out);
  run<8>("LDS.128 20 lanes", 4, out);
  run<9>("ldmatrix.x4", 4, out);
  run<10>("STS.64", 4, out);
  run<3>("STS.128", 4, out);
  run<4>("FFMA", 16, out);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int spread : {1 << 12, 1 << 20, 1 << 24}) {
    float ms;
    cudaEventRecord(e0); red4<<<148 * 8, 256>>>((float4*)out, 256, spread); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("REDG.v4 spread %8d float4: %.3f ms, %.1f G lane-ops/s (%.1f GB/s of payload)\n", spread, ms,
           148.0 * 8 * 256 * 256 / ms / 1e6, 148.0 * 8 * 256 * 256 * 16 / ms / 1e6);
    cudaEventRecord(e0); red1<<<148 * 8, 256>>>(out, 256, spread * 4); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("REDG.32 spread %8d float : %.3f ms, %.1f G lane-ops/s (%.1f GB/s of payload)\n", spread * 4, ms,
           148.0 * 8 * 256 * 256 / ms / 1e6, 148.0 * 8 * 256 * 256 * 4 / ms / 1e6);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
