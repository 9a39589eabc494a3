// EWA projection of 3-D Gaussians with RaDe ray-space depth/normal terms, forward and backward.
// Replaces gsplat-rade `fully_fused_projection` (packed=False, pinhole): SURVEY.md rows a5/a12,
// reference call sites collab_splats/models/rade_gs_model.py:373-389 (direct) and :439-465 (via
// rasterization()).  The per-element math lives in rade_math.cuh (shared with the host tests).
//
// HBM-bound streaming kernels: 104 B per (camera, Gaussian) forward, 136 B backward.  One thread per
// (camera, Gaussian); 3-float rows (means, scales, conics, normals) are moved through shared memory
// so that every global access is a coalesced 16-byte vector access.
#include "common.cuh"
#include "rade_math.cuh"

namespace {

constexpr int PB = 256;  // threads per block

// Cooperative, coalesced load/store of `count` rows of 3 floats through shared memory (row-major).
// Full blocks whose global base is 16-byte aligned move float4s; tails and odd bases move scalars.
__device__ __forceinline__ void load_rows3(float* s, const float* __restrict__ src, int count, int t) {
  if (count == PB && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    if (t < PB * 3 / 4) reinterpret_cast<float4*>(s)[t] = __ldg(reinterpret_cast<const float4*>(src) + t);
  } else {
    for (int i = t; i < count * 3; i += PB) s[i] = __ldg(src + i);
  }
}
__device__ __forceinline__ void store_rows3(const float* s, float* __restrict__ dst, int count, int t) {
  if (count == PB && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    if (t < PB * 3 / 4) reinterpret_cast<float4*>(dst)[t] = reinterpret_cast<const float4*>(s)[t];
  } else {
    for (int i = t; i < count * 3; i += PB) dst[i] = s[i];
  }
}

__device__ __forceinline__ void load_cam(const float* __restrict__ viewmats, const float* __restrict__ Ks, int c,
                                         rs::Cam<float>& cam) {
  const float* vm = viewmats + (size_t)c * 16;
  const float* K = Ks + (size_t)c * 9;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) cam.W[i * 3 + j] = __ldg(vm + i * 4 + j);
    cam.t[i] = __ldg(vm + i * 4 + 3);
  }
  cam.fx = __ldg(K + 0); cam.fy = __ldg(K + 4); cam.cx = __ldg(K + 2); cam.cy = __ldg(K + 5);
}

__global__ void __launch_bounds__(PB)
project_fwd_kernel(const float* __restrict__ means, const float* __restrict__ quats,
                   const float* __restrict__ scales, const float* __restrict__ viewmats,
                   const float* __restrict__ Ks, int C, int N, rs::ProjParams<float> pp, int calc_comp,
                   int32_t* __restrict__ radii, float* __restrict__ means2d, float* __restrict__ depths,
                   float* __restrict__ conics, float* __restrict__ comps, float* __restrict__ ray_ts,
                   float* __restrict__ ray_planes, float* __restrict__ normals) {
  __shared__ __align__(16) float sbuf[PB * 3];
  const int t = threadIdx.x;
  const int n0 = blockIdx.x * PB;
  const int c = blockIdx.y;
  const int count = min(PB, N - n0);
  const int n = n0 + t;
  const bool active = t < count;

  float mean[3] = {0.f, 0.f, 1.f}, scale[3] = {1.f, 1.f, 1.f}, quat[4] = {1.f, 0.f, 0.f, 0.f};
  load_rows3(sbuf, means + (size_t)n0 * 3, count, t);
  __syncthreads();
  if (active) { mean[0] = sbuf[t * 3]; mean[1] = sbuf[t * 3 + 1]; mean[2] = sbuf[t * 3 + 2]; }
  __syncthreads();
  load_rows3(sbuf, scales + (size_t)n0 * 3, count, t);
  __syncthreads();
  if (active) { scale[0] = sbuf[t * 3]; scale[1] = sbuf[t * 3 + 1]; scale[2] = sbuf[t * 3 + 2]; }
  __syncthreads();
  if (active) {
    float4 q = __ldg(reinterpret_cast<const float4*>(quats) + n);
    quat[0] = q.x; quat[1] = q.y; quat[2] = q.z; quat[3] = q.w;
  }
  rs::Cam<float> cam;
  load_cam(viewmats, Ks, c, cam);
  rs::ProjOut<float> o;
  rs::project_fwd_one(mean, quat, scale, cam, pp, o);

  const size_t e = (size_t)c * N + n;
  const size_t e0 = (size_t)c * N + n0;
  if (active) {
    reinterpret_cast<int2*>(radii)[e] = make_int2(o.rx, o.ry);
    reinterpret_cast<float2*>(means2d)[e] = make_float2(o.m2x, o.m2y);
    depths[e] = o.depth;
    if (calc_comp) comps[e] = o.comp;
    ray_ts[e] = o.ray_t;
    reinterpret_cast<float2*>(ray_planes)[e] = make_float2(o.rp0, o.rp1);
    sbuf[t * 3] = o.ca; sbuf[t * 3 + 1] = o.cb; sbuf[t * 3 + 2] = o.cc;
  }
  __syncthreads();
  store_rows3(sbuf, conics + e0 * 3, count, t);
  __syncthreads();
  if (active) { sbuf[t * 3] = o.nx; sbuf[t * 3 + 1] = o.ny; sbuf[t * 3 + 2] = o.nz; }
  __syncthreads();
  store_rows3(sbuf, normals + e0 * 3, count, t);
}

// One thread per Gaussian, looping over cameras: the per-Gaussian gradients are summed over C in
// registers (no atomics).  Camera gradients (optional) are warp-reduced then atomically added.
// Capped at 128 registers (two 256-thread CTAs per SM): the VJP is ~1500 instructions per element, so the
// kernel is latency-bound rather than HBM-bound unless enough warps are resident.
__global__ void __launch_bounds__(PB, 2)
project_bwd_kernel(const float* __restrict__ means, const float* __restrict__ quats,
                   const float* __restrict__ scales, const float* __restrict__ viewmats,
                   const float* __restrict__ Ks, int C, int N, rs::ProjParams<float> pp,
                   const float* __restrict__ v_means2d, const float* __restrict__ v_depths,
                   const float* __restrict__ v_conics, const float* __restrict__ v_comps,
                   const float* __restrict__ v_ray_ts, const float* __restrict__ v_ray_planes,
                   const float* __restrict__ v_normals, float* __restrict__ v_means,
                   float* __restrict__ v_quats, float* __restrict__ v_scales, float* __restrict__ v_viewmats) {
  __shared__ __align__(16) float sbuf[PB * 3];
  const int t = threadIdx.x;
  const int n0 = blockIdx.x * PB;
  const int count = min(PB, N - n0);
  const int n = n0 + t;
  const bool active = t < count;
  const int lane = t & 31;

  float mean[3] = {0.f, 0.f, 1.f}, scale[3] = {1.f, 1.f, 1.f}, quat[4] = {1.f, 0.f, 0.f, 0.f};
  load_rows3(sbuf, means + (size_t)n0 * 3, count, t);
  __syncthreads();
  if (active) { mean[0] = sbuf[t * 3]; mean[1] = sbuf[t * 3 + 1]; mean[2] = sbuf[t * 3 + 2]; }
  __syncthreads();
  load_rows3(sbuf, scales + (size_t)n0 * 3, count, t);
  __syncthreads();
  if (active) { scale[0] = sbuf[t * 3]; scale[1] = sbuf[t * 3 + 1]; scale[2] = sbuf[t * 3 + 2]; }
  __syncthreads();
  if (active) {
    float4 q = __ldg(reinterpret_cast<const float4*>(quats) + n);
    quat[0] = q.x; quat[1] = q.y; quat[2] = q.z; quat[3] = q.w;
  }
  float vm[3] = {0.f, 0.f, 0.f}, vq[4] = {0.f, 0.f, 0.f, 0.f}, vs[3] = {0.f, 0.f, 0.f};
  for (int c = 0; c < C; ++c) {
    rs::Cam<float> cam;
    load_cam(viewmats, Ks, c, cam);
    float vW[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, vt[3] = {0.f, 0.f, 0.f};
    if (active) {
      const size_t e = (size_t)c * N + n;
      rs::ProjGradIn<float> g;
      float2 a = __ldg(reinterpret_cast<const float2*>(v_means2d) + e);
      g.v_m2x = a.x; g.v_m2y = a.y;
      g.v_depth = v_depths ? __ldg(v_depths + e) : 0.f;
      g.v_ca = __ldg(v_conics + e * 3); g.v_cb = __ldg(v_conics + e * 3 + 1); g.v_cc = __ldg(v_conics + e * 3 + 2);
      g.v_comp = v_comps ? __ldg(v_comps + e) : 0.f;
      g.v_ray_t = v_ray_ts ? __ldg(v_ray_ts + e) : 0.f;
      if (v_ray_planes) {
        float2 b = __ldg(reinterpret_cast<const float2*>(v_ray_planes) + e);
        g.v_rp0 = b.x; g.v_rp1 = b.y;
      } else {
        g.v_rp0 = g.v_rp1 = 0.f;
      }
      if (v_normals) {
        g.v_nx = __ldg(v_normals + e * 3); g.v_ny = __ldg(v_normals + e * 3 + 1); g.v_nz = __ldg(v_normals + e * 3 + 2);
      } else {
        g.v_nx = g.v_ny = g.v_nz = 0.f;
      }
      rs::project_bwd_one(mean, quat, scale, cam, pp, g, vm, vq, vs, v_viewmats ? vW : (float*)nullptr, vt);
    }
    if (v_viewmats) {  // warp-uniform branch
      float r[16];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) r[i * 4 + j] = vW[i * 3 + j];
        r[i * 4 + 3] = vt[i];
      }
      r[12] = r[13] = r[14] = r[15] = 0.f;
      rs::warp_reduce_scatter<16>(r, lane);
      if ((lane & 1) == 0 && (lane >> 1) < 12) atomicAdd(v_viewmats + (size_t)c * 16 + (lane >> 1), r[0]);
    }
  }
  // coalesced writes of the 3-float rows through shared memory (v_means/v_scales start 16-byte aligned)
  if (active) { sbuf[t * 3] = vm[0]; sbuf[t * 3 + 1] = vm[1]; sbuf[t * 3 + 2] = vm[2]; }
  __syncthreads();
  store_rows3(sbuf, v_means + (size_t)n0 * 3, count, t);
  __syncthreads();
  if (active) { sbuf[t * 3] = vs[0]; sbuf[t * 3 + 1] = vs[1]; sbuf[t * 3 + 2] = vs[2]; }
  __syncthreads();
  store_rows3(sbuf, v_scales + (size_t)n0 * 3, count, t);
  if (active) reinterpret_cast<float4*>(v_quats)[n] = make_float4(vq[0], vq[1], vq[2], vq[3]);
}

}  // namespace

extern "C" int rs_project_fwd(const float* means, const float* quats, const float* scales, const float* viewmats,
                              const float* Ks, int C, int N, int width, int height, float eps2d, float near_plane,
                              float far_plane, float radius_clip, int calc_compensations, int32_t* radii,
                              float* means2d, float* depths, float* conics, float* compensations, float* ray_ts,
                              float* ray_planes, float* normals, void* stream) {
  RsSpan span__("rs_project_fwd", stream);
  if (C < 0 || N < 0 || width <= 0 || height <= 0) return RS_ERR_BAD_ARG;
  if (C == 0 || N == 0) return RS_OK;
  if (!means || !quats || !scales || !viewmats || !Ks || !radii || !means2d || !depths || !conics || !ray_ts ||
      !ray_planes || !normals || (calc_compensations && !compensations))
    return RS_ERR_BAD_ARG;
  rs::ProjParams<float> pp{(float)width, (float)height, eps2d, near_plane, far_plane, radius_clip};
  dim3 grid(rs_div_up(N, PB), C);
  project_fwd_kernel<<<grid, PB, 0, (cudaStream_t)stream>>>(means, quats, scales, viewmats, Ks, C, N, pp,
                                                           calc_compensations, radii, means2d, depths, conics,
                                                           compensations, ray_ts, ray_planes, normals);
  RS_RETURN_LAST_ERROR();
}

extern "C" int rs_project_bwd(const float* means, const float* quats, const float* scales, const float* viewmats,
                              const float* Ks, int C, int N, int width, int height, float eps2d, float near_plane,
                              float far_plane, float radius_clip, const float* v_means2d, const float* v_depths,
                              const float* v_conics, const float* v_compensations, const float* v_ray_ts,
                              const float* v_ray_planes, const float* v_normals, float* v_means, float* v_quats,
                              float* v_scales, float* v_viewmats, void* stream) {
  RsSpan span__("rs_project_bwd", stream);
  if (C < 0 || N < 0 || width <= 0 || height <= 0) return RS_ERR_BAD_ARG;
  if (N == 0) return RS_OK;
  if (!means || !quats || !scales || !viewmats || !Ks || !v_means2d || !v_conics || !v_means || !v_quats ||
      !v_scales)
    return RS_ERR_BAD_ARG;
  rs::ProjParams<float> pp{(float)width, (float)height, eps2d, near_plane, far_plane, radius_clip};
  if (v_viewmats) {
    cudaError_t e = cudaMemsetAsync(v_viewmats, 0, sizeof(float) * 16 * (size_t)C, (cudaStream_t)stream);
    if (e != cudaSuccess) { rs_set_last_cuda_error((int)e); return RS_ERR_LAUNCH; }
  }
  project_bwd_kernel<<<rs_div_up(N, PB), PB, 0, (cudaStream_t)stream>>>(
      means, quats, scales, viewmats, Ks, C, N, pp, v_means2d, v_depths, v_conics, v_compensations, v_ray_ts,
      v_ray_planes, v_normals, v_means, v_quats, v_scales, v_viewmats);
  RS_RETURN_LAST_ERROR();
}
