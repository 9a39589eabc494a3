// Per-Gaussian bookkeeping that follows the rasterizer in the training step:
//  * densify_stats_kernel: the statistics gsplat's DefaultStrategy accumulates after every backward
//    (collab_splats/models/rade_gs_model.py:191-198 -> strategy.step_pre/post_backward): sum of the 2-D
//    positional gradient norms (in normalised device coordinates, scaled by the number of cameras), visibility
//    counts and the running maximum screen radius -- one pass, thread per Gaussian looping over the cameras, no
//    atomics, instead of ~12 torch ops with two boolean-mask gathers and three index_add / index_put scatters;
//  * project_lookup_kernel: the pixel lookup of collab_splats/utils/utils.py:13-40 (project_gaussians): rounded,
//    clamped flat pixel index of every projected centre and the "radius > 1 px" visibility mask, on the device.
// Both are HBM streams: 16 B (+8 B radii) per (camera, Gaussian) in, 8-12 B per Gaussian out.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
densify_stats_kernel(const float2* __restrict__ grads, const int2* __restrict__ radii, int C, int N, float sx, float sy,
                     float inv_extent, float* __restrict__ grad2d, float* __restrict__ count,
                     float* __restrict__ radii_max) {
  const int n = blockIdx.x * 256 + threadIdx.x;
  if (n >= N) return;
  float g = 0.f, cnt = 0.f, rmax = 0.f;
  for (int c = 0; c < C; ++c) {
    const size_t e = (size_t)c * N + n;
    const int2 r = __ldg(radii + e);
    if (r.x > 0 && r.y > 0) {
      const float2 v = __ldg(grads + e);
      const float gx = v.x * sx, gy = v.y * sy;
      g += sqrtf(gx * gx + gy * gy);
      cnt += 1.f;
      rmax = fmaxf(rmax, (float)max(r.x, r.y) * inv_extent);
    }
  }
  if (cnt > 0.f) {
    grad2d[n] += g;
    count[n] += cnt;
    if (radii_max) radii_max[n] = fmaxf(radii_max[n], rmax);
  }
}

__global__ void __launch_bounds__(256)
project_lookup_kernel(const float2* __restrict__ means2d, const int2* __restrict__ radii, int N, int W, int H,
                      long long* __restrict__ flat, unsigned char* __restrict__ valid) {
  const int n = blockIdx.x * 256 + threadIdx.x;
  if (n >= N) return;
  const float2 m = __ldg(means2d + n);
  const int2 r = __ldg(radii + n);
  // torch.round (half to even) -> long -> clamp; the float -> integer conversion saturates instead of wrapping
  const long long x = min(max((long long)rintf(m.x), 0ll), (long long)W - 1);
  const long long y = min(max((long long)rintf(m.y), 0ll), (long long)H - 1);
  flat[n] = x + y * (long long)W;
  valid[n] = (r.x > 1 || r.y > 1) ? 1 : 0;
}

}  // namespace

// grads [C,N,2] (means2d.grad or .absgrad), radii [C,N,2] i32; grad2d / count / radii_max [N] are ACCUMULATED
// (radii_max may be NULL).  sx = width/2 * n_cameras, sy = height/2 * n_cameras, inv_extent = 1/max(width,height).
extern "C" int rs_densify_stats(const float* grads, const int32_t* radii, int C, int N, float sx, float sy,
                                float inv_extent, float* grad2d, float* count, float* radii_max, void* stream) {
  RsSpan span__("rs_densify_stats", stream);
  if (C < 0 || N < 0) return RS_ERR_BAD_ARG;
  if (C == 0 || N == 0) return RS_OK;
  if (!grads || !radii || !grad2d || !count) return RS_ERR_BAD_ARG;
  densify_stats_kernel<<<rs_div_up(N, 256), 256, 0, (cudaStream_t)stream>>>(
      (const float2*)grads, (const int2*)radii, C, N, sx, sy, inv_extent, grad2d, count, radii_max);
  RS_RETURN_LAST_ERROR();
}

// means2d [N,2], radii [N,2] i32 of ONE camera -> flat pixel index [N] i64, visibility mask [N] u8
extern "C" int rs_project_lookup(const float* means2d, const int32_t* radii, int N, int width, int height,
                                 long long* proj_flattened, unsigned char* valid_mask, void* stream) {
  RsSpan span__("rs_project_lookup", stream);
  if (N < 0 || width <= 0 || height <= 0) return RS_ERR_BAD_ARG;
  if (N == 0) return RS_OK;
  if (!means2d || !radii || !proj_flattened || !valid_mask) return RS_ERR_BAD_ARG;
  project_lookup_kernel<<<rs_div_up(N, 256), 256, 0, (cudaStream_t)stream>>>(
      (const float2*)means2d, (const int2*)radii, N, width, height, proj_flattened, valid_mask);
  RS_RETURN_LAST_ERROR();
}
