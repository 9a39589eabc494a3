// Fused post-render loss: L1 on RGB (+ optional background blend and clamp) and the RaDe-GS depth-normal
// consistency loss, forward AND gradients in one pass over the images.
//
// This is the step right after the rasterizer in the reference's training loop:
//   depth -> points -> central differences -> cross -> normalise   collab_splats/utils/camera_utils.py:176-279
//   err = 1 - <rendered normal, depth normal>, two maps             collab_splats/models/rade_gs_model.py:202-219
//   loss = lambda * ((1-ratio) * mean(err_exp) + ratio * mean(err_med))   rade_gs_model.py:292-307
// which the reference runs as ~60 elementwise torch kernels plus two host-side camera rebuilds per step.
// Here one thread per pixel reads the 4-neighbourhood of both depth maps, evaluates both normals, accumulates
// the three loss sums (block reduction + one atomicAdd per block) and scatters the depth gradients with
// float atomics (8 per interior pixel).  HBM-bound: ~90 B per pixel.
#include "common.cuh"
#include "dn_stencil.cuh"

namespace {
using rs::V3;
using rs::dn_term;

constexpr int LB = 256;

struct LossArgs {
  const float* render;      // [H,W,D], first 3 channels are RGB
  const float* alphas;      // [H,W]
  const float* exp_depth;   // [H,W]
  const float* med_depth;   // [H,W]
  const float* normals;     // [H,W,3]
  const unsigned char* gt;  // [H,W,3] uint8
  const float* background;  // [3] or null
  float fx, fy;             // focal lengths (principal point is the image centre, rade_gs_model.py:327-334)
  int W, H, D;
  float w_l1, w_exp, w_med; // d(loss)/d(sum) weights: 1/(3P), lambda*(1-ratio)/P, lambda*ratio/P
  int use_dn;               // depth-normal term on/off (regularization_from_iter, rade_gs_model.py:202-205)
  float* sums;              // [4]: w_l1*sum|rgb-gt|, w_exp*sum err_exp, w_med*sum err_med, total loss
  float* v_render;          // [H,W,D]
  float* v_alphas;          // [H,W]
  float* v_exp_depth;       // [H,W]  (zero-initialised by the caller: receives atomics)
  float* v_med_depth;       // [H,W]  (zero-initialised by the caller)
  float* v_normals;         // [H,W,3]
};

__global__ void __launch_bounds__(LB) rade_loss_kernel(const LossArgs a) {
  __shared__ float s_red[4][LB / 32];
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  float l1 = 0.f, e0 = 0.f, e1 = 0.f;
  if (x < a.W && y < a.H) {
    const size_t p = (size_t)y * a.W + x;
    const float alpha = __ldg(a.alphas + p);
    float v_alpha = 0.f;
    const float* rc = a.render + p * a.D;
    float* vr = a.v_render + p * a.D;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float bg = a.background ? __ldg(a.background + k) : 0.f;
      const float raw = __ldg(rc + k) + (1.f - alpha) * bg;
      const float rgb = fminf(fmaxf(raw, 0.f), 1.f);
      const float diff = rgb - (float)__ldg(a.gt + p * 3 + k) * (1.0f / 255.0f);
      l1 += fabsf(diff);
      float g = (diff > 0.f ? a.w_l1 : (diff < 0.f ? -a.w_l1 : 0.f));
      if (!(raw >= 0.f && raw <= 1.f)) g = 0.f;  // clamp passes gradient only inside [0,1]
      vr[k] = g;
      v_alpha -= g * bg;
    }
    if (a.D <= 8)   // (wider rows -- rade-features -- are zero-filled by one memset before the launch)
      for (int k = 3; k < a.D; ++k) vr[k] = 0.f;
    a.v_alphas[p] = v_alpha;
    V3 vN = {0.f, 0.f, 0.f};
    if (a.use_dn) {
      const V3 N = {__ldg(a.normals + p * 3), __ldg(a.normals + p * 3 + 1), __ldg(a.normals + p * 3 + 2)};
      const float ifx = 1.0f / a.fx, ify = 1.0f / a.fy;
      e0 = dn_term(a.exp_depth, a.v_exp_depth, x, y, a.W, a.H, ifx, ify, N, a.w_exp, vN);
      e1 = dn_term(a.med_depth, a.v_med_depth, x, y, a.W, a.H, ifx, ify, N, a.w_med, vN);
    }
    a.v_normals[p * 3] = vN.x; a.v_normals[p * 3 + 1] = vN.y; a.v_normals[p * 3 + 2] = vN.z;
  }
  // block reduction of the (weighted) sums
  l1 *= a.w_l1; e0 *= a.w_exp; e1 *= a.w_med;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    l1 += __shfl_xor_sync(RS_FULL_MASK, l1, d);
    e0 += __shfl_xor_sync(RS_FULL_MASK, e0, d);
    e1 += __shfl_xor_sync(RS_FULL_MASK, e1, d);
  }
  if (lane == 0) { s_red[0][warp] = l1; s_red[1][warp] = e0; s_red[2][warp] = e1; s_red[3][warp] = l1 + e0 + e1; }
  __syncthreads();
  if (threadIdx.x < 4) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < LB / 32; ++w) s += s_red[threadIdx.x][w];
    atomicAdd(a.sums + threadIdx.x, s);
  }
}

}  // namespace

// One camera.  `sums` [4] and v_exp_depth / v_med_depth [H,W] must be zero-filled by the caller.
// On return sums = (w_l1*sum|rgb-gt|, w_exp*sum err_exp, w_med*sum err_med, loss = their total); the gradients
// written are d(loss)/d(input).
extern "C" int rs_rade_loss_fwd_bwd(const float* render, const float* alphas, const float* exp_depth,
                                    const float* med_depth, const float* normals, const unsigned char* gt_rgb_u8,
                                    const float* background, float fx, float fy, int width, int height, int D,
                                    float w_l1, float w_exp, float w_med, int use_depth_normal, float* sums,
                                    float* v_render, float* v_alphas, float* v_exp_depth, float* v_med_depth,
                                    float* v_normals, void* stream) {
  RsSpan span__("rs_rade_loss_fwd_bwd", stream);
  if (width <= 0 || height <= 0 || D < 3) return RS_ERR_BAD_ARG;
  if (!render || !alphas || !exp_depth || !med_depth || !normals || !gt_rgb_u8 || !sums || !v_render || !v_alphas ||
      !v_exp_depth || !v_med_depth || !v_normals)
    return RS_ERR_BAD_ARG;
  LossArgs a{render, alphas, exp_depth, med_depth, normals, gt_rgb_u8, background, fx, fy, width, height, D,
             w_l1, w_exp, w_med, use_depth_normal, sums, v_render, v_alphas, v_exp_depth, v_med_depth, v_normals};
  if (D > 8) {   // the gradient of the non-rgb channels is zero: 4 D bytes per pixel are cheaper to memset than to store
    cudaError_t e = cudaMemsetAsync(v_render, 0, sizeof(float) * (size_t)width * height * D, (cudaStream_t)stream);
    if (e != cudaSuccess) { rs_set_last_cuda_error((int)e); return RS_ERR_LAUNCH; }
  }
  dim3 grid(rs_div_up(width, 32), rs_div_up(height, 8));
  rade_loss_kernel<<<grid, LB, 0, (cudaStream_t)stream>>>(a);
  RS_RETURN_LAST_ERROR();
}
