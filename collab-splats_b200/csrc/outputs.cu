// get_outputs epilogue of the rade-gs model: everything RadegsModel.get_outputs does to the
// rasterizer's images after the call (collab_splats/models/rade_gs_model.py:200-271), for one camera:
//   normal_error_map[k] = 1 - <expected_normals, depth_double_to_normal(camera, expected, median)[k]>     :202-214
//   normals = (expected_normals + 1) / 2                                                                  :221
//   rgb = clamp(render[..., :3] + (1 - alpha) * background, 0, 1)                                         :227-229
//   depth_im / depth / median_depth / normals = where(alpha > 0, x, x.detach().max())                     :237-254
// The reference runs this as ~25 elementwise / reduction torch kernels plus the host-side camera rebuild of
// depth_double_to_normal; here it is two launches forward (the four global maxima, then one pass over the pixels)
// and one backward.  HBM-bound: ~100 B per pixel forward.
#include "common.cuh"
#include "dn_stencil.cuh"

namespace {
using rs::V3;
using rs::dn_term;

constexpr int OB = 256;

// monotone float <-> unsigned key, so that the float maximum is an unsigned atomicMax
__device__ __forceinline__ unsigned f2key(float f) {
  const unsigned b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

struct OutArgs {
  const float* render;      // [H,W,D]
  const float* alphas;      // [H,W]
  const float* exp_depth;   // [H,W]
  const float* med_depth;   // [H,W]
  const float* normals;     // [H,W,3]
  const float* background;  // [3]
  float fx, fy;
  int W, H, D, depth_ch, use_dn;   // depth_ch: channel of render holding the "ED" depth image, -1 if none
  unsigned* maxima;         // [4] keys: depth_im, expected depth, median depth, (normals + 1) / 2
  // forward outputs
  float* rgb;               // [H,W,3]
  float* depth;             // [H,W]
  float* median;            // [H,W]
  float* depth_im;          // [H,W] or null
  float* normals_out;       // [H,W,3]
  float* err;               // [2,H,W]
  // backward: upstream gradients of the outputs (null = zero) and gradients of the inputs
  const float *v_rgb, *v_depth, *v_median, *v_depth_im, *v_normals_out, *v_err, *v_accum;
  float *g_render, *g_alphas, *g_exp, *g_med, *g_normals;   // g_exp / g_med are zero-filled by the caller (atomics)
};

__global__ void __launch_bounds__(OB) outputs_max_kernel(const OutArgs a) {
  __shared__ unsigned s_red[4][OB / 32];
  const long long P = (long long)a.W * a.H;
  unsigned m[4] = {0u, 0u, 0u, 0u};   // key 0 is below every float's key
  for (long long p = (long long)blockIdx.x * OB + threadIdx.x; p < P; p += (long long)gridDim.x * OB) {
    if (a.depth_ch >= 0) m[0] = max(m[0], f2key(__ldg(a.render + p * a.D + a.depth_ch)));
    m[1] = max(m[1], f2key(__ldg(a.exp_depth + p)));
    m[2] = max(m[2], f2key(__ldg(a.med_depth + p)));
#pragma unroll
    for (int k = 0; k < 3; ++k) m[3] = max(m[3], f2key((__ldg(a.normals + p * 3 + k) + 1.f) / 2.f));
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) m[i] = max(m[i], __shfl_xor_sync(RS_FULL_MASK, m[i], d));
    if (lane == 0) s_red[i][warp] = m[i];
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    unsigned v = 0u;
#pragma unroll
    for (int w = 0; w < OB / 32; ++w) v = max(v, s_red[threadIdx.x][w]);
    atomicMax(a.maxima + threadIdx.x, v);
  }
}

__global__ void __launch_bounds__(OB) outputs_fwd_kernel(const OutArgs a) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= a.W || y >= a.H) return;
  const size_t p = (size_t)y * a.W + x;
  const size_t P = (size_t)a.W * a.H;
  const float alpha = __ldg(a.alphas + p);
  const bool hit = alpha > 0.f;
  const float* rc = a.render + p * a.D;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    // (separately rounded product and sum, as the reference's two torch kernels: no FMA contraction)
    const float raw = __fadd_rn(__ldg(rc + k), __fmul_rn(1.f - alpha, __ldg(a.background + k)));
    a.rgb[p * 3 + k] = fminf(fmaxf(raw, 0.f), 1.f);
  }
  if (a.depth_im) a.depth_im[p] = hit ? __ldg(rc + a.depth_ch) : key2f(a.maxima[0]);
  a.depth[p] = hit ? __ldg(a.exp_depth + p) : key2f(a.maxima[1]);
  a.median[p] = hit ? __ldg(a.med_depth + p) : key2f(a.maxima[2]);
  const V3 N = {__ldg(a.normals + p * 3), __ldg(a.normals + p * 3 + 1), __ldg(a.normals + p * 3 + 2)};
  const float nmax = key2f(a.maxima[3]);
  a.normals_out[p * 3] = hit ? (N.x + 1.f) / 2.f : nmax;
  a.normals_out[p * 3 + 1] = hit ? (N.y + 1.f) / 2.f : nmax;
  a.normals_out[p * 3 + 2] = hit ? (N.z + 1.f) / 2.f : nmax;
  float e0 = 0.f, e1 = 0.f;
  if (a.use_dn) {
    V3 unused = {0.f, 0.f, 0.f};
    const float ifx = 1.0f / a.fx, ify = 1.0f / a.fy;
    e0 = dn_term(a.exp_depth, nullptr, x, y, a.W, a.H, ifx, ify, N, 0.f, unused);
    e1 = dn_term(a.med_depth, nullptr, x, y, a.W, a.H, ifx, ify, N, 0.f, unused);
  }
  a.err[p] = e0;
  a.err[P + p] = e1;
}

__global__ void __launch_bounds__(OB) outputs_bwd_kernel(const OutArgs a) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= a.W || y >= a.H) return;
  const size_t p = (size_t)y * a.W + x;
  const size_t P = (size_t)a.W * a.H;
  const float alpha = __ldg(a.alphas + p);
  const bool hit = alpha > 0.f;
  const float* rc = a.render + p * a.D;
  float* gr = a.g_render + p * a.D;
  float g_alpha = a.v_accum ? __ldg(a.v_accum + p) : 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float bg = __ldg(a.background + k);
    const float raw = __fadd_rn(__ldg(rc + k), __fmul_rn(1.f - alpha, bg));
    float g = a.v_rgb ? __ldg(a.v_rgb + p * 3 + k) : 0.f;
    if (!(raw >= 0.f && raw <= 1.f)) g = 0.f;   // clamp passes the gradient only inside [0,1]
    gr[k] = g;
    g_alpha -= g * bg;
  }
  for (int k = 3; k < a.D; ++k)
    gr[k] = (k == a.depth_ch && hit && a.v_depth_im) ? __ldg(a.v_depth_im + p) : 0.f;
  a.g_alphas[p] = g_alpha;   // where(alpha > 0, ...) has no gradient with respect to alpha
  if (hit && a.v_depth) atomicAdd(a.g_exp + p, __ldg(a.v_depth + p));
  if (hit && a.v_median) atomicAdd(a.g_med + p, __ldg(a.v_median + p));
  V3 vN = {0.f, 0.f, 0.f};
  if (hit && a.v_normals_out) {
    vN.x = 0.5f * __ldg(a.v_normals_out + p * 3);
    vN.y = 0.5f * __ldg(a.v_normals_out + p * 3 + 1);
    vN.z = 0.5f * __ldg(a.v_normals_out + p * 3 + 2);
  }
  if (a.use_dn && a.v_err) {
    const V3 N = {__ldg(a.normals + p * 3), __ldg(a.normals + p * 3 + 1), __ldg(a.normals + p * 3 + 2)};
    const float ifx = 1.0f / a.fx, ify = 1.0f / a.fy;
    dn_term(a.exp_depth, a.g_exp, x, y, a.W, a.H, ifx, ify, N, __ldg(a.v_err + p), vN);
    dn_term(a.med_depth, a.g_med, x, y, a.W, a.H, ifx, ify, N, __ldg(a.v_err + P + p), vN);
  }
  a.g_normals[p * 3] = vN.x; a.g_normals[p * 3 + 1] = vN.y; a.g_normals[p * 3 + 2] = vN.z;
}

}  // namespace

// One camera.  `maxima` (4 x u32, device) must be zero-filled by the caller; it receives the four masked-fill values
// (as order-preserving keys) and is read again by the same call's second launch.
extern "C" int rs_rade_outputs_fwd(const float* render, const float* alphas, const float* exp_depth,
                                   const float* med_depth, const float* normals, const float* background, float fx,
                                   float fy, int width, int height, int D, int depth_channel, int use_depth_normal,
                                   uint32_t* maxima, float* rgb, float* depth, float* median_depth, float* depth_im,
                                   float* normals_out, float* error_maps, void* stream) {
  RsSpan span__("rs_rade_outputs_fwd", stream);
  if (width <= 0 || height <= 0 || D < 3 || depth_channel >= D) return RS_ERR_BAD_ARG;
  if (!render || !alphas || !exp_depth || !med_depth || !normals || !background || !maxima || !rgb || !depth ||
      !median_depth || !normals_out || !error_maps || (depth_channel >= 0 && !depth_im))
    return RS_ERR_BAD_ARG;
  OutArgs a{};
  a.render = render; a.alphas = alphas; a.exp_depth = exp_depth; a.med_depth = med_depth; a.normals = normals;
  a.background = background; a.fx = fx; a.fy = fy; a.W = width; a.H = height; a.D = D;
  a.depth_ch = depth_channel; a.use_dn = use_depth_normal; a.maxima = maxima;
  a.rgb = rgb; a.depth = depth; a.median = median_depth; a.depth_im = depth_channel >= 0 ? depth_im : nullptr;
  a.normals_out = normals_out; a.err = error_maps;
  const long long P = (long long)width * height;
  const int blocks = (int)(P / OB < 1 ? 1 : (P / OB > 148 * 8 ? 148 * 8 : P / OB));
  outputs_max_kernel<<<blocks, OB, 0, (cudaStream_t)stream>>>(a);
  dim3 grid(rs_div_up(width, 32), rs_div_up(height, 8));
  outputs_fwd_kernel<<<grid, OB, 0, (cudaStream_t)stream>>>(a);
  rs_count_launches(1);
  RS_RETURN_LAST_ERROR();
}

// Gradients of the inputs for upstream gradients of the outputs (any of them may be NULL = zero); g_exp_depth and
// g_med_depth must be zero-filled by the caller (the depth -> normal stencil scatters into them).
extern "C" int rs_rade_outputs_bwd(const float* render, const float* alphas, const float* exp_depth,
                                   const float* med_depth, const float* normals, const float* background, float fx,
                                   float fy, int width, int height, int D, int depth_channel, int use_depth_normal,
                                   const float* v_rgb, const float* v_depth, const float* v_median_depth,
                                   const float* v_depth_im, const float* v_normals_out, const float* v_error_maps,
                                   const float* v_accumulation, float* g_render, float* g_alphas, float* g_exp_depth,
                                   float* g_med_depth, float* g_normals, void* stream) {
  RsSpan span__("rs_rade_outputs_bwd", stream);
  if (width <= 0 || height <= 0 || D < 3 || depth_channel >= D) return RS_ERR_BAD_ARG;
  if (!render || !alphas || !exp_depth || !med_depth || !normals || !background || !g_render || !g_alphas ||
      !g_exp_depth || !g_med_depth || !g_normals)
    return RS_ERR_BAD_ARG;
  OutArgs a{};
  a.render = render; a.alphas = alphas; a.exp_depth = exp_depth; a.med_depth = med_depth; a.normals = normals;
  a.background = background; a.fx = fx; a.fy = fy; a.W = width; a.H = height; a.D = D;
  a.depth_ch = depth_channel; a.use_dn = use_depth_normal;
  a.v_rgb = v_rgb; a.v_depth = v_depth; a.v_median = v_median_depth; a.v_depth_im = v_depth_im;
  a.v_normals_out = v_normals_out; a.v_err = v_error_maps; a.v_accum = v_accumulation;
  a.g_render = g_render; a.g_alphas = g_alphas; a.g_exp = g_exp_depth; a.g_med = g_med_depth; a.g_normals = g_normals;
  dim3 grid(rs_div_up(width, 32), rs_div_up(height, 8));
  outputs_bwd_kernel<<<grid, OB, 0, (cudaStream_t)stream>>>(a);
  RS_RETURN_LAST_ERROR();
}
