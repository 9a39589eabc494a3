// Feature decode + cosine loss of the rade-features model, forward and backward.
// Replaces, as run every training step after rasterization():
//   collab_splats/models/rade_features_model.py:149-189  decode_features: permute, F.interpolate(bilinear,
//       align_corners=False) of the rendered [H,W,F] feature map to the main branch's feature-map size, TwoLayerMLP,
//       F.interpolate of every other branch's output to its own size;
//   collab_splats/utils/features.py:408-449             TwoLayerMLP: 1x1 conv F->Hd, ReLU, one 1x1 conv Hd->C_b per branch;
//   collab_splats/models/rade_features_model.py:564-582  sum_b w_b * mean_q(1 - cosine_similarity(pred_b, gt_b, dim=0)) * lambda.
// The reference does this with ~25 framework kernels forward (+ autograd) over a permuted 4 F H W byte copy of the
// render; here a handful of kernels read the render in place (channel offset inside the [H,W,ld] rows) and write gradients
// straight into the render's gradient image:
//   feat_hidden_kernel      x[Pm,F] = bilinear(render), h[Pm,Hd] = relu(W1 x + b1)
//   feat_branch_fwd_kernel  per branch: hq = bilinear(h) (the 1x1 conv commutes with the resize: both are linear and the
//                           taps sum to 1, so resize(conv(h)) = conv(resize(h))), y = W2 hq + b2, optional decoded
//                           output, per-pixel <y,y>, <y,g>, <g,g>
//   feat_cosine_kernel      per pixel: loss term and the two coefficients of dL/dy
//   feat_branch_bwd_kernel  y recomputed -> dy -> v_W2, v_b2, v_h (scattered through the taps)
//   feat_hidden_bwd_kernel  v_h -> relu mask -> v_W1, v_b1, v_x -> scattered into v_render through the taps
// Tiles: 32 pixels x 64 output channels per step, operands in shared memory, 8 pixels per thread in registers;
// plain FP32 FMA (the whole step is ~1.5 GFLOP at the reference's sizes: launch- and latency-bound, not a
// tensor-core problem).  Sums across CTAs are fp32 atomics (order-dependent in the last bits).
#include "common.cuh"

namespace {

constexpr int FD_TP = 32;       // pixels per CTA
constexpr int FD_CC = 64;       // output channels per step
constexpr int FD_MAXD = 128;    // max F and max Hd
constexpr int FD_PPAD = 36;     // row stride of the [k][pixel] tiles (float4-aligned, spreads banks)
constexpr float FD_COS_EPS = 1e-8f;   // torch.nn.functional.cosine_similarity default eps

struct Tap2 {
  int idx[4];
  float w[4];
};

// torch's bilinear source index (align_corners=False): src = in/out * (dst + 0.5) - 0.5 clamped at 0; i1 = i0 + 1
// unless i0 is the last row/column
__device__ __forceinline__ void tap1(int dst, int in_size, int out_size, int& i0, int& i1, float& l0, float& l1) {
  const float scale = (float)in_size / (float)out_size;
  float src = scale * ((float)dst + 0.5f) - 0.5f;
  src = src < 0.f ? 0.f : src;
  i0 = min((int)src, in_size - 1);
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = src - (float)i0;
  l0 = 1.f - l1;
}

__device__ __forceinline__ Tap2 make_tap(long long q, int Hin, int Win, int Hout, int Wout) {
  Tap2 t;
  const int oy = (int)(q / Wout), ox = (int)(q % Wout);
  int y0, y1, x0, x1;
  float ly0, ly1, lx0, lx1;
  tap1(oy, Hin, Hout, y0, y1, ly0, ly1);
  tap1(ox, Win, Wout, x0, x1, lx0, lx1);
  t.idx[0] = y0 * Win + x0; t.w[0] = ly0 * lx0;
  t.idx[1] = y0 * Win + x1; t.w[1] = ly0 * lx1;
  t.idx[2] = y1 * Win + x0; t.w[2] = ly1 * lx0;
  t.idx[3] = y1 * Win + x1; t.w[3] = ly1 * lx1;
  return t;
}

// ---------------------------------------------------------------------------------------------- hidden layer, forward
// smem: xs[FD_TP][F] | W1s[Hd][F+1]
__global__ void __launch_bounds__(256)
feat_hidden_kernel(const float* __restrict__ render, int H, int W, int ld, int ch0, int F, int Hm, int Wm,
                   const float* __restrict__ W1, const float* __restrict__ b1, int Hd, float* __restrict__ xg,
                   float* __restrict__ hg) {
  extern __shared__ float sm[];
  float* xs = sm;
  float* W1s = sm + FD_TP * F;
  const int t = threadIdx.x;
  const long long Pm = (long long)Hm * Wm, p0 = (long long)blockIdx.x * FD_TP;
  for (int e = t; e < Hd * F; e += 256) W1s[(e / F) * (F + 1) + e % F] = __ldg(W1 + e);
  for (int e = t; e < FD_TP * F; e += 256) {
    const int p = e / F, f = e % F;
    float v = 0.f;
    if (p0 + p < Pm) {
      const Tap2 tp = make_tap(p0 + p, H, W, Hm, Wm);
#pragma unroll
      for (int k = 0; k < 4; ++k) v += tp.w[k] * __ldg(render + (size_t)tp.idx[k] * ld + ch0 + f);
      xg[(size_t)(p0 + p) * F + f] = v;
    }
    xs[e] = v;
  }
  __syncthreads();
  const int pg = t >> 6;
  for (int j = t & 63; j < Hd; j += 64) {
    float acc[8];
    const float b = __ldg(b1 + j);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = b;
    const float* wr = W1s + j * (F + 1);
    for (int f = 0; f < F; ++f) {
      const float w = wr[f];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(w, xs[(pg * 8 + i) * F + f], acc[i]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (p0 + pg * 8 + i < Pm) hg[(size_t)(p0 + pg * 8 + i) * Hd + j] = fmaxf(acc[i], 0.f);
  }
}

// ---------------------------------------------------------------------------------------------- one branch
struct BranchArgs {
  const float* h;        // [Hm*Wm, Hd]
  int Hm, Wm, Hd;
  const float* W2;       // [C, Hd]
  const float* b2;       // [C]
  int C;
  int c_per_cta;         // channels handled by one CTA of the y dimension of the grid (multiple of FD_CC)
  const float* gt;       // [C, Hb*Wb] or NULL (decode only)
  int Hb, Wb;
  float scale;           // branch weight * loss lambda / (Hb*Wb)
  float* loss;           // [1], accumulated
  float* psum;           // [Hb*Wb, 3] scratch: <y,y>, <y,g>, <g,g>, then (ca, cb, -)
  float* v_h;            // [Hm*Wm, Hd], accumulated (NULL: forward only)
  float* v_W2;           // [C, Hd], accumulated
  float* v_b2;           // [C], accumulated
  float* decoded;        // [C, Hb*Wb] or NULL
};

// y[8 pixels of this thread's group][channel c0 + (t & 63)] = b2 + W2 hq
__device__ __forceinline__ void branch_tile_gemm(const float* W2s, const float* hqT, int Hd, int cl, int pg, float bias,
                                                 float (&acc)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = bias;
  const float* wr = W2s + cl * (Hd + 1);
  for (int k = 0; k < Hd; ++k) {
    const float w = wr[k];
    const float4 a = *reinterpret_cast<const float4*>(hqT + k * FD_PPAD + pg * 8);
    const float4 b = *reinterpret_cast<const float4*>(hqT + k * FD_PPAD + pg * 8 + 4);
    acc[0] = fmaf(w, a.x, acc[0]); acc[1] = fmaf(w, a.y, acc[1]); acc[2] = fmaf(w, a.z, acc[2]);
    acc[3] = fmaf(w, a.w, acc[3]); acc[4] = fmaf(w, b.x, acc[4]); acc[5] = fmaf(w, b.y, acc[5]);
    acc[6] = fmaf(w, b.z, acc[6]); acc[7] = fmaf(w, b.w, acc[7]);
  }
}

// taps of the CTA's 32 pixels + hqT[k][p] = bilinear(h)
__device__ __forceinline__ void branch_stage_pixels(const BranchArgs& A, long long q0, long long Pb, float* hqT,
                                                    float* tapw, int* tapi) {
  const int t = threadIdx.x, Hd = A.Hd;
  if (t < FD_TP) {
    Tap2 tp;
    if (q0 + t < Pb) tp = make_tap(q0 + t, A.Hm, A.Wm, A.Hb, A.Wb);
    else
      for (int k = 0; k < 4; ++k) { tp.idx[k] = 0; tp.w[k] = 0.f; }
    for (int k = 0; k < 4; ++k) { tapi[t * 4 + k] = tp.idx[k]; tapw[t * 4 + k] = tp.w[k]; }
  }
  __syncthreads();
  for (int e = t; e < FD_TP * Hd; e += 256) {
    const int p = e / Hd, k = e % Hd;
    float v = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float w = tapw[p * 4 + j];
      if (w != 0.f) v = fmaf(w, __ldg(A.h + (size_t)tapi[p * 4 + j] * Hd + k), v);
    }
    hqT[k * FD_PPAD + p] = v;
  }
}

__device__ __forceinline__ void branch_load_w2(const BranchArgs& A, int c0, float* W2s) {
  const int Hd = A.Hd;
  for (int e = threadIdx.x; e < FD_CC * Hd; e += 256) {
    const int r = e / Hd, k = e % Hd;
    W2s[r * (Hd + 1) + k] = (c0 + r < A.C) ? __ldg(A.W2 + (size_t)(c0 + r) * Hd + k) : 0.f;
  }
}

// grid (pixel tiles, channel splits).  smem: hqT[Hd][FD_PPAD] | W2s[FD_CC][Hd+1] | tapw[FD_TP][4] | tapi[FD_TP][4]
__global__ void __launch_bounds__(256)
feat_branch_fwd_kernel(BranchArgs A) {
  extern __shared__ float sm[];
  const int Hd = A.Hd;
  float* hqT = sm;
  float* W2s = hqT + Hd * FD_PPAD;
  float* tapw = W2s + FD_CC * (Hd + 1);
  int* tapi = reinterpret_cast<int*>(tapw + FD_TP * 4);
  const int t = threadIdx.x, cl = t & 63, pg = t >> 6, lane = t & 31;
  const long long Pb = (long long)A.Hb * A.Wb, q0 = (long long)blockIdx.x * FD_TP;
  const int c_begin = blockIdx.y * A.c_per_cta, c_end = min(A.C, c_begin + A.c_per_cta);
  branch_stage_pixels(A, q0, Pb, hqT, tapw, tapi);
  float syy[8], syg[8], sgg[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) syy[i] = syg[i] = sgg[i] = 0.f;
  for (int c0 = c_begin; c0 < c_end; c0 += FD_CC) {
    __syncthreads();
    branch_load_w2(A, c0, W2s);
    __syncthreads();
    const int c = c0 + cl;
    if (c < c_end) {
      float acc[8];
      branch_tile_gemm(W2s, hqT, Hd, cl, pg, __ldg(A.b2 + c), acc);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const long long q = q0 + pg * 8 + i;
        if (q < Pb) {
          if (A.decoded) A.decoded[(size_t)c * Pb + q] = acc[i];
          if (A.gt) {
            const float g = __ldg(A.gt + (size_t)c * Pb + q);
            syy[i] = fmaf(acc[i], acc[i], syy[i]); syg[i] = fmaf(acc[i], g, syg[i]); sgg[i] = fmaf(g, g, sgg[i]);
          }
        }
      }
    }
  }
  if (!A.gt) return;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
      syy[i] += __shfl_xor_sync(RS_FULL_MASK, syy[i], d);
      syg[i] += __shfl_xor_sync(RS_FULL_MASK, syg[i], d);
      sgg[i] += __shfl_xor_sync(RS_FULL_MASK, sgg[i], d);
    }
    const long long q = q0 + pg * 8 + i;
    if (lane == 0 && q < Pb) {
      atomicAdd(A.psum + q * 3, syy[i]);
      atomicAdd(A.psum + q * 3 + 1, syg[i]);
      atomicAdd(A.psum + q * 3 + 2, sgg[i]);
    }
  }
}

// per pixel: the three sums -> loss term and the two coefficients of dL/dy_c = ca * g_c + cb * y_c (in place)
__global__ void __launch_bounds__(256)
feat_cosine_kernel(float* __restrict__ psum, long long Pb, float scale, float* __restrict__ loss) {
  const long long q = (long long)blockIdx.x * 256 + threadIdx.x;
  float l = 0.f;
  if (q < Pb) {
    const float ny = sqrtf(psum[q * 3]), ng = sqrtf(psum[q * 3 + 2]);
    const float nyc = fmaxf(ny, FD_COS_EPS), ngc = fmaxf(ng, FD_COS_EPS);
    const float cosv = psum[q * 3 + 1] / (nyc * ngc);
    l = scale * (1.f - cosv);
    psum[q * 3] = -scale / (nyc * ngc);
    psum[q * 3 + 1] = ny > FD_COS_EPS ? scale * cosv / (nyc * nyc) : 0.f;
  }
  __shared__ float part[8];
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) l += __shfl_xor_sync(RS_FULL_MASK, l, d);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = l;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += part[i];
    atomicAdd(loss, s);
  }
}

// grid (pixel tiles, channel splits).  smem: hqT | W2s | dys[FD_TP][FD_CC+1] | dW2s[FD_CC][Hd+1] | coef[FD_TP][2] |
// tapw | tapi
__global__ void __launch_bounds__(256)
feat_branch_bwd_kernel(BranchArgs A) {
  extern __shared__ float sm[];
  const int Hd = A.Hd;
  float* hqT = sm;
  float* W2s = hqT + Hd * FD_PPAD;
  float* dys = W2s + FD_CC * (Hd + 1);
  float* dW2s = dys + FD_TP * (FD_CC + 1);
  float* coef = dW2s + FD_CC * (Hd + 1);
  float* tapw = coef + FD_TP * 2;
  int* tapi = reinterpret_cast<int*>(tapw + FD_TP * 4);
  const int t = threadIdx.x, cl = t & 63, pg = t >> 6;
  const long long Pb = (long long)A.Hb * A.Wb, q0 = (long long)blockIdx.x * FD_TP;
  const int c_begin = blockIdx.y * A.c_per_cta, c_end = min(A.C, c_begin + A.c_per_cta);
  if (t < FD_TP) {
    const bool ok = q0 + t < Pb;
    coef[t * 2] = ok ? A.psum[(q0 + t) * 3] : 0.f;
    coef[t * 2 + 1] = ok ? A.psum[(q0 + t) * 3 + 1] : 0.f;
  }
  branch_stage_pixels(A, q0, Pb, hqT, tapw, tapi);
  float dh[2][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) dh[0][i] = dh[1][i] = 0.f;
  const int kper = (Hd + 3) / 4;       // v_W2 tile: this thread owns channel cl and hidden units pg*kper ..
  for (int c0 = c_begin; c0 < c_end; c0 += FD_CC) {
    __syncthreads();
    branch_load_w2(A, c0, W2s);
    __syncthreads();
    const int c = c0 + cl;
    float dy[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) dy[i] = 0.f;
    if (c < c_end) {
      float acc[8];
      branch_tile_gemm(W2s, hqT, Hd, cl, pg, __ldg(A.b2 + c), acc);
      float sb = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const long long q = q0 + pg * 8 + i;
        if (q < Pb) {
          const float g = __ldg(A.gt + (size_t)c * Pb + q);
          dy[i] = fmaf(coef[(pg * 8 + i) * 2], g, coef[(pg * 8 + i) * 2 + 1] * acc[i]);
          sb += dy[i];
        }
      }
      atomicAdd(A.v_b2 + c, sb);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) dys[(pg * 8 + i) * (FD_CC + 1) + cl] = dy[i];
    __syncthreads();
    {  // v_W2[c][k] partial over the CTA's 32 pixels: dy column in registers, hq rows broadcast
      float col[FD_TP];
#pragma unroll
      for (int p = 0; p < FD_TP; ++p) col[p] = dys[p * (FD_CC + 1) + cl];
      for (int k = pg * kper; k < min(Hd, (pg + 1) * kper); ++k) {
        float s = 0.f;
#pragma unroll
        for (int p4 = 0; p4 < FD_TP / 4; ++p4) {
          const float4 a = *reinterpret_cast<const float4*>(hqT + k * FD_PPAD + p4 * 4);
          s = fmaf(col[p4 * 4], a.x, s); s = fmaf(col[p4 * 4 + 1], a.y, s);
          s = fmaf(col[p4 * 4 + 2], a.z, s); s = fmaf(col[p4 * 4 + 3], a.w, s);
        }
        dW2s[cl * (Hd + 1) + k] = s;
      }
    }
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      const int k = kk * 64 + cl;
      if (k < Hd) {
        for (int r = 0; r < FD_CC; ++r) {
          const float w = W2s[r * (Hd + 1) + k];
#pragma unroll
          for (int i = 0; i < 8; ++i) dh[kk][i] = fmaf(w, dys[(pg * 8 + i) * (FD_CC + 1) + r], dh[kk][i]);
        }
      }
    }
    __syncthreads();
    for (int e = t; e < FD_CC * Hd; e += 256) {
      const int r = e / Hd, k = e % Hd;
      if (c0 + r < c_end) atomicAdd(A.v_W2 + (size_t)(c0 + r) * Hd + k, dW2s[r * (Hd + 1) + k]);
    }
  }
#pragma unroll
  for (int kk = 0; kk < 2; ++kk) {
    const int k = kk * 64 + cl;
    if (k < Hd) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int p = pg * 8 + i;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float w = tapw[p * 4 + j];
          if (w != 0.f) atomicAdd(A.v_h + (size_t)tapi[p * 4 + j] * Hd + k, w * dh[kk][i]);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- hidden layer, backward
// smem: xs[FD_TP][F] | W1s[Hd][F+1] | dhp[FD_TP][Hd+1]
__global__ void __launch_bounds__(256)
feat_hidden_bwd_kernel(const float* __restrict__ xg, const float* __restrict__ hg, const float* __restrict__ v_h,
                       int Hm, int Wm, int Hd, int F, const float* __restrict__ W1, float* __restrict__ v_W1,
                       float* __restrict__ v_b1, float* __restrict__ v_render, int H, int W, int ld, int ch0,
                       const float* __restrict__ gscale) {
  extern __shared__ float sm[];
  float* xs = sm;
  float* W1s = xs + FD_TP * F;
  float* dhp = W1s + Hd * (F + 1);
  const int t = threadIdx.x, pg = t >> 6;
  const long long Pm = (long long)Hm * Wm, p0 = (long long)blockIdx.x * FD_TP;
  const float gs = gscale ? __ldg(gscale) : 1.f;
  for (int e = t; e < Hd * F; e += 256) W1s[(e / F) * (F + 1) + e % F] = __ldg(W1 + e);
  for (int e = t; e < FD_TP * F; e += 256) xs[e] = (p0 + e / F < Pm) ? __ldg(xg + (size_t)p0 * F + e) : 0.f;
  for (int e = t; e < FD_TP * Hd; e += 256) {
    const int p = e / Hd, j = e % Hd;
    float v = 0.f;
    if (p0 + p < Pm && __ldg(hg + (size_t)p0 * Hd + e) > 0.f) v = __ldg(v_h + (size_t)p0 * Hd + e) * gs;
    dhp[p * (Hd + 1) + j] = v;
  }
  __syncthreads();
  if (t < Hd) {
    float s = 0.f;
    for (int p = 0; p < FD_TP; ++p) s += dhp[p * (Hd + 1) + t];
    atomicAdd(v_b1 + t, s);
  }
  for (int e = t; e < Hd * F; e += 256) {
    const int j = e / F, f = e % F;
    float s = 0.f;
#pragma unroll 8
    for (int p = 0; p < FD_TP; ++p) s = fmaf(dhp[p * (Hd + 1) + j], xs[p * F + f], s);
    atomicAdd(v_W1 + e, s);
  }
  if (!v_render) return;
  for (int f = t & 63; f < F; f += 64) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int j = 0; j < Hd; ++j) {
      const float w = W1s[j * (F + 1) + f];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(w, dhp[(pg * 8 + i) * (Hd + 1) + j], acc[i]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const long long p = p0 + pg * 8 + i;
      if (p < Pm) {
        const Tap2 tp = make_tap(p, H, W, Hm, Wm);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (tp.w[k] != 0.f) atomicAdd(v_render + (size_t)tp.idx[k] * ld + ch0 + f, tp.w[k] * acc[i]);
      }
    }
  }
}

template <typename K>
int opt_in_smem(K kernel, size_t bytes) {
  if (bytes > 227 * 1024) return RS_ERR_BAD_ARG;
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) { rs_set_last_cuda_error((int)e); return RS_ERR_LAUNCH; }
  }
  return RS_OK;
}

}  // namespace

// render [H,W,ld] (the rasterizer's colour output; feature channels are columns ch0 .. ch0+F-1) -> x [Hm*Wm, F]
// (the bilinearly resized feature map, kept for the backward) and h [Hm*Wm, Hd] = relu(W1 x + b1).  F, Hd <= 128.
extern "C" int rs_feature_hidden_fwd(const float* render, int H, int W, int ld, int ch0, int F, int Hm, int Wm,
                                     const float* W1, const float* b1, int Hd, float* x, float* h, void* stream) {
  RsSpan span__("rs_feature_hidden_fwd", stream);
  if (H <= 0 || W <= 0 || F <= 0 || F > FD_MAXD || Hd <= 0 || Hd > FD_MAXD || ch0 < 0 || ch0 + F > ld || Hm <= 0 ||
      Wm <= 0 || (long long)H * W > 0x7fffffffll || (long long)Hm * Wm > 0x7fffffffll)
    return RS_ERR_BAD_ARG;
  if (!render || !W1 || !b1 || !x || !h) return RS_ERR_BAD_ARG;
  const size_t smem = sizeof(float) * ((size_t)FD_TP * F + (size_t)Hd * (F + 1));
  int rc = opt_in_smem(feat_hidden_kernel, smem);
  if (rc != RS_OK) return rc;
  feat_hidden_kernel<<<rs_div_up((long long)Hm * Wm, FD_TP), 256, smem, (cudaStream_t)stream>>>(
      render, H, W, ld, ch0, F, Hm, Wm, W1, b1, Hd, x, h);
  RS_RETURN_LAST_ERROR();
}

// One decoder branch on the hidden map h [Hm*Wm, Hd]: y = W2 * bilinear(h -> Hb x Wb) + b2, [C, Hb*Wb] channel-first.
// decoded (optional) receives y.  With gt [C, Hb*Wb] (and psum, a zero-filled [Hb*Wb,3] scratch):
// *loss += scale * sum_q (1 - cos(y_q, gt_q)); with v_h as well (training) the call also accumulates v_W2 [C,Hd],
// v_b2 [C] and v_h [Hm*Wm,Hd] (zero-filled or carrying other branches' sums).
// scale = branch weight * features_loss_lambda / (Hb*Wb).  The grid splits pixels (32 per CTA) and, when the map is
// small, channels, so that at least two waves of 148 CTAs exist.
extern "C" int rs_feature_branch(const float* h, int Hm, int Wm, int Hd, const float* W2, const float* b2, int C,
                                 const float* gt, int Hb, int Wb, float scale, float* loss, float* psum, float* v_h,
                                 float* v_W2, float* v_b2, float* decoded, void* stream) {
  RsSpan span__("rs_feature_branch", stream);
  if (Hm <= 0 || Wm <= 0 || Hd <= 0 || Hd > FD_MAXD || C <= 0 || Hb <= 0 || Wb <= 0 ||
      (long long)Hm * Wm > 0x7fffffffll || (long long)Hb * Wb > 0x7fffffffll)
    return RS_ERR_BAD_ARG;
  if (!h || !W2 || !b2 || (!gt && !decoded) || (gt && (!loss || !psum)) || (v_h && (!gt || !v_W2 || !v_b2)))
    return RS_ERR_BAD_ARG;
  const long long Pb = (long long)Hb * Wb;
  const int tiles = rs_div_up(Pb, FD_TP), chunks = rs_div_up(C, FD_CC);
  const int splits = max(1, min(chunks, rs_div_up(2 * 148, tiles)));
  const int c_per_cta = rs_div_up(chunks, splits) * FD_CC;
  BranchArgs A{h, Hm, Wm, Hd, W2, b2, C, c_per_cta, gt, Hb, Wb, scale, loss, psum, v_h, v_W2, v_b2, decoded};
  const dim3 grid(tiles, rs_div_up(C, c_per_cta));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem_f = sizeof(float) * ((size_t)Hd * FD_PPAD + (size_t)FD_CC * (Hd + 1) + FD_TP * 8);
  int rc = opt_in_smem(feat_branch_fwd_kernel, smem_f);
  if (rc != RS_OK) return rc;
  feat_branch_fwd_kernel<<<grid, 256, smem_f, st>>>(A);
  if (gt) {
    feat_cosine_kernel<<<rs_div_up(Pb, 256), 256, 0, st>>>(psum, Pb, scale, loss);
    rs_count_launches(1);
    if (v_h) {
      const size_t smem_b = smem_f + sizeof(float) * ((size_t)FD_TP * (FD_CC + 1) + (size_t)FD_CC * (Hd + 1) + FD_TP * 2);
      rc = opt_in_smem(feat_branch_bwd_kernel, smem_b);
      if (rc != RS_OK) return rc;
      feat_branch_bwd_kernel<<<grid, 256, smem_b, st>>>(A);
      rs_count_launches(1);
    }
  }
  RS_RETURN_LAST_ERROR();
}

// v_h [Hm*Wm,Hd] -> v_W1 [Hd,F], v_b1 [Hd] (accumulated) and, through the relu mask, W1^T and the bilinear taps,
// v_render [H,W,ld] columns ch0.. (accumulated; NULL skips it).  gscale (device scalar or NULL = 1) multiplies v_h on
// load: the upstream gradient of the loss, applied without a pass over the image.
extern "C" int rs_feature_hidden_bwd(const float* x, const float* h, const float* v_h, int Hm, int Wm, int Hd, int F,
                                     const float* W1, float* v_W1, float* v_b1, float* v_render, int H, int W, int ld,
                                     int ch0, const float* gscale, void* stream) {
  RsSpan span__("rs_feature_hidden_bwd", stream);
  if (H <= 0 || W <= 0 || F <= 0 || F > FD_MAXD || Hd <= 0 || Hd > FD_MAXD || ch0 < 0 || ch0 + F > ld || Hm <= 0 ||
      Wm <= 0 || (long long)H * W > 0x7fffffffll || (long long)Hm * Wm > 0x7fffffffll)
    return RS_ERR_BAD_ARG;
  if (!x || !h || !v_h || !W1 || !v_W1 || !v_b1) return RS_ERR_BAD_ARG;
  const size_t smem = sizeof(float) * ((size_t)FD_TP * F + (size_t)Hd * (F + 1) + (size_t)FD_TP * (Hd + 1));
  int rc = opt_in_smem(feat_hidden_bwd_kernel, smem);
  if (rc != RS_OK) return rc;
  feat_hidden_bwd_kernel<<<rs_div_up((long long)Hm * Wm, FD_TP), 256, smem, (cudaStream_t)stream>>>(
      x, h, v_h, Hm, Wm, Hd, F, W1, v_W1, v_b1, v_render, H, W, ld, ch0, gscale);
  RS_RETURN_LAST_ERROR();
}
