// View-dependent colours for the rasterizer in one kernel: camera position from the view matrix, view
// direction, real SH (degree <= 3), +0.5, clamp at 0, and the optional depth channel of the "RGB+D/ED" render
// modes, written straight into the 4-channel padded colour rows the compositing kernel gathers.
// Replaces, inside gsplat `rasterization()` (reference call rade_gs_model.py:439-465 with
// sh_degree = 0..3), the chain  inverse(viewmats) -> dirs -> spherical_harmonics -> +0.5 -> clamp_min -> cat(depth)
// and its autograd (8 kernels forward, ~12 backward upstream) with one forward and one backward kernel.
//
// HBM-bound: 12*K B of coefficients per Gaussian dominate (192 B at K = 16).  Coefficient rows are staged
// through shared memory with coalesced 16-byte accesses (row stride padded to an odd word count, so the
// per-thread row walks are bank-conflict free); the backward sums over cameras in registers (no atomics).
#include "common.cuh"
#include "rade_math.cuh"

namespace {

constexpr int CB = 128;  // Gaussians per block

__device__ __forceinline__ void camera_position(const float* __restrict__ vm, float& px, float& py, float& pz) {
  // campos = -R^-1 t  (= inverse(viewmat)[:3,3]); R^-1 by cofactors so non-rigid view matrices work too
  const float a = __ldg(vm + 0), b = __ldg(vm + 1), c = __ldg(vm + 2), tx = __ldg(vm + 3);
  const float d = __ldg(vm + 4), e = __ldg(vm + 5), f = __ldg(vm + 6), ty = __ldg(vm + 7);
  const float g = __ldg(vm + 8), h = __ldg(vm + 9), i = __ldg(vm + 10), tz = __ldg(vm + 11);
  const float A = e * i - f * h, B = -(d * i - f * g), Cc = d * h - e * g;
  const float idet = 1.0f / (a * A + b * B + c * Cc);
  const float i00 = A * idet, i01 = (c * h - b * i) * idet, i02 = (b * f - c * e) * idet;
  const float i10 = B * idet, i11 = (a * i - c * g) * idet, i12 = (c * d - a * f) * idet;
  const float i20 = Cc * idet, i21 = (b * g - a * h) * idet, i22 = (a * e - b * d) * idet;
  px = -(i00 * tx + i01 * ty + i02 * tz);
  py = -(i10 * tx + i11 * ty + i12 * tz);
  pz = -(i20 * tx + i21 * ty + i22 * tz);
}

// cooperative, coalesced copy of `count` coefficient rows (row = K*3 floats) global <-> shared (stride RS).
// The flat index -> (row, column) split is a division by `row`: ROW > 0 makes it a compile-time constant
// (K = 16 -> 48 floats, the sh3 case that carries the bandwidth), ROW == 0 keeps the generic run-time form.
template <int ROW>
__device__ __forceinline__ void rows_to_smem_t(float* s, const float* __restrict__ src, int count, int row_rt, int RS_rt,
                                               int t) {
  const int row = ROW > 0 ? ROW : row_rt;
  const int RS = ROW > 0 ? (ROW | 1) : RS_rt;
  const int total = count * row;
  if ((total & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    for (int i = t; i < total / 4; i += CB) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
      const float vv[4] = {v.x, v.y, v.z, v.w};
      const int f0 = i * 4, r0 = f0 / row, c0 = f0 - r0 * row;   // one division per 16 bytes
      if constexpr (ROW > 0 && ROW % 4 == 0) {                   // a 16-byte piece never straddles two rows
        float* d = s + r0 * RS + c0;
        d[0] = vv[0]; d[1] = vv[1]; d[2] = vv[2]; d[3] = vv[3];
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          int r = r0, c = c0 + k;
          if (c >= row) { c -= row; ++r; }
          s[r * RS + c] = vv[k];
        }
      }
    }
  } else {
    for (int f = t; f < total; f += CB) s[(f / row) * RS + (f % row)] = __ldg(src + f);
  }
}
template <int ROW>
__device__ __forceinline__ void smem_to_rows_t(const float* s, float* __restrict__ dst, int count, int row_rt, int RS_rt,
                                               int t) {
  const int row = ROW > 0 ? ROW : row_rt;
  const int RS = ROW > 0 ? (ROW | 1) : RS_rt;
  const int total = count * row;
  if ((total & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    for (int i = t; i < total / 4; i += CB) {
      float vv[4];
      const int f0 = i * 4, r0 = f0 / row, c0 = f0 - r0 * row;
      if constexpr (ROW > 0 && ROW % 4 == 0) {
        const float* d = s + r0 * RS + c0;
        vv[0] = d[0]; vv[1] = d[1]; vv[2] = d[2]; vv[3] = d[3];
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          int r = r0, c = c0 + k;
          if (c >= row) { c -= row; ++r; }
          vv[k] = s[r * RS + c];
        }
      }
      reinterpret_cast<float4*>(dst)[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
    }
  } else {
    for (int f = t; f < total; f += CB) dst[f] = s[(f / row) * RS + (f % row)];
  }
}
__device__ __forceinline__ void rows_to_smem(float* s, const float* __restrict__ src, int count, int row, int RS, int t) {
  if (row == 48) rows_to_smem_t<48>(s, src, count, row, RS, t);
  else rows_to_smem_t<0>(s, src, count, row, RS, t);
}
__device__ __forceinline__ void smem_to_rows(const float* s, float* __restrict__ dst, int count, int row, int RS, int t) {
  if (row == 48) smem_to_rows_t<48>(s, dst, count, row, RS, t);
  else smem_to_rows_t<0>(s, dst, count, row, RS, t);
}

__global__ void __launch_bounds__(CB)
sh_colors_fwd_kernel(int degree, int K, int C, int N, const float* __restrict__ means,
                     const float* __restrict__ coeffs, const float* __restrict__ viewmats,
                     const int2* __restrict__ radii, const float* __restrict__ depths, float4* __restrict__ colors4) {
  extern __shared__ float s_rows[];
  const int t = threadIdx.x;
  const int n0 = blockIdx.x * CB;
  const int count = min(CB, N - n0);
  const int row = K * 3, RS = row | 1;
  rows_to_smem(s_rows, coeffs + (size_t)n0 * row, count, row, RS, t);
  __syncthreads();
  if (t >= count) return;
  const int n = n0 + t;
  const float mx = __ldg(means + n * 3), my = __ldg(means + n * 3 + 1), mz = __ldg(means + n * 3 + 2);
  const float* cf = s_rows + t * RS;
  const int nb = (degree + 1) * (degree + 1);
  for (int c = 0; c < C; ++c) {
    const size_t e = (size_t)c * N + n;
    float r = 0.f, g = 0.f, b = 0.f;
    const int2 rad = __ldg(radii + e);
    if (rad.x > 0 && rad.y > 0) {
      float cx, cy, cz;
      camera_position(viewmats + c * 16, cx, cy, cz);
      const float x = mx - cx, y = my - cy, z = mz - cz;
      const float inv = 1.f / fmaxf(sqrtf(x * x + y * y + z * z), 1e-12f);
      float basis[16];
      rs::sh_basis(degree, x * inv, y * inv, z * inv, basis);
#pragma unroll
      for (int q = 0; q < 16; ++q)
        if (q < nb) { r += basis[q] * cf[q * 3]; g += basis[q] * cf[q * 3 + 1]; b += basis[q] * cf[q * 3 + 2]; }
    }
    colors4[e] = make_float4(fmaxf(r + 0.5f, 0.f), fmaxf(g + 0.5f, 0.f), fmaxf(b + 0.5f, 0.f),
                             depths ? __ldg(depths + e) : 0.f);
  }
}

__global__ void __launch_bounds__(CB)
sh_colors_bwd_kernel(int degree, int K, int C, int N, const float* __restrict__ means,
                     const float* __restrict__ coeffs, const float* __restrict__ viewmats,
                     const int2* __restrict__ radii, const float4* __restrict__ v_colors4, int has_depth,
                     float* __restrict__ v_coeffs, float* __restrict__ v_means, float* __restrict__ v_depths) {
  extern __shared__ float s_rows[];
  const int t = threadIdx.x;
  const int n0 = blockIdx.x * CB;
  const int count = min(CB, N - n0);
  const int row = K * 3, RS = row | 1;
  rows_to_smem(s_rows, coeffs + (size_t)n0 * row, count, row, RS, t);
  __syncthreads();
  const int nb = (degree + 1) * (degree + 1);
  // Coefficient gradients accumulate in shared memory, not in 48 registers: the kernel is HBM-bound and needs
  // the occupancy.  With one camera they overwrite the coefficients in place (every read of a row entry
  // precedes its write); with several cameras they go to a second, zero-initialised row.
  const bool in_place = (C == 1);
  float* s_out = in_place ? s_rows : s_rows + CB * RS;
  const int n = n0 + t;
  if (t < count) {
    float* o = s_out + t * RS;
    if (!in_place)
      for (int i = 0; i < row; ++i) o[i] = 0.f;
    float vmx = 0.f, vmy = 0.f, vmz = 0.f;
    const float mx = __ldg(means + n * 3), my = __ldg(means + n * 3 + 1), mz = __ldg(means + n * 3 + 2);
    const float* cf = s_rows + t * RS;
    for (int c = 0; c < C; ++c) {
      const size_t e = (size_t)c * N + n;
      const float4 vc = __ldg(v_colors4 + e);
      if (has_depth) v_depths[e] = vc.w;
      const int2 rad = __ldg(radii + e);
      const bool live = rad.x > 0 && rad.y > 0;  // masked entries contribute exact zeros
      float cx, cy, cz;
      camera_position(viewmats + c * 16, cx, cy, cz);
      const float x = mx - cx, y = my - cy, z = mz - cz;
      const float inv = 1.f / fmaxf(sqrtf(x * x + y * y + z * z), 1e-12f);
      const float ux = x * inv, uy = y * inv, uz = z * inv;
      float basis[16], gq[16];
      rs::sh_basis(degree, ux, uy, uz, basis);
      float r = 0.f, g = 0.f, b = 0.f;
#pragma unroll
      for (int q = 0; q < 16; ++q)
        if (q < nb) { r += basis[q] * cf[q * 3]; g += basis[q] * cf[q * 3 + 1]; b += basis[q] * cf[q * 3 + 2]; }
      // clamp_min(sh + 0.5, 0): gradient passes where the un-clamped value is >= 0
      const float vr = (live && r + 0.5f >= 0.f) ? vc.x : 0.f, vg = (live && g + 0.5f >= 0.f) ? vc.y : 0.f,
                  vb = (live && b + 0.5f >= 0.f) ? vc.z : 0.f;
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        gq[q] = 0.f;
        if (q < K) {
          const float bq = q < nb ? basis[q] : 0.f;
          if (q < nb) gq[q] = vr * cf[q * 3] + vg * cf[q * 3 + 1] + vb * cf[q * 3 + 2];
          if (in_place) {
            o[q * 3] = bq * vr; o[q * 3 + 1] = bq * vg; o[q * 3 + 2] = bq * vb;
          } else {
            o[q * 3] += bq * vr; o[q * 3 + 1] += bq * vg; o[q * 3 + 2] += bq * vb;
          }
        }
      }
      float bx, by, bz;
      rs::sh_basis_vjp(degree, ux, uy, uz, gq, bx, by, bz);
      const float dd = ux * bx + uy * by + uz * bz;
      vmx += (bx - ux * dd) * inv; vmy += (by - uy * dd) * inv; vmz += (bz - uz * dd) * inv;
    }
    v_means[n * 3] = vmx; v_means[n * 3 + 1] = vmy; v_means[n * 3 + 2] = vmz;
  }
  __syncthreads();
  smem_to_rows(s_out, v_coeffs + (size_t)n0 * row, count, row, RS, t);
}


// ------------------------------------------------------------------------------------------------ camera-sharded split
// Multi-GPU (camera-sharded) form of the backward.  The SH-coefficient gradient of one camera is the outer product
// Y_k(dir(n, camera)) x v_rgb[n]: 48 floats per Gaussian that carry 3 floats of information.  So instead of
// all-reducing 192 B per Gaussian, every rank publishes its masked colour gradients (12 B per Gaussian and camera)
// in peer-visible memory and every rank rebuilds the SUM OVER ALL CAMERAS OF ALL RANKS of the outer products
// itself, reading the other ranks' rows straight over NVLink (or from an all-gathered copy):
//   sh_colors_bwd_local_kernel : v_colors4 -> vrgb[C,N,3] (clamp/visibility-masked rgb gradient),
//                                v_means (direction part), v_depths; also writes the camera positions to the header
//   sh_coeffs_gather_kernel    : v_coeffs[n,k,:] = sum over sources (rank g, camera c) of Y_k(dir(n, campos_gc)) *
//                                vrgb_gc[n,:], in a fixed (g, c) order, so every rank gets bit-identical sums.
// A source region is [header: RS_PEER_HEADER_BYTES, float4 campos per camera][float vrgb[cams][N][3]] -- 12 B per Gaussian
// and camera (round 2: was a padded float4; at 8 GPUs the pushes are NVLink-bandwidth-bound, a quarter fewer bytes).
constexpr int RS_PEER_HEADER_BYTES = 1024;  // up to 64 cameras per rank
constexpr int RS_MAX_PEERS = 16;
constexpr int RS_MAX_SOURCES = 64;   // cameras of all ranks in one step

struct PeerSources {
  const char* base[RS_MAX_PEERS];
  int cams[RS_MAX_PEERS];
  int n;
};

__global__ void __launch_bounds__(CB)
sh_colors_bwd_local_kernel(int degree, int K, int C, int N, const float* __restrict__ means,
                           const float* __restrict__ coeffs, const float* __restrict__ viewmats,
                           const int2* __restrict__ radii, const float4* __restrict__ v_colors4, int has_depth,
                           char* __restrict__ region, float* __restrict__ v_means, float* __restrict__ v_depths) {
  extern __shared__ float s_rows[];
  const int t = threadIdx.x;
  const int n0 = blockIdx.x * CB;
  const int count = min(CB, N - n0);
  const int row = K * 3, RS = row | 1;
  rows_to_smem(s_rows, coeffs + (size_t)n0 * row, count, row, RS, t);
  if (blockIdx.x == 0 && t < C) {
    float cx, cy, cz;
    camera_position(viewmats + t * 16, cx, cy, cz);
    reinterpret_cast<float4*>(region)[t] = make_float4(cx, cy, cz, 1.f);
  }
  __syncthreads();
  if (t >= count) return;
  const int nb = (degree + 1) * (degree + 1);
  const int n = n0 + t;
  float* vrgb = reinterpret_cast<float*>(region + RS_PEER_HEADER_BYTES);
  float vmx = 0.f, vmy = 0.f, vmz = 0.f;
  const float mx = __ldg(means + n * 3), my = __ldg(means + n * 3 + 1), mz = __ldg(means + n * 3 + 2);
  const float* cf = s_rows + t * RS;
  for (int c = 0; c < C; ++c) {
    const size_t e = (size_t)c * N + n;
    const float4 vc = __ldg(v_colors4 + e);
    if (has_depth) v_depths[e] = vc.w;
    const int2 rad = __ldg(radii + e);
    const bool live = rad.x > 0 && rad.y > 0;
    float cx, cy, cz;
    camera_position(viewmats + c * 16, cx, cy, cz);
    const float x = mx - cx, y = my - cy, z = mz - cz;
    const float inv = 1.f / fmaxf(sqrtf(x * x + y * y + z * z), 1e-12f);
    const float ux = x * inv, uy = y * inv, uz = z * inv;
    float basis[16], gq[16];
    rs::sh_basis(degree, ux, uy, uz, basis);
    float r = 0.f, g = 0.f, b = 0.f;
#pragma unroll
    for (int q = 0; q < 16; ++q)
      if (q < nb) { r += basis[q] * cf[q * 3]; g += basis[q] * cf[q * 3 + 1]; b += basis[q] * cf[q * 3 + 2]; }
    const float vr = (live && r + 0.5f >= 0.f) ? vc.x : 0.f, vg = (live && g + 0.5f >= 0.f) ? vc.y : 0.f,
                vb = (live && b + 0.5f >= 0.f) ? vc.z : 0.f;
    vrgb[e * 3] = vr; vrgb[e * 3 + 1] = vg; vrgb[e * 3 + 2] = vb;
#pragma unroll
    for (int q = 0; q < 16; ++q) gq[q] = q < nb ? vr * cf[q * 3] + vg * cf[q * 3 + 1] + vb * cf[q * 3 + 2] : 0.f;
    float bx, by, bz;
    rs::sh_basis_vjp(degree, ux, uy, uz, gq, bx, by, bz);
    const float dd = ux * bx + uy * by + uz * bz;
    vmx += (bx - ux * dd) * inv; vmy += (by - uy * dd) * inv; vmz += (bz - uz * dd) * inv;
  }
  v_means[n * 3] = vmx; v_means[n * 3 + 1] = vmy; v_means[n * 3 + 2] = vmz;
}

template <int W, int MINB>
__global__ void __launch_bounds__(CB, MINB)
sh_coeffs_gather_kernel(int degree, int K, int N, const float* __restrict__ means, const PeerSources src,
                        float* __restrict__ v_coeffs) {
  extern __shared__ float s_rows[];
  __shared__ float4 s_campos[RS_MAX_SOURCES];
  __shared__ const float* s_rowptr[RS_MAX_SOURCES];
  __shared__ int s_total;
  const int t = threadIdx.x;
  const int n0 = blockIdx.x * CB;
  const int count = min(CB, N - n0);
  const int row = K * 3, RS = row | 1;
  const int nb = (degree + 1) * (degree + 1);
  // flat list of sources (rank g, camera c), in (g, c) order: camera position + base of its vrgb rows
  if (t == 0) {
    int k = 0;
    for (int g = 0; g < src.n; ++g)
      for (int c = 0; c < src.cams[g] && k < RS_MAX_SOURCES; ++c, ++k)
        s_rowptr[k] = reinterpret_cast<const float*>(src.base[g] + RS_PEER_HEADER_BYTES) + (size_t)c * N * 3;
    s_total = k;
  }
  __syncthreads();
  const int total = s_total;
  if (t < total) {
    int k = t, g = 0;
    while (k >= src.cams[g]) { k -= src.cams[g]; ++g; }
    s_campos[t] = __ldcg(reinterpret_cast<const float4*>(src.base[g]) + k);
  }
  float acc[48];   // the sums live in registers
#pragma unroll
  for (int i = 0; i < 48; ++i) acc[i] = 0.f;
  __syncthreads();
  const int n = n0 + t;
  if (t < count) {
    const float mx = __ldg(means + n * 3), my = __ldg(means + n * 3 + 1), mz = __ldg(means + n * 3 + 2);
    // W (possibly remote) 12-byte rows in flight per thread before any is consumed
    for (int k0 = 0; k0 < total; k0 += W) {
      float3 v[W];
#pragma unroll
      for (int i = 0; i < W; ++i) {
        v[i] = make_float3(0.f, 0.f, 0.f);
        if (k0 + i < total) {
          const float* r = s_rowptr[k0 + i] + (size_t)n * 3;
          v[i] = make_float3(__ldcg(r), __ldcg(r + 1), __ldcg(r + 2));
        }
      }
#pragma unroll
      for (int i = 0; i < W; ++i) {
        if (v[i].x == 0.f && v[i].y == 0.f && v[i].z == 0.f) continue;   // culled / clamped: exact zeros
        const float4 cp = s_campos[k0 + i];
        const float x = mx - cp.x, y = my - cp.y, z = mz - cp.z;
        const float inv = 1.f / fmaxf(sqrtf(x * x + y * y + z * z), 1e-12f);
        float basis[16];
        rs::sh_basis(degree, x * inv, y * inv, z * inv, basis);
#pragma unroll
        for (int q = 0; q < 16; ++q)
          if (q < nb) {
            acc[q * 3] = fmaf(basis[q], v[i].x, acc[q * 3]);
            acc[q * 3 + 1] = fmaf(basis[q], v[i].y, acc[q * 3 + 1]);
            acc[q * 3 + 2] = fmaf(basis[q], v[i].z, acc[q * 3 + 2]);
          }
      }
    }
  }
  float* o = s_rows + t * RS;
#pragma unroll
  for (int i = 0; i < 48; ++i)
    if (i < row) o[i] = acc[i];      // coefficients beyond the active degree stay zero
  __syncthreads();
  smem_to_rows(s_rows, v_coeffs + (size_t)n0 * row, count, row, RS, t);
}

}  // namespace

// colors4[C,N,4] = (max(SH(dir)+0.5, 0) rgb, depth or 0); coeffs[N,K,3] shared by all cameras; entries whose
// radii are 0 skip the SH sum (their colour is the clamp of 0.5, as upstream's masked call leaves it).
extern "C" int rs_sh_colors_fwd(int degree, int K, int C, int N, const float* means, const float* coeffs,
                                const float* viewmats, const int32_t* radii, const float* depths, float* colors4,
                                void* stream) {
  RsSpan span__("rs_sh_colors_fwd", stream);
  if (degree < 0 || degree > 3 || K < (degree + 1) * (degree + 1) || K > 16 || C < 0 || N < 0) return RS_ERR_BAD_ARG;
  if (C == 0 || N == 0) return RS_OK;
  if (!means || !coeffs || !viewmats || !radii || !colors4) return RS_ERR_BAD_ARG;
  const size_t smem = sizeof(float) * CB * ((K * 3) | 1);
  sh_colors_fwd_kernel<<<rs_div_up(N, CB), CB, smem, (cudaStream_t)stream>>>(
      degree, K, C, N, means, coeffs, viewmats, (const int2*)radii, depths, (float4*)colors4);
  RS_RETURN_LAST_ERROR();
}

// VJP: v_colors4[C,N,4] -> v_coeffs[N,K,3] and v_means[N,3] (both summed over cameras, overwritten) and, if
// has_depth, v_depths[C,N] (the 4th channel's gradient).
extern "C" int rs_sh_colors_bwd(int degree, int K, int C, int N, const float* means, const float* coeffs,
                                const float* viewmats, const int32_t* radii, const float* v_colors4, int has_depth,
                                float* v_coeffs, float* v_means, float* v_depths, void* stream) {
  RsSpan span__("rs_sh_colors_bwd", stream);
  if (degree < 0 || degree > 3 || K < (degree + 1) * (degree + 1) || K > 16 || C < 0 || N < 0) return RS_ERR_BAD_ARG;
  if (N == 0) return RS_OK;
  if (!means || !coeffs || !viewmats || !radii || !v_colors4 || !v_coeffs || !v_means || (has_depth && !v_depths))
    return RS_ERR_BAD_ARG;
  const size_t smem = (C == 1 ? 1 : 2) * sizeof(float) * CB * ((K * 3) | 1);
  if (smem > 48 * 1024) {  // two staging rows at K = 16 need the opt-in shared-memory limit
    cudaError_t e = cudaFuncSetAttribute(sh_colors_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { rs_set_last_cuda_error((int)e); return RS_ERR_LAUNCH; }
  }
  sh_colors_bwd_kernel<<<rs_div_up(N, CB), CB, smem, (cudaStream_t)stream>>>(
      degree, K, C, N, means, coeffs, viewmats, (const int2*)radii, (const float4*)v_colors4, has_depth, v_coeffs,
      v_means, v_depths);
  RS_RETURN_LAST_ERROR();
}

// ---- camera-sharded split of rs_sh_colors_bwd (see the kernels above).  `region` is the caller's peer-visible
// buffer of rs_sh_region_bytes(C, N) bytes; v_coeffs is NOT produced here but by rs_sh_coeffs_gather on every rank.
extern "C" long long rs_sh_region_bytes(int C, int N) {
  const long long rows = 12ll * (long long)(C > 0 ? C : 0) * (long long)(N > 0 ? N : 0);
  return (long long)RS_PEER_HEADER_BYTES + (rows + 15) / 16 * 16;
}

extern "C" int rs_sh_colors_bwd_local(int degree, int K, int C, int N, const float* means, const float* coeffs,
                                      const float* viewmats, const int32_t* radii, const float* v_colors4,
                                      int has_depth, void* region, float* v_means, float* v_depths, void* stream) {
  RsSpan span__("rs_sh_colors_bwd_local", stream);
  if (degree < 0 || degree > 3 || K < (degree + 1) * (degree + 1) || K > 16 || C < 0 || N < 0) return RS_ERR_BAD_ARG;
  if (C > RS_PEER_HEADER_BYTES / 16) return RS_ERR_UNSUPPORTED;
  if (N == 0 || C == 0) return RS_OK;
  if (!means || !coeffs || !viewmats || !radii || !v_colors4 || !region || !v_means || (has_depth && !v_depths))
    return RS_ERR_BAD_ARG;
  const size_t smem = sizeof(float) * CB * ((K * 3) | 1);
  sh_colors_bwd_local_kernel<<<rs_div_up(N, CB), CB, smem, (cudaStream_t)stream>>>(
      degree, K, C, N, means, coeffs, viewmats, (const int2*)radii, (const float4*)v_colors4, has_depth, (char*)region,
      v_means, v_depths);
  RS_RETURN_LAST_ERROR();
}

// regions[n_sources]: device-visible base pointers (local, peer-mapped or all-gathered copies) of the regions the
// ranks filled with rs_sh_colors_bwd_local, cams[n_sources]: cameras in each.  v_coeffs[N,K,3] is overwritten with
// the sum over all sources, accumulated in source order.
extern "C" int rs_sh_coeffs_gather(int degree, int K, int N, const float* means, const void* const* regions,
                                   const int* cams, int n_sources, float* v_coeffs, void* stream) {
  RsSpan span__("rs_sh_coeffs_gather", stream);
  if (degree < 0 || degree > 3 || K < (degree + 1) * (degree + 1) || K > 16 || N < 0) return RS_ERR_BAD_ARG;
  if (n_sources < 0 || n_sources > RS_MAX_PEERS) return RS_ERR_UNSUPPORTED;
  if (N == 0) return RS_OK;
  if (!means || !v_coeffs || (n_sources > 0 && (!regions || !cams))) return RS_ERR_BAD_ARG;
  PeerSources src;
  src.n = n_sources;
  long long total_cams = 0;
  for (int g = 0; g < n_sources; ++g) total_cams += cams[g] > 0 ? cams[g] : 0;
  if (total_cams > RS_MAX_SOURCES || total_cams > CB) return RS_ERR_UNSUPPORTED;
  for (int g = 0; g < RS_MAX_PEERS; ++g) {
    src.base[g] = g < n_sources ? (const char*)regions[g] : nullptr;
    src.cams[g] = g < n_sources ? cams[g] : 0;
    if (g < n_sources && (!regions[g] || cams[g] < 0 || cams[g] > RS_PEER_HEADER_BYTES / 16)) return RS_ERR_BAD_ARG;
  }
  const size_t smem = sizeof(float) * CB * ((K * 3) | 1);
  // 4 source rows in flight per thread and 5 CTAs per SM (96 registers): 0.113 ms for 8 sources of 1 M Gaussians, against
  // 0.123 ms with 8 in flight at 4 CTAs per SM; storing the rows straight from registers (no staging) is slower (0.127)
  sh_coeffs_gather_kernel<4, 5><<<rs_div_up(N, CB), CB, smem, (cudaStream_t)stream>>>(degree, K, N, means, src, v_coeffs);
  RS_RETURN_LAST_ERROR();
}
