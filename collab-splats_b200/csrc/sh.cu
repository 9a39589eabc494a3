// Real spherical harmonics (degree <= 3) -> RGB, forward and backward.
// Replaces gsplat `spherical_harmonics`: SURVEY.md row a6; reference call sites
// collab_splats/models/rade_features_model.py:430-434 (direct) and the sh_degree path of
// rasterization() (collab_splats/models/rade_gs_model.py:454).
//
// HBM-bound: 12*K + 24 B per element forward.  One thread per element.  Coefficients may be shared
// by all cameras (`n_coeff_rows` < n_elems: row = element % n_coeff_rows), which avoids the [C,N,K,3]
// expansion upstream materialises; the backward then loops over cameras per Gaussian so the shared
// coefficient gradient is accumulated in registers without atomics.
#include "common.cuh"
#include "rade_math.cuh"

namespace {

constexpr int SB = 256;

__global__ void __launch_bounds__(SB)
sh_fwd_kernel(int degree, int K, long long n_elems, long long n_rows, const float* __restrict__ dirs,
              const float* __restrict__ coeffs, const uint8_t* __restrict__ masks, float* __restrict__ colors) {
  const long long e = (long long)blockIdx.x * SB + threadIdx.x;
  if (e >= n_elems) return;
  float r = 0.f, g = 0.f, b = 0.f;
  if (!masks || masks[e]) {
    float x = __ldg(dirs + e * 3), y = __ldg(dirs + e * 3 + 1), z = __ldg(dirs + e * 3 + 2);
    float inv = 1.f / fmaxf(sqrtf(x * x + y * y + z * z), 1e-12f);
    float basis[16];
    rs::sh_basis(degree, x * inv, y * inv, z * inv, basis);
    const float* cf = coeffs + (e % n_rows) * (long long)K * 3;
    const int nb = (degree + 1) * (degree + 1);
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      if (q < nb) {
        r += basis[q] * __ldg(cf + q * 3);
        g += basis[q] * __ldg(cf + q * 3 + 1);
        b += basis[q] * __ldg(cf + q * 3 + 2);
      }
    }
  }
  colors[e * 3] = r; colors[e * 3 + 1] = g; colors[e * 3 + 2] = b;
}

// thread per coefficient row; loops over the n_elems / n_rows elements that share it
__global__ void __launch_bounds__(SB)
sh_bwd_kernel(int degree, int K, long long n_elems, long long n_rows, const float* __restrict__ dirs,
              const float* __restrict__ coeffs, const uint8_t* __restrict__ masks,
              const float* __restrict__ v_colors, float* __restrict__ v_coeffs, float* __restrict__ v_dirs) {
  const long long row = (long long)blockIdx.x * SB + threadIdx.x;
  if (row >= n_rows) return;
  const int nb = (degree + 1) * (degree + 1);
  const float* cf = coeffs + row * (long long)K * 3;
  float acc[48];
#pragma unroll
  for (int i = 0; i < 48; ++i) acc[i] = 0.f;
  for (long long e = row; e < n_elems; e += n_rows) {
    float vx = 0.f, vy = 0.f, vz = 0.f;
    if (!masks || masks[e]) {
      float x = __ldg(dirs + e * 3), y = __ldg(dirs + e * 3 + 1), z = __ldg(dirs + e * 3 + 2);
      float inv = 1.f / fmaxf(sqrtf(x * x + y * y + z * z), 1e-12f);
      float ux = x * inv, uy = y * inv, uz = z * inv;
      float vr = __ldg(v_colors + e * 3), vg = __ldg(v_colors + e * 3 + 1), vb = __ldg(v_colors + e * 3 + 2);
      float basis[16], gq[16];
      rs::sh_basis(degree, ux, uy, uz, basis);
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        gq[q] = 0.f;
        if (q < nb) {
          acc[q * 3] += basis[q] * vr; acc[q * 3 + 1] += basis[q] * vg; acc[q * 3 + 2] += basis[q] * vb;
          if (v_dirs) gq[q] = vr * __ldg(cf + q * 3) + vg * __ldg(cf + q * 3 + 1) + vb * __ldg(cf + q * 3 + 2);
        }
      }
      if (v_dirs) {
        float bx, by, bz;
        rs::sh_basis_vjp(degree, ux, uy, uz, gq, bx, by, bz);
        float dd = ux * bx + uy * by + uz * bz;
        vx = (bx - ux * dd) * inv; vy = (by - uy * dd) * inv; vz = (bz - uz * dd) * inv;
      }
    }
    if (v_dirs) { v_dirs[e * 3] = vx; v_dirs[e * 3 + 1] = vy; v_dirs[e * 3 + 2] = vz; }
  }
  float* vc = v_coeffs + row * (long long)K * 3;
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    if (q < K) {
      vc[q * 3] = q < nb ? acc[q * 3] : 0.f;
      vc[q * 3 + 1] = q < nb ? acc[q * 3 + 1] : 0.f;
      vc[q * 3 + 2] = q < nb ? acc[q * 3 + 2] : 0.f;
    }
  }
}

}  // namespace

extern "C" int rs_sh_fwd(int degree, int K, long long n_elems, long long n_coeff_rows, const float* dirs,
                         const float* coeffs, const uint8_t* masks, float* colors, void* stream) {
  RsSpan span__("rs_sh_fwd", stream);
  if (degree < 0 || degree > 3 || K < (degree + 1) * (degree + 1) || K > 16 || n_elems < 0 || n_coeff_rows <= 0)
    return n_elems == 0 ? RS_OK : RS_ERR_BAD_ARG;
  if (n_elems == 0) return RS_OK;
  if (!dirs || !coeffs || !colors || (n_elems % n_coeff_rows) != 0) return RS_ERR_BAD_ARG;
  sh_fwd_kernel<<<rs_div_up(n_elems, SB), SB, 0, (cudaStream_t)stream>>>(degree, K, n_elems, n_coeff_rows, dirs,
                                                                        coeffs, masks, colors);
  RS_RETURN_LAST_ERROR();
}

extern "C" int rs_sh_bwd(int degree, int K, long long n_elems, long long n_coeff_rows, const float* dirs,
                         const float* coeffs, const uint8_t* masks, const float* v_colors, float* v_coeffs,
                         float* v_dirs, void* stream) {
  RsSpan span__("rs_sh_bwd", stream);
  if (degree < 0 || degree > 3 || K < (degree + 1) * (degree + 1) || K > 16 || n_elems < 0 || n_coeff_rows <= 0)
    return n_elems == 0 ? RS_OK : RS_ERR_BAD_ARG;
  if (n_elems == 0) return RS_OK;
  if (!dirs || !coeffs || !v_colors || !v_coeffs || (n_elems % n_coeff_rows) != 0) return RS_ERR_BAD_ARG;
  sh_bwd_kernel<<<rs_div_up(n_coeff_rows, SB), SB, 0, (cudaStream_t)stream>>>(
      degree, K, n_elems, n_coeff_rows, dirs, coeffs, masks, v_colors, v_coeffs, v_dirs);
  RS_RETURN_LAST_ERROR();
}
