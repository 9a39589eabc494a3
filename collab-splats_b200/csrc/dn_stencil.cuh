// Depth -> normal stencil of the RaDe-GS depth-normal consistency term, shared by the fused loss (loss.cu) and the
// get_outputs epilogue (outputs.cu).  Restates collab_splats/utils/camera_utils.py:176-279 for one interior pixel:
// points P = depth * K^-1 (x + .5, y + .5, 1), central differences along rows and columns, cross product, normalise
// (border pixels get a zero normal, i.e. an error of 1 and no gradient).
#pragma once
#include <cuda_runtime.h>

namespace rs {

struct V3 { float x, y, z; };
__device__ __forceinline__ V3 cross3(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

// One depth map: returns err = 1 - <N, n_depth>; with upstream weight g it accumulates -g * n_depth into vN (the
// gradient of the rendered normal) and scatters g's gradient into the four neighbouring depths (v_depth may be
// null: forward only).
__device__ __forceinline__ float dn_term(const float* __restrict__ depth, float* __restrict__ v_depth, int x, int y,
                                         int W, int H, float ifx, float ify, V3 N, float g, V3& vN) {
  if (x < 1 || y < 1 || x >= W - 1 || y >= H - 1) return 1.0f;  // border normals are 0 -> err = 1, no gradient
  const float cx = 0.5f * W, cy = 0.5f * H;
  const float rxm = (x - 0.5f - cx) * ifx, rx0 = (x + 0.5f - cx) * ifx, rxp = (x + 1.5f - cx) * ifx;
  const float rym = (y - 0.5f - cy) * ify, ry0 = (y + 0.5f - cy) * ify, ryp = (y + 1.5f - cy) * ify;
  const size_t p = (size_t)y * W + x;
  const float du = __ldg(depth + p - W), dd = __ldg(depth + p + W), dl = __ldg(depth + p - 1), dr = __ldg(depth + p + 1);
  // d_row = P(x,y+1) - P(x,y-1);  d_col = P(x+1,y) - P(x-1,y);  P = depth * (rx, ry, 1)
  const V3 a = {(dd - du) * rx0, dd * ryp - du * rym, dd - du};
  const V3 b = {dr * rxp - dl * rxm, (dr - dl) * ry0, dr - dl};
  const V3 c = cross3(a, b);
  // 1 / max(|c|, 1e-12) and |c| from one SFU reciprocal square root (2 ulp) instead of an IEEE sqrt and an IEEE division
  const float l2 = c.x * c.x + c.y * c.y + c.z * c.z;
  const float inv = rsqrtf(fmaxf(l2, 1e-24f));
  const float len = l2 * inv;
  const V3 n = {c.x * inv, c.y * inv, c.z * inv};
  const float dot = N.x * n.x + N.y * n.y + N.z * n.z;
  // gradients: err = 1 - N.n, upstream weight g
  vN.x -= g * n.x; vN.y -= g * n.y; vN.z -= g * n.z;
  if (len > 1e-12f && v_depth) {
    const V3 vn = {-g * N.x, -g * N.y, -g * N.z};
    const float nd = n.x * vn.x + n.y * vn.y + n.z * vn.z;
    const V3 vc = {(vn.x - n.x * nd) * inv, (vn.y - n.y * nd) * inv, (vn.z - n.z * nd) * inv};
    const V3 va = cross3(b, vc);   // d(a x b)/da ^T vc = b x vc
    const V3 vb = cross3(vc, a);   // d(a x b)/db ^T vc = vc x a
    // a depends on dd (+) and du (-); b on dr (+) and dl (-)
    atomicAdd(v_depth + p + W, va.x * rx0 + va.y * ryp + va.z);
    atomicAdd(v_depth + p - W, -(va.x * rx0 + va.y * rym + va.z));
    atomicAdd(v_depth + p + 1, vb.x * rxp + vb.y * ry0 + vb.z);
    atomicAdd(v_depth + p - 1, -(vb.x * rxm + vb.y * ry0 + vb.z));
  }
  return 1.0f - dot;
}


}  // namespace rs
