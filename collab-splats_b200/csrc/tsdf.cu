// TSDF fusion of rendered depth / colour maps on the device: the per-frame loop of the
// reference's meshing exporter, collab_splats/utils/mesh.py:1562-1632, which copies every rendered frame to the host
// and integrates it into Open3D's ScalableTSDFVolume on the CPU.  Same algorithm, restated for the GPU:
//   * the volume is a hash map (open addressing, 64-bit packed unit coordinates) from "volume unit" (16^3 voxels,
//     unit length = 16 * voxel_length) to a slot of a caller-owned pool (tsdf, weight, rgb per voxel);
//   * tsdf_touch_kernel: every `stride`-th depth pixel is back-projected to the world and the units within
//     +-sdf_trunc of the point are inserted (allocated on first touch) and appended ONCE per frame to the frame's
//     touched list (Open3D: ScalableTSDFVolume::Integrate, first half);
//   * tsdf_integrate_kernel: persistent CTAs walk the touched list; every voxel of a touched unit is projected
//     into the frame, and if the pixel holds a depth and the signed distance along the ray is > -sdf_trunc the
//     running weighted mean of tsdf and colour is updated, weight += 1 (Open3D:
//     UniformTSDFVolume::IntegrateWithDepthToCameraDistanceMultiplier).
// Arithmetic contract (shared with oracle/tsdf_oracle.py, so the two agree bit for bit): fp32, every operation
// individually rounded in the order written below (no FMA contraction).  HBM-bound: 20 B read + 20 B written per
// voxel of a touched unit, 80 KB per unit and frame.
#include "common.cuh"

namespace {

constexpr int TSDF_RES = 16;
constexpr int TSDF_VOX = TSDF_RES * TSDF_RES * TSDF_RES;
constexpr unsigned long long TSDF_EMPTY = ~0ull;

struct TsdfFrame {
  float fx, fy, cx, cy;
  float E[12];    // world -> camera, rows of [R|t]
  float P[12];    // camera -> world
  int W, H;
  float depth_trunc, voxel_length, sdf_trunc;
};

__device__ __forceinline__ unsigned long long pack_unit(int x, int y, int z) {
  const unsigned long long B = 1ull << 20;
  return ((unsigned long long)(x + (long long)B) << 42) | ((unsigned long long)(y + (long long)B) << 21) |
         (unsigned long long)(z + (long long)B);
}

__device__ __forceinline__ unsigned int hash_unit(unsigned long long k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return (unsigned int)k;
}

// find-or-insert; returns the hash-table index of the unit or -1 (table full / pool full -> *overflow raised).  The
// thread that wins the key allocates the pool slot and publishes it in vals[h]; nobody waits for it inside this
// kernel -- the frame's stamp and touched list are kept per hash index, and the integrate kernel (next launch) reads
// vals[h].
__device__ int unit_index(unsigned long long key, int ux, int uy, int uz, unsigned long long* keys, int* vals,
                          unsigned int cap_mask, int* pool_count, int max_units, int* unit_xyz, int* overflow) {
  unsigned int h = hash_unit(key) & cap_mask;
  for (unsigned int probe = 0; probe <= cap_mask; ++probe, h = (h + 1) & cap_mask) {
    unsigned long long cur = keys[h];
    if (cur == TSDF_EMPTY) {
      cur = atomicCAS(keys + h, TSDF_EMPTY, key);
      if (cur == TSDF_EMPTY) {   // ours: allocate a pool slot and publish it
        const int slot = atomicAdd(pool_count, 1);
        if (slot >= max_units) { *overflow = 1; vals[h] = -2; return -1; }
        unit_xyz[slot * 3] = ux; unit_xyz[slot * 3 + 1] = uy; unit_xyz[slot * 3 + 2] = uz;
        vals[h] = slot;
        return (int)h;
      }
    }
    if (cur == key) return (int)h;
  }
  *overflow = 1;
  return -1;
}

__global__ void __launch_bounds__(256)
tsdf_touch_kernel(const float* __restrict__ depth, TsdfFrame f, int stride, int frame_id, unsigned long long* keys,
                  int* vals, unsigned int cap_mask, int* counters /* {pool_count, touched_count, overflow} */,
                  int max_units, int* unit_xyz, int* stamps, int* touched) {
  const int sw = (f.W + stride - 1) / stride, sh = (f.H + stride - 1) / stride;
  const int s = blockIdx.x * 256 + threadIdx.x;
  if (s >= sw * sh) return;
  const int i = (s / sw) * stride, j = (s % sw) * stride;   // row, column
  const float d = __ldg(depth + (size_t)i * f.W + j);
  if (!(d > 0.f) || d >= f.depth_trunc) return;
  const float x = __fdiv_rn(__fmul_rn(__fsub_rn((float)j, f.cx), d), f.fx);
  const float y = __fdiv_rn(__fmul_rn(__fsub_rn((float)i, f.cy), d), f.fy);
  float w[3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
    w[r] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(f.P[r * 4], x), __fmul_rn(f.P[r * 4 + 1], y)),
                               __fmul_rn(f.P[r * 4 + 2], d)), f.P[r * 4 + 3]);
  const float ul = __fmul_rn(f.voxel_length, (float)TSDF_RES);
  int lo[3], hi[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    lo[r] = (int)floorf(__fdiv_rn(__fsub_rn(w[r], f.sdf_trunc), ul));
    hi[r] = (int)floorf(__fdiv_rn(__fadd_rn(w[r], f.sdf_trunc), ul));
  }
  for (int ux = lo[0]; ux <= hi[0]; ++ux)
    for (int uy = lo[1]; uy <= hi[1]; ++uy)
      for (int uz = lo[2]; uz <= hi[2]; ++uz) {
        if (abs(ux) >= (1 << 20) || abs(uy) >= (1 << 20) || abs(uz) >= (1 << 20)) { counters[2] = 1; continue; }
        const int h = unit_index(pack_unit(ux, uy, uz), ux, uy, uz, keys, vals, cap_mask, counters, max_units,
                                 unit_xyz, counters + 2);
        if (h < 0) continue;
        if (atomicExch(stamps + h, frame_id) != frame_id) {
          const int at = atomicAdd(counters + 1, 1);
          if (at < max_units) touched[at] = h; else counters[2] = 1;
        }
      }
}

__global__ void __launch_bounds__(256)
tsdf_integrate_kernel(const float* __restrict__ depth, const unsigned char* __restrict__ color_u8,
                      const float* __restrict__ color_f32, TsdfFrame f, const int* __restrict__ counters,
                      const int* __restrict__ touched, const int* __restrict__ vals, int max_units,
                      const int* __restrict__ unit_xyz, float* __restrict__ tsdf, float* __restrict__ weight,
                      float* __restrict__ rgb) {
  const int n_touched = min(counters[1], max_units);
  const float inv_fx = __fdiv_rn(1.f, f.fx), inv_fy = __fdiv_rn(1.f, f.fy), inv_trunc = __fdiv_rn(1.f, f.sdf_trunc);
  const float safe_w = __fsub_rn((float)f.W, 0.0001f), safe_h = __fsub_rn((float)f.H, 0.0001f);
  for (int ti = blockIdx.x; ti < n_touched; ti += gridDim.x) {
    const int slot = __ldg(vals + __ldg(touched + ti));
    if (slot < 0) continue;
    const int ox = __ldg(unit_xyz + slot * 3) * TSDF_RES, oy = __ldg(unit_xyz + slot * 3 + 1) * TSDF_RES,
              oz = __ldg(unit_xyz + slot * 3 + 2) * TSDF_RES;
    for (int v = threadIdx.x; v < TSDF_VOX; v += 256) {   // voxel index x*256 + y*16 + z: z fastest, coalesced
      const int vx = v >> 8, vy = (v >> 4) & 15, vz = v & 15;
      const float wx = __fmul_rn(__fadd_rn((float)(ox + vx), 0.5f), f.voxel_length);
      const float wy = __fmul_rn(__fadd_rn((float)(oy + vy), 0.5f), f.voxel_length);
      const float wz = __fmul_rn(__fadd_rn((float)(oz + vz), 0.5f), f.voxel_length);
      float pc[3];
#pragma unroll
      for (int r = 0; r < 3; ++r)
        pc[r] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(f.E[r * 4], wx), __fmul_rn(f.E[r * 4 + 1], wy)),
                                    __fmul_rn(f.E[r * 4 + 2], wz)), f.E[r * 4 + 3]);
      if (!(pc[2] > 0.f)) continue;
      const float u_f = __fadd_rn(__fadd_rn(__fdiv_rn(__fmul_rn(pc[0], f.fx), pc[2]), f.cx), 0.5f);
      const float v_f = __fadd_rn(__fadd_rn(__fdiv_rn(__fmul_rn(pc[1], f.fy), pc[2]), f.cy), 0.5f);
      if (!(u_f >= 0.0001f && u_f < safe_w && v_f >= 0.0001f && v_f < safe_h)) continue;
      const int u = (int)u_f, vv = (int)v_f;
      const size_t pix = (size_t)vv * f.W + u;
      const float d = __ldg(depth + pix);
      if (!(d > 0.f) || d >= f.depth_trunc) continue;
      const float xn = __fmul_rn(__fsub_rn((float)u, f.cx), inv_fx), yn = __fmul_rn(__fsub_rn((float)vv, f.cy), inv_fy);
      const float mult = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(xn, xn), __fmul_rn(yn, yn)), 1.f));
      const float sdf = __fmul_rn(__fsub_rn(d, pc[2]), mult);
      if (!(sdf > -f.sdf_trunc)) continue;
      const float t = fminf(1.f, __fmul_rn(sdf, inv_trunc));
      const size_t e = (size_t)slot * TSDF_VOX + v;
      const float w0 = weight[e], w1 = __fadd_rn(w0, 1.f);
      tsdf[e] = __fdiv_rn(__fadd_rn(__fmul_rn(tsdf[e], w0), t), w1);
      if (rgb) {
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          const float cnew = color_u8 ? (float)__ldg(color_u8 + pix * 3 + ch) : __ldg(color_f32 + pix * 3 + ch);
          rgb[e * 3 + ch] = __fdiv_rn(__fadd_rn(__fmul_rn(rgb[e * 3 + ch], w0), cnew), w1);
        }
      }
      weight[e] = w1;
    }
  }
}

}  // namespace

// One frame into the volume.  Caller-owned state (all device memory, zero / 0xFF initialised as noted):
//   keys u64[capacity] (all 0xFF), vals i32[capacity] (all -1), capacity a power of two > max_units (2x or more keeps probes short);
//   counters i32[4] = {units allocated, units touched by this frame, overflow flag, -} (zero at volume creation;
//   this call resets [1]); unit_xyz i32[max_units,3]; stamps i32[capacity] (zero; frame ids start at 1, increasing);
//   touched i32[max_units]; tsdf / weight f32[max_units,4096], rgb f32[max_units,4096,3] or NULL (zero).
// depth f32[H,W] (0 = no measurement; values >= depth_trunc are ignored), colour u8[H,W,3] or f32[H,W,3] (one of the
// two, or both NULL when rgb is NULL); intrinsics (fx, fy, cx, cy); extrinsic = world -> camera [3x4 row major],
// pose = its inverse.  No host synchronisation: the integrate pass reads the touched count from device memory.
extern "C" int rs_tsdf_integrate(const float* depth, const unsigned char* color_u8, const float* color_f32, int width,
                                 int height, float fx, float fy, float cx, float cy, const float* extrinsic_3x4,
                                 const float* pose_3x4, float voxel_length, float sdf_trunc, float depth_trunc,
                                 int stride, int frame_id, unsigned long long* keys, int* vals, long long capacity,
                                 int* counters, int max_units, int* unit_xyz, int* stamps, int* touched, float* tsdf,
                                 float* weight, float* rgb, void* stream) {
  RsSpan span__("rs_tsdf_integrate", stream);
  if (!depth || width <= 0 || height <= 0 || !extrinsic_3x4 || !pose_3x4 || !(voxel_length > 0.f) ||
      !(sdf_trunc > 0.f) || stride <= 0 || frame_id <= 0 || !keys || !vals || !counters || !unit_xyz || !stamps ||
      !touched || !tsdf || !weight || max_units <= 0)
    return RS_ERR_BAD_ARG;
  if (capacity < 2 || (capacity & (capacity - 1)) != 0 || capacity <= max_units || capacity > (1ll << 31))
    return RS_ERR_BAD_ARG;
  if (rgb && !color_u8 && !color_f32) return RS_ERR_BAD_ARG;
  TsdfFrame f;
  f.fx = fx; f.fy = fy; f.cx = cx; f.cy = cy; f.W = width; f.H = height;
  f.depth_trunc = depth_trunc; f.voxel_length = voxel_length; f.sdf_trunc = sdf_trunc;
  for (int i = 0; i < 12; ++i) { f.E[i] = extrinsic_3x4[i]; f.P[i] = pose_3x4[i]; }
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(counters + 1, 0, sizeof(int), st);
  if (e != cudaSuccess) { rs_set_last_cuda_error((int)e); return RS_ERR_LAUNCH; }
  const int sw = (width + stride - 1) / stride, sh = (height + stride - 1) / stride;
  tsdf_touch_kernel<<<rs_div_up((long long)sw * sh, 256), 256, 0, st>>>(depth, f, stride, frame_id, keys, vals,
                                                                        (unsigned int)(capacity - 1), counters,
                                                                        max_units, unit_xyz, stamps, touched);
  tsdf_integrate_kernel<<<148 * 8, 256, 0, st>>>(depth, color_u8, color_f32, f, counters, touched, vals, max_units,
                                                unit_xyz, tsdf, weight, rgb);
  rs_count_launches(1);
  RS_RETURN_LAST_ERROR();
}
