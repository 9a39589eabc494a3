// Per-tile front-to-back alpha compositing (forward) and its back-to-front gradient (backward) with the
// RaDe-GS outputs: colours (N-D), alpha, expected ray-depth, median depth, normals.
// Replaces gsplat-rade `rasterize_to_pixels` fwd/bwd: SURVEY.md rows a10/a11, Appendix A8/A9; reached
// from collab_splats/models/rade_gs_model.py:439-465 and rade_features_model.py:450-476.
//
// Design (B200, FP32 SIMT -- compositing is not a dense contraction):
//  * per-(camera,Gaussian) inputs are first packed into one 64-byte geometry record (+ a padded colour
//    row), so a tile gathers each Gaussian with 16-byte cp.async copies straight into shared memory;
//    conics are pre-scaled by log2(e) (alpha = o * 2^-sigma') and each record carries the half-extents of
//    its alpha >= 1/255 footprint;
//  * one CTA per 16x16 tile, 8 warps, each warp owns an 8x4 pixel block; batches are double-buffered
//    (ids two batches ahead, records one batch ahead) so the gather overlaps the blend;
//  * each warp ballots the batch against its own 8x4 rectangle (conservative footprint test, so the
//    result is identical to testing every pair) and only evaluates the survivors; warps stop as soon
//    as all 32 pixels have saturated, the CTA as soon as all warps have;
//  * backward: the D+4 blended channels collapse to ONE suffix accumulator per pixel
//    (w_i = <v_out, val_i>), per-Gaussian gradients are reduced across the warp with a recursive-halving
//    shuffle tree (K-1 shuffles for K values instead of 5K) and committed with one coalesced 64-byte
//    RED per (warp, Gaussian) into a packed gradient record.
#include "common.cuh"

namespace {

constexpr int RT = 256;  // threads per CTA (16x16 pixels)
#define RS_LOG2E 1.4426950408889634f
#define RS_LN2 0.6931471805599453f

template <int DP> struct Batch { static constexpr int value = DP <= 8 ? 256 : (DP <= 32 ? 128 : 64); };

// ------------------------------------------------------------------------------------------------ pack
__global__ void __launch_bounds__(256)
pack_geom_kernel(const float2* __restrict__ means2d, const float* __restrict__ conics,
                 const float* __restrict__ opacities, int opac_per_cam, const float* __restrict__ compensations,
                 int N, const float* __restrict__ ray_ts, const float2* __restrict__ ray_planes,
                 const float* __restrict__ normals, const int2* __restrict__ radii, long long n_elems,
                 int cull_mode, float4* __restrict__ geom) {
  const long long e = (long long)blockIdx.x * 256 + threadIdx.x;
  if (e >= n_elems) return;
  // a culled Gaussian never passes either footprint test (mode 0: negative extents; mode 1: tau' = -inf)
  float4 q0 = make_float4(0.f, 0.f, cull_mode == 0 ? -1e30f : 0.f, cull_mode == 0 ? -1e30f : 0.f);
  float4 q1 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 q2 = make_float4(0.f, 0.f, 0.f, 0.f), q3 = make_float4(0.f, 0.f, 0.f, -1e30f);
  bool live = true;
  if (radii) { int2 r = __ldg(radii + e); live = r.x > 0 && r.y > 0; }
  if (live) {
    float2 m = __ldg(means2d + e);
    float a = __ldg(conics + e * 3), b = __ldg(conics + e * 3 + 1), c = __ldg(conics + e * 3 + 2);
    float o = __ldg(opacities + (opac_per_cam ? e : e % N));
    if (compensations) o *= __ldg(compensations + e);  // antialiased mode: opacity * sqrt(det0/det)
    // footprint of alpha >= 1/255: sigma <= tau = ln(255 o).  Two conservative tests are supported (both can only
    // ever keep more pairs than the exact per-pair test would):
    //  cull_mode 0: bbox half extents sqrt(2 tau Sigma_ii) of the footprint ellipse, padded -> (hx, hy);
    //  cull_mode 1: exact ellipse-vs-rectangle test: sigma' (log2 units) is minimised over the rectangle of pixel
    //               centres (on the one or two edges facing the centre) and compared with tau' = log2(255 o) padded;
    //               the record carries the two edge-minimiser slopes (-b'/2c', -b'/2a') and tau'.
    float hx = 1e30f, hy = 1e30f, taup = 1e30f;
    float tau = logf(255.f * o);
    if (tau < -0.02f) {
      hx = -1e30f; hy = -1e30f; taup = -1e30f;  // alpha < 1/255 everywhere
    } else if (tau == tau) {
      float det = a * c - b * b;
      if (det > 0.f && isfinite(det) && a > 0.f && c > 0.f) {
        if (cull_mode == 0) {
          float s = 2.f * (tau + 0.03f) / det;
          float ex = sqrtf(s * c) * 1.001f + 0.02f, ey = sqrtf(s * a) * 1.001f + 0.02f;
          if (isfinite(ex) && isfinite(ey)) { hx = ex; hy = ey; }
        } else {
          float rx = -b / c, ry = -b / a;   // = -b'/(2c'), -b'/(2a') with the pre-scaled conic
          float tp = (tau + 0.03f) * RS_LOG2E * 1.0001f + 0.01f;
          if (isfinite(rx) && isfinite(ry) && isfinite(tp)) { hx = rx; hy = ry; taup = tp; } else { hx = 0.f; hy = 0.f; }
        }
      } else if (cull_mode != 0) { hx = 0.f; hy = 0.f; }
    } else if (cull_mode != 0) { hx = 0.f; hy = 0.f; }
    q0 = make_float4(m.x, m.y, hx, hy);
    q1 = make_float4(0.5f * RS_LOG2E * a, RS_LOG2E * b, 0.5f * RS_LOG2E * c, o);
    float2 rp = __ldg(ray_planes + e);
    q2 = make_float4(__ldg(ray_ts + e), rp.x, rp.y, 0.f);  // .w is overwritten with the flatten id below
    q3 = make_float4(__ldg(normals + e * 3), __ldg(normals + e * 3 + 1), __ldg(normals + e * 3 + 2), taup);
  }
  q2.w = __int_as_float((int)e);  // the record carries its own flatten id (bit pattern, never used as a float)
  geom[e * 4 + 0] = q0; geom[e * 4 + 1] = q1; geom[e * 4 + 2] = q2; geom[e * 4 + 3] = q3;
}

__global__ void __launch_bounds__(256)
pack_colors_kernel(const float* __restrict__ colors, long long rows, int D, int DP, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= rows * DP) return;
  const long long r = i / DP;
  const int k = (int)(i - r * DP);
  out[i] = k < D ? __ldg(colors + r * D + k) : 0.f;
}

// gradient record layout (16 floats per (camera, Gaussian)), S = sum over the pixels the Gaussian was blended into:
//   0 S v_sigma*dx  1 S v_sigma*dy | 2 ga 3 gb 4 gc (raw conic) | 5 S v_sigma | 6 g_ray_t 7 g_rpx 8 g_rpy |
//   9 gnx 10 gny 11 gnz | 12..15 colour channels 0..3 (only when the colours are padded to 4 channels, DP == 4)
// The compositing kernels leave the two per-Gaussian linear maps to this kernel (they only need per-Gaussian
// constants, which are in the geometry record):
//   v_means2d = (a*S0 + b*S1 + rpx*g_ray_t, b*S0 + c*S1 + rpy*g_ray_t)   (d sigma/d(dx,dy) with the raw conic + the
//                                                                          ray-plane term of t = ray_t + rp.d)
//   v_opacity = -S5 / o                                                   (alpha = o e^-sigma: v_o = -v_sigma / o)
// thread per Gaussian, looping over cameras, so the opacity gradient (shared by all cameras when the input
// opacity is [N]) is summed in a register.
__global__ void __launch_bounds__(256)
unpack_geom_grad_kernel(const float4* __restrict__ gg, const float4* __restrict__ geom,
                        const float2* __restrict__ abs_grad, int C, int N,
                        const float* __restrict__ opacities, int opac_per_cam,
                        const float* __restrict__ compensations, float2* __restrict__ v_means2d,
                        float2* __restrict__ v_means2d_abs, float* __restrict__ v_conics,
                        float* __restrict__ v_opacities, float* __restrict__ v_compensations,
                        float* __restrict__ v_ray_ts, float2* __restrict__ v_ray_planes,
                        float* __restrict__ v_normals, float* __restrict__ v_colors4, int color_per_cam, int D) {
  const int n = blockIdx.x * 256 + threadIdx.x;
  if (n >= N) return;
  float vo_sum = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
  for (int c = 0; c < C; ++c) {
    const long long e = (long long)c * N + n;
    const float4 g0 = gg[e * 4], g1 = gg[e * 4 + 1], g2 = gg[e * 4 + 2], g3 = gg[e * 4 + 3];
    const float4 q1 = __ldg(geom + e * 4 + 1), q2 = __ldg(geom + e * 4 + 2);
    const float ca = 2.f * RS_LN2 * q1.x, cb = RS_LN2 * q1.y, cc = 2.f * RS_LN2 * q1.z;  // raw conic
    v_means2d[e] = make_float2(ca * g0.x + cb * g0.y + q2.y * g1.z, cb * g0.x + cc * g0.y + q2.z * g1.z);
    if (v_means2d_abs) v_means2d_abs[e] = abs_grad[e];
    v_conics[e * 3] = g0.z; v_conics[e * 3 + 1] = g0.w; v_conics[e * 3 + 2] = g1.x;
    float go = q1.w > 0.f ? -g1.y / q1.w : 0.f;
    if (compensations) {
      const float comp = __ldg(compensations + e);
      v_compensations[e] = go * __ldg(opacities + (opac_per_cam ? e : n));
      go *= comp;
    }
    if (opac_per_cam) v_opacities[e] = go; else vo_sum += go;
    v_ray_ts[e] = g1.z;
    v_ray_planes[e] = make_float2(g1.w, g2.x);
    v_normals[e * 3] = g2.y; v_normals[e * 3 + 1] = g2.z; v_normals[e * 3 + 2] = g2.w;
    if (v_colors4) {
      if (color_per_cam) {
        float* o = v_colors4 + e * D;
        if (D > 0) o[0] = g3.x;
        if (D > 1) o[1] = g3.y;
        if (D > 2) o[2] = g3.z;
        if (D > 3) o[3] = g3.w;
      } else {
        c0 += g3.x; c1 += g3.y; c2 += g3.z; c3 += g3.w;
      }
    }
  }
  if (!opac_per_cam) v_opacities[n] = vo_sum;
  if (v_colors4 && !color_per_cam) {
    float* o = v_colors4 + (long long)n * D;
    if (D > 0) o[0] = c0;
    if (D > 1) o[1] = c1;
    if (D > 2) o[2] = c2;
    if (D > 3) o[3] = c3;
  }
}

__global__ void __launch_bounds__(256)
unpack_colors_grad_kernel(const float* __restrict__ cg, long long rows, int D, int DP, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= rows * D) return;
  const long long r = i / D;
  const int k = (int)(i - r * D);
  out[i] = cg[r * DP + k];
}

// ------------------------------------------------------------------------------------------------ shared
struct RasterArgs {
  const float4* geom;        // [C*N][4]
  const float* colors;       // [color_rows][DP]
  const float* backgrounds;  // [C][D] or null
  const float* Ks;           // [C][9]
  int C, N, W, H, tile_w, tile_h, D, color_per_cam;
  int ed_channel;            // >= 0: that colour channel is divided by max(alpha, 1e-10) ("ED" modes); -1: none
  int exact_cull;            // footprint test the records were packed for (!RS_RASTER_CULL_BBOX)
  int flags;                 // RS_RASTER_* options of the call
  const int* offsets;        // [C*tile_h*tile_w]
  const int* flatten_ids;    // [M]
  int M;
  // forward outputs == backward saved tensors
  float* out_colors;   // [C,H,W,D]
  float* out_alphas;   // [C,H,W]
  float* out_dexp;     // [C,H,W]
  float* out_dmed;     // [C,H,W]
  float* out_normals;  // [C,H,W,3]
  float* out_T;        // [C,H,W]   final transmittance
  int* last_ids;       // [C,H,W]
  int* median_ids;     // [C,H,W]
  // backward inputs / outputs
  const float* v_colors; const float* v_alphas; const float* v_dexp; const float* v_dmed; const float* v_normals;
  float* geom_grad;    // [C*N][16]  (layout: see unpack_geom_grad_kernel)
  float* color_grad;   // [color_rows][DP]  (DP > 4 only; DP == 4 colours travel in the geometry record)
  float* abs_grad;     // [C*N][2] or null
  unsigned long long* stats;  // forward, optional: {Q pairs a per-pixel loop would visit, Qc contributing pairs,
                              //                     warp evaluations, warp evaluations that blended}
  const int* end_dev;  // sync-free callers: device int32 = number of valid list entries (<= M, which is then the
                       // buffers' capacity), written by rs_offset_encode_dev; null: M is the count
};

template <int DP, int BATCH, int S = 2> struct Smem {
  float4 q0[S][BATCH], q1[S][BATCH], q2[S][BATCH], q3[S][BATCH];
  float col[S][BATCH][DP];
  int ids[2][BATCH];
  int red[8];
  // S > 2: ring of S staged batches with mbarrier hand-over instead of one CTA barrier per batch (see ring_* below)
  unsigned long long full[S], empty[S];
  int done;
};

// Footprint test of ONE Gaussian (per lane) against the warp's rectangle of pixel centres
// [rcx - hw, rcx + hw] x [rcy - hh, rcy + hh].  Conservative in both modes (see pack_geom_kernel).
template <int DP, int BATCH, int S>
__device__ __forceinline__ bool footprint_hit(const Smem<DP, BATCH, S>& s, int buf, int j, bool exact, float rcx, float rcy,
                                              float hw, float hh) {
  const float4 f = s.q0[buf][j];
  const float cx = f.x - rcx, cy = f.y - rcy;   // d = centre - pixel ranges over [cx - hw, cx + hw] x [cy - hh, cy + hh]
  if (!exact) return fabsf(cx) <= f.z + hw && fabsf(cy) <= f.w + hh;
  const float4 q = s.q1[buf][j];
  const float taup = s.q3[buf][j].w;
  const float xlo = cx - hw, xhi = cx + hw, ylo = cy - hh, yhi = cy + hh;
  const float dxe = fminf(fmaxf(0.f, xlo), xhi), dye = fminf(fmaxf(0.f, ylo), yhi);   // closest approach per axis
  // minimum of sigma' over the vertical line x = dxe (clamped to the rectangle) and the horizontal line y = dye:
  // the constrained minimum of a convex quadratic over a box lies on an edge facing its centre
  const float dyv = fminf(fmaxf(f.z * dxe, ylo), yhi);
  const float dxh = fminf(fmaxf(f.w * dye, xlo), xhi);
  const float sv = fmaf(fmaf(q.z, dyv, q.y * dxe), dyv, q.x * dxe * dxe);
  const float sh = fmaf(fmaf(q.x, dxh, q.y * dye), dxh, q.z * dye * dye);
  return fminf(sv, sh) <= taup;
}

// gather one batch (ids already in s.ids[buf]) with 16-byte cp.async copies; consecutive threads copy
// consecutive 16-byte chunks of one record, so each 64-byte record is one coalesced request
template <int DP, int BATCH, int NT = RT, int S = 2>
__device__ __forceinline__ void issue_gather(Smem<DP, BATCH, S>& s, int buf, int count, const RasterArgs& a, int t) {
  constexpr int CH = 4 + DP / 4;
  const int total = count * CH;
  for (int i = t; i < total; i += NT) {
    const int slot = i / CH, ch = i - slot * CH;
    const int id = s.ids[buf][slot];
    if (ch < 4) {
      float4* dst = (ch == 0 ? s.q0[buf] : ch == 1 ? s.q1[buf] : ch == 2 ? s.q2[buf] : s.q3[buf]) + slot;
      rs::cp_async16(dst, a.geom + (size_t)id * 4 + ch);
    } else {
      const int row = a.color_per_cam ? id : id % a.N;
      rs::cp_async16(&s.col[buf][slot][(ch - 4) * 4], a.colors + (size_t)row * DP + (ch - 4) * 4);
    }
  }
  rs::cp_async_commit();
}

// ---- staged-batch ring (S > 2 stages) with mbarrier hand-over.  The CTA barrier per batch couples the tile's four
// warps (13.8 % / 15.2 % of the backward's / forward's stall samples are warps waiting at it for the slowest one).  In
// the ring every warp gathers ITS 32 slots of a batch (ids in registers, distributed with shuffles), announces them
// with cp.async.mbarrier.arrive (the arrival fires when the copies have landed: nobody waits for its own copies), and
// a stage is refilled once all warps have released it -- so a warp may run up to S - 2 batches ahead of the slowest.
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared.b64 st, [%0];\n}" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_after_cp_async(unsigned long long* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared.b64 [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok) : "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// waits for the phase; gives up (returns false) when `*done` has reached `all` (every warp of the CTA is finished)
__device__ __forceinline__ bool mbar_wait_or_done(unsigned long long* bar, unsigned parity, const volatile int* done, int all) {
  while (!mbar_try_wait(bar, parity))
    if (*done >= all) return false;
  return true;
}
// this warp's 32 slots of a batch: lane l owns slot 32 warp + l (its flatten id in `my_id`); consecutive lanes copy
// consecutive 16-byte chunks of a record, as issue_gather does
template <int DP, int BATCH, int S>
__device__ __forceinline__ void ring_gather_share(Smem<DP, BATCH, S>& s, int stage, int count, int my_id,
                                                  const RasterArgs& a, int lane, int warp) {
  constexpr int CH = 4 + DP / 4;
#pragma unroll
  for (int r = 0; r < CH; ++r) {
    const int c = r * 32 + lane;
    const int sl = c / CH, ch = c - sl * CH;
    const int id = __shfl_sync(RS_FULL_MASK, my_id, sl);
    const int slot = warp * 32 + sl;
    if (slot < count) {
      if (ch < 4) {
        float4* dst = (ch == 0 ? s.q0[stage] : ch == 1 ? s.q1[stage] : ch == 2 ? s.q2[stage] : s.q3[stage]) + slot;
        rs::cp_async16(dst, a.geom + (size_t)id * 4 + ch);
      } else {
        const int row = a.color_per_cam ? id : id % a.N;
        rs::cp_async16(&s.col[stage][slot][(ch - 4) * 4], a.colors + (size_t)row * DP + (ch - 4) * 4);
      }
    }
  }
  mbar_arrive_after_cp_async(&s.full[stage]);
}

// End of a tile's list: the next tile's offset, or for the last tile the number of intersections -- a kernel parameter
// (DEV = false) or, when the caller never learned it (sync-free), a device word (DEV = true: one load through a selected
// pointer, no branch).  The 2-pixel kernels are compiled for both (a run-time test here costs the forward 8 registers
// and a CTA per SM); the generic kernels test at run time.
template <bool DEV>
__device__ __forceinline__ int list_end_of(const RasterArgs& a, int tile_id, int n_tiles) {
  if constexpr (DEV) {
    return __ldg(tile_id + 1 < n_tiles ? a.offsets + tile_id + 1 : a.end_dev);
  } else {
    return tile_id + 1 < n_tiles ? __ldg(a.offsets + tile_id + 1) : a.M;
  }
}
__device__ __forceinline__ int list_end_rt(const RasterArgs& a, int tile_id, int n_tiles) {
  return a.end_dev ? list_end_of<true>(a, tile_id, n_tiles) : list_end_of<false>(a, tile_id, n_tiles);
}

struct TileCtx {
  int cam, start, end, x0, y0, pxi, pyi;
  bool inside;
  float px, py, rcx, rcy;
};

__device__ __forceinline__ TileCtx tile_ctx(const RasterArgs& a, int lane, int warp) {
  TileCtx c;
  const int tile_id = blockIdx.x;
  const int tiles_per_cam = a.tile_w * a.tile_h;
  c.cam = tile_id / tiles_per_cam;
  const int tl = tile_id - c.cam * tiles_per_cam;
  const int tyi = tl / a.tile_w, txi = tl - tyi * a.tile_w;
  c.start = __ldg(a.offsets + tile_id);
  c.end = list_end_rt(a, tile_id, a.C * tiles_per_cam);
  c.x0 = txi * RS_TILE + (warp & 1) * 8;
  c.y0 = tyi * RS_TILE + (warp >> 1) * 4;
  c.pxi = c.x0 + (lane & 7);
  c.pyi = c.y0 + (lane >> 3);
  c.inside = c.pxi < a.W && c.pyi < a.H;
  c.px = c.pxi + 0.5f; c.py = c.pyi + 0.5f;
  c.rcx = c.x0 + 4.0f; c.rcy = c.y0 + 2.0f;
  return c;
}

// 1 / sqrt(((px-cx)/fx)^2 + ((py-cy)/fy)^2 + 1): ray distance -> z depth
__device__ __forceinline__ float inv_ray_len(const RasterArgs& a, int cam, float px, float py) {
  const float* K = a.Ks + (size_t)cam * 9;
  const float fx = __ldg(K), fy = __ldg(K + 4), cx = __ldg(K + 2), cy = __ldg(K + 5);
  const float rx = (px - cx) / fx, ry = (py - cy) / fy;
  return 1.0f / sqrtf(rx * rx + ry * ry + 1.0f);
}

// ---- TF32 tensor-core helpers for the wide (rade-features) colour rows: mma.sync m16n8k8 with both operands split
// into a TF32 head and tail (x = hi + lo); the three significant partial products are accumulated in fp32
// ("3xTF32": the dropped lo*lo term is ~2^-22 relative), so the blends keep fp32 accuracy.
__device__ __forceinline__ unsigned tf32_hi(float x) {
  unsigned r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void tf32_split(float x, unsigned& hi, unsigned& lo) {
  hi = tf32_hi(x);
  lo = tf32_hi(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int CM_PEND = 8;      // Gaussians parked per warp before a contraction
constexpr int CM_VSTRIDE = 36;  // row stride of the parked visibilities (conflict-free B-fragment loads)


// ------------------------------------------------------------------------------------------------ forward, wide rows
// Forward for DP >= 32 with the colour blend on the tensor cores.  The blend of a warp's 32 pixels with 8 Gaussians,
//     out[p][ch] += sum_g vis[p][g] * colour[g][ch],
// is M = 32 pixels (two 16-row tiles) x N = DP channels (8-column tiles) x K = 8 Gaussians: the warp parks the
// visibilities of the Gaussians it blends (and their slot in the staged batch) and contracts every 8th one -- or what
// is parked when the batch ends, since the staged colour rows are recycled -- against the colour rows in shared
// memory.  The DP accumulators per pixel become C fragments (each lane holds 4 pixels x DP/4 channels).  Everything
// else (alpha test, transmittance, depth / normal / median bookkeeping, early termination) is as in
// rasterize_fwd_kernel, written branch-free so that a skipped pair is an exact no-op.
template <int DP>
__device__ __forceinline__ void flush_blend_mma(float (&cf)[2][(DP + 7) / 8][4], const float* __restrict__ vis_w,
                                                const int* __restrict__ slot_w, const float (*col)[DP], int lane) {
  constexpr int NT = (DP + 7) / 8;
  const int gid = lane >> 2, tig = lane & 3;
  unsigned ah[2][4], al[2][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    tf32_split(vis_w[tig * CM_VSTRIDE + gid + 16 * mt], ah[mt][0], al[mt][0]);
    tf32_split(vis_w[tig * CM_VSTRIDE + gid + 8 + 16 * mt], ah[mt][1], al[mt][1]);
    tf32_split(vis_w[(tig + 4) * CM_VSTRIDE + gid + 16 * mt], ah[mt][2], al[mt][2]);
    tf32_split(vis_w[(tig + 4) * CM_VSTRIDE + gid + 8 + 16 * mt], ah[mt][3], al[mt][3]);
  }
  const float* c0 = col[slot_w[tig]];
  const float* c1 = col[slot_w[tig + 4]];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const int ch = min(gid + 8 * nt, DP - 1);   // columns past DP are never written out
    unsigned b0h, b0l, b1h, b1l;
    tf32_split(c0[ch], b0h, b0l);
    tf32_split(c1[ch], b1h, b1l);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      mma_tf32_16x8x8(cf[mt][nt], al[mt], b0h, b1h);
      mma_tf32_16x8x8(cf[mt][nt], ah[mt], b0l, b1l);
      mma_tf32_16x8x8(cf[mt][nt], ah[mt], b0h, b1h);
    }
  }
}

template <int DP, int BATCH>
__global__ void __launch_bounds__(RT, 2) rasterize_fwd_mma_kernel(const RasterArgs a) {
  constexpr int NT = (DP + 7) / 8;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem<DP, BATCH>& s = *reinterpret_cast<Smem<DP, BATCH>*>(smem_raw);
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  float* cm_vis = reinterpret_cast<float*>(smem_raw + sizeof(Smem<DP, BATCH>)) + warp * (CM_PEND * CM_VSTRIDE);
  int* cm_slot = reinterpret_cast<int*>(reinterpret_cast<float*>(smem_raw + sizeof(Smem<DP, BATCH>)) +
                                        (RT / 32) * (CM_PEND * CM_VSTRIDE)) + warp * CM_PEND;
  const TileCtx c = tile_ctx(a, lane, warp);
  const int start = c.start, end = c.end;
  const float px = c.px, py = c.py;

  float T = c.inside ? 1.f : 0.f, T_out = 1.f, dsum = 0.f, tmed = 0.f, nx = 0.f, ny = 0.f, nz = 0.f;
  float cf[2][NT][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) cf[mt][nt][0] = cf[mt][nt][1] = cf[mt][nt][2] = cf[mt][nt][3] = 0.f;
  int last_id = start - 1, med_id = -1, cm_n = 0;

  const int nb = (end - start + BATCH - 1) / BATCH;
  if (nb > 0) {
    if (t < BATCH) { const int i = start + t; s.ids[0][t] = i < end ? __ldg(a.flatten_ids + i) : 0; }
    __syncthreads();
    issue_gather<DP, BATCH>(s, 0, min(BATCH, end - start), a, t);
    if (nb > 1 && t < BATCH) { const int i = start + BATCH + t; s.ids[1][t] = i < end ? __ldg(a.flatten_ids + i) : 0; }
  }
  bool warp_done = !__any_sync(RS_FULL_MASK, T != 0.f);
  for (int b = 0; b < nb; ++b) {
    rs::cp_async_wait_all();
    if (__syncthreads_count(T != 0.f) == 0) break;
    int next_id = 0;
    if (b + 1 < nb) {
      issue_gather<DP, BATCH>(s, (b + 1) & 1, min(BATCH, end - (start + (b + 1) * BATCH)), a, t);
      if (b + 2 < nb && t < BATCH) {
        const int i = start + (b + 2) * BATCH + t;
        next_id = i < end ? __ldg(a.flatten_ids + i) : 0;
      }
    }
    if (!warp_done) {
      const int buf = b & 1;
      const int base_idx = start + b * BATCH;
      const int bcount = min(BATCH, end - base_idx);
      for (int g0 = 0; g0 < bcount; g0 += 32) {
        const int j = g0 + lane;
        bool hit = false;
        if (j < bcount) hit = footprint_hit(s, buf, j, a.exact_cull != 0, c.rcx, c.rcy, 3.5f, 1.5f);
        unsigned m = __ballot_sync(RS_FULL_MASK, hit);
        while (m) {
          const int jj = g0 + __ffs(m) - 1;
          m &= m - 1;
          const float4 q0 = s.q0[buf][jj], q1 = s.q1[buf][jj];
          const float dx = q0.x - px, dy = q0.y - py;
          const float sig = q1.x * dx * dx + q1.z * dy * dy + q1.y * dx * dy;
          const float alpha = fminf(RS_ALPHA_MAX, q1.w * rs::fast_exp2(-sig));
          const bool ok = sig >= 0.f && alpha >= RS_ALPHA_MIN;
          if (!__any_sync(RS_FULL_MASK, ok)) continue;
          // branch-free: a failed pair acts as alpha = 0, a dead pixel (T == 0) discards its update
          const float am = ok ? alpha : 0.f;
          const float nT = T * (1.f - am);
          const bool live = nT > RS_T_STOP;              // false for dead pixels and for the terminating pair
          T_out = (!live && T != 0.f) ? T : T_out;
          const float vis = live ? am * T : 0.f;
          if (__any_sync(RS_FULL_MASK, vis != 0.f)) {
            const float4 q2 = s.q2[buf][jj], q3 = s.q3[buf][jj];
            const float tt = q2.x + q2.y * dx + q2.z * dy;
            dsum += vis * tt;
            nx += vis * q3.x; ny += vis * q3.y; nz += vis * q3.z;
#if RS_MEDIAN_INCLUSIVE
            const bool med = live && T > 0.5f && nT <= 0.5f;
#else
            const bool med = live && T > 0.5f && nT < 0.5f;
#endif
            tmed = med ? tt : tmed;
            med_id = med ? base_idx + jj : med_id;
            last_id = (live && ok) ? base_idx + jj : last_id;
            cm_vis[cm_n * CM_VSTRIDE + lane] = vis;
            if (lane == 0) cm_slot[cm_n] = jj;
            if (++cm_n == CM_PEND) {
              __syncwarp();
              flush_blend_mma<DP>(cf, cm_vis, cm_slot, s.col[buf], lane);
              __syncwarp();
              cm_n = 0;
            }
          }
          T = live ? nT : 0.f;
        }
        if (!__any_sync(RS_FULL_MASK, T != 0.f)) { warp_done = true; break; }
      }
      if (cm_n > 0) {   // the staged colour rows of this batch are about to be recycled: contract what is parked
        for (int g = cm_n; g < CM_PEND; ++g) {
          cm_vis[g * CM_VSTRIDE + lane] = 0.f;
          if (lane == 0) cm_slot[g] = 0;
        }
        __syncwarp();
        flush_blend_mma<DP>(cf, cm_vis, cm_slot, s.col[buf], lane);
        __syncwarp();
        cm_n = 0;
      }
    }
    if (b + 2 < nb && t < BATCH) s.ids[b & 1][t] = next_id;
  }
  rs::cp_async_wait_all();

  if (T == 0.f) T = T_out;   // (outside pixels: T_out = 1)
  // colours: the C fragments hold, per lane, pixels gid + 8h (h = 0..3) of the warp's 8x4 block and channels
  // 2 tig + 8 nt (+1); the per-pixel transmittance comes from the lane that owns the pixel
  {
    const int gid = lane >> 2, tig = lane & 3;
    const float* bg = a.backgrounds ? a.backgrounds + (size_t)c.cam * a.D : nullptr;
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const int mp = gid + 8 * h;
      const float Tm = __shfl_sync(RS_FULL_MASK, T, mp);
      const int pxm = c.x0 + (mp & 7), pym = c.y0 + (mp >> 3);
      if (pxm < a.W && pym < a.H) {
        float* oc = a.out_colors + (((size_t)c.cam * a.H + pym) * a.W + pxm) * a.D;
        const float ed_scale = 1.f / fmaxf(1.f - Tm, 1e-10f);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int ch = 2 * tig + 8 * nt + e;
            if (ch < a.D) {
              float v = cf[h >> 1][nt][(h & 1) * 2 + e] + (bg ? Tm * __ldg(bg + ch) : 0.f);
              v *= (ch == a.ed_channel) ? ed_scale : 1.f;
              oc[ch] = v;
            }
          }
        }
      }
    }
  }
  if (c.inside) {
    const size_t pix = ((size_t)c.cam * a.H + c.pyi) * a.W + c.pxi;
    const float il = inv_ray_len(a, c.cam, px, py);
    a.out_alphas[pix] = 1.f - T;
    a.out_T[pix] = T;
#if RS_NORMALIZE_EXPECTED_DEPTH
    a.out_dexp[pix] = dsum * il / fmaxf(1.f - T, 1e-10f);
#else
    a.out_dexp[pix] = dsum * il;
#endif
    a.out_dmed[pix] = tmed * il;
    a.out_normals[pix * 3] = nx; a.out_normals[pix * 3 + 1] = ny; a.out_normals[pix * 3 + 2] = nz;
    a.last_ids[pix] = last_id;
    a.median_ids[pix] = med_id;
  }
}

// ------------------------------------------------------------------------------------------------ forward
template <int DP, int BATCH, bool STATS>
__global__ void __launch_bounds__(RT) rasterize_fwd_kernel(const RasterArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem<DP, BATCH>& s = *reinterpret_cast<Smem<DP, BATCH>*>(smem_raw);
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const TileCtx c = tile_ctx(a, lane, warp);
  const int start = c.start, end = c.end;
  const float px = c.px, py = c.py;

  // T is the live transmittance and drops to exactly 0 once the pixel has terminated (T_out keeps the value
  // it had), so "done" needs no flag: a dead pixel fails the T*(1-alpha) > T_STOP test by itself.
  float T = c.inside ? 1.f : 0.f, T_out = 1.f, dsum = 0.f, tmed = 0.f, nx = 0.f, ny = 0.f, nz = 0.f;
  float acc[DP];
#pragma unroll
  for (int k = 0; k < DP; ++k) acc[k] = 0.f;
  int last_id = start - 1, med_id = -1;
  int st_contrib = 0, st_term = -1, st_evals = 0, st_blend = 0;  // STATS only

  const int nb = (end - start + BATCH - 1) / BATCH;
  if (nb > 0) {
    if (t < BATCH) { const int i = start + t; s.ids[0][t] = i < end ? __ldg(a.flatten_ids + i) : 0; }
    __syncthreads();
    issue_gather<DP, BATCH>(s, 0, min(BATCH, end - start), a, t);
    if (nb > 1 && t < BATCH) { const int i = start + BATCH + t; s.ids[1][t] = i < end ? __ldg(a.flatten_ids + i) : 0; }
  }
  bool warp_done = !__any_sync(RS_FULL_MASK, T != 0.f);
  for (int b = 0; b < nb; ++b) {
    rs::cp_async_wait_all();
    // barrier: batch b has landed, ids[(b+1)&1] are visible, every warp has finished batch b-1
    if (__syncthreads_count(T != 0.f) == 0) break;
    int next_id = 0;
    if (b + 1 < nb) {
      issue_gather<DP, BATCH>(s, (b + 1) & 1, min(BATCH, end - (start + (b + 1) * BATCH)), a, t);
      if (b + 2 < nb && t < BATCH) {
        const int i = start + (b + 2) * BATCH + t;
        next_id = i < end ? __ldg(a.flatten_ids + i) : 0;
      }
    }
    if (!warp_done) {
      const int buf = b & 1;
      const int base_idx = start + b * BATCH;
      const int bcount = min(BATCH, end - base_idx);
      for (int g0 = 0; g0 < bcount; g0 += 32) {
        const int j = g0 + lane;
        bool hit = false;
        if (j < bcount) {
          hit = footprint_hit(s, buf, j, a.exact_cull != 0, c.rcx, c.rcy, 3.5f, 1.5f);
        }
        unsigned m = __ballot_sync(RS_FULL_MASK, hit);
        while (m) {
          const int jj = g0 + __ffs(m) - 1;
          m &= m - 1;
          const float4 q0 = s.q0[buf][jj], q1 = s.q1[buf][jj];
          const float dx = q0.x - px, dy = q0.y - py;
          const float sig = q1.x * dx * dx + q1.z * dy * dy + q1.y * dx * dy;
          const float alpha = fminf(RS_ALPHA_MAX, q1.w * rs::fast_exp2(-sig));
          if constexpr (STATS) {
            ++st_evals;
            st_blend += __any_sync(RS_FULL_MASK, sig >= 0.f && alpha >= RS_ALPHA_MIN && T * (1.f - alpha) > RS_T_STOP);
          }
          if (sig >= 0.f && alpha >= RS_ALPHA_MIN) {
            const float nT = T * (1.f - alpha);
            if (!(nT > RS_T_STOP)) {
              if (T != 0.f) { T_out = T; T = 0.f; if constexpr (STATS) st_term = base_idx + jj; }  // terminate
            } else {
              if constexpr (STATS) ++st_contrib;
              const float vis = alpha * T;
              const float4 q2 = s.q2[buf][jj], q3 = s.q3[buf][jj];
              const float tt = q2.x + q2.y * dx + q2.z * dy;
              dsum += vis * tt;
              nx += vis * q3.x; ny += vis * q3.y; nz += vis * q3.z;
              const float4* cp = reinterpret_cast<const float4*>(&s.col[buf][jj][0]);
#pragma unroll
              for (int k = 0; k < DP / 4; ++k) {
                const float4 cc = cp[k];
                acc[4 * k] += vis * cc.x; acc[4 * k + 1] += vis * cc.y;
                acc[4 * k + 2] += vis * cc.z; acc[4 * k + 3] += vis * cc.w;
              }
#if RS_MEDIAN_INCLUSIVE
              if (T > 0.5f && nT <= 0.5f) { tmed = tt; med_id = base_idx + jj; }
#else
              if (T > 0.5f && nT < 0.5f) { tmed = tt; med_id = base_idx + jj; }
#endif
              last_id = base_idx + jj;
              T = nT;
            }
          }
        }
        if (!__any_sync(RS_FULL_MASK, T != 0.f)) { warp_done = true; break; }  // all 32 pixels saturated
      }
    }
    if (b + 2 < nb && t < BATCH) s.ids[b & 1][t] = next_id;  // batch b's ids are dead (gather issued last iteration)
  }
  rs::cp_async_wait_all();

  if constexpr (STATS) {
    // Q: list entries a per-pixel loop (the reference's) visits = up to and including the terminating Gaussian,
    // or the whole list if the pixel never saturates;  Qc: pairs that were blended
    unsigned long long q = 0, qc = 0;
    if (c.inside) { q = (unsigned long long)((st_term >= 0 ? st_term + 1 : end) - start); qc = (unsigned long long)st_contrib; }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
      q += __shfl_xor_sync(RS_FULL_MASK, q, d);
      qc += __shfl_xor_sync(RS_FULL_MASK, qc, d);
    }
    if (lane == 0) {
      atomicAdd(a.stats + 0, q); atomicAdd(a.stats + 1, qc);
      atomicAdd(a.stats + 2, (unsigned long long)st_evals); atomicAdd(a.stats + 3, (unsigned long long)st_blend);
    }
  }

  if (c.inside) {
    if (T == 0.f) T = T_out;  // terminated pixel: transmittance in front of the Gaussian that stopped it
    const size_t pix = ((size_t)c.cam * a.H + c.pyi) * a.W + c.pxi;
    const float il = inv_ray_len(a, c.cam, px, py);
    float* oc = a.out_colors + pix * a.D;
    const float* bg = a.backgrounds ? a.backgrounds + (size_t)c.cam * a.D : nullptr;
    const float ed_scale = 1.f / fmaxf(1.f - T, 1e-10f);
#pragma unroll
    for (int k = 0; k < DP; ++k)
      if (k < a.D) {
        float v = acc[k] + (bg ? T * __ldg(bg + k) : 0.f);
        v *= (k == a.ed_channel) ? ed_scale : 1.f;  // expected depth: accumulated depth / alpha
        oc[k] = v;
      }
    a.out_alphas[pix] = 1.f - T;
    a.out_T[pix] = T;
#if RS_NORMALIZE_EXPECTED_DEPTH
    a.out_dexp[pix] = dsum * il / fmaxf(1.f - T, 1e-10f);
#else
    a.out_dexp[pix] = dsum * il;
#endif
    a.out_dmed[pix] = tmed * il;
    a.out_normals[pix * 3] = nx; a.out_normals[pix * 3 + 1] = ny; a.out_normals[pix * 3 + 2] = nz;
    a.last_ids[pix] = last_id;
    a.median_ids[pix] = med_id;
  }
}

// sigma' of the two-pixels-per-lane kernels: the x-dependent terms (adx2 = a'dx^2, bdx = b'dx) are shared by the
// two pixels of a column.  Every operation is individually rounded / explicitly fused, so the forward and the
// backward kernel take the same alpha >= 1/255 decisions.
__device__ __forceinline__ float sigma_col(float adx2, float bdx, float c, float dy) {
  return fmaf(fmaf(c, dy, bdx), dy, adx2);
}

// ------------------------------------------------------------------------------------------------ forward, 2 px/lane
// Variant of the forward kernel in which a warp owns an 8x8 pixel block and every lane blends TWO pixels of the
// same column, (x, y) and (x, y+4).  The Gaussian record loads, the loop control and the x-dependent part of
// the quadratic form (dx, a'dx^2, b'dx) are shared by the two pixels, which cuts the per-pixel instruction
// count of the (issue-bound) alpha test; 4 warps (128 threads) per tile, 128-Gaussian batches.
constexpr int RT2 = 128;

template <int DP, int BATCH, bool STATS, int NW, int S = 2, bool EDEV = false>
// (S == 2, the default barrier staging: capped at 72 registers = 7 CTAs of 128 threads per SM; the allocation drifts to
// 80 registers / 6 CTAs otherwise, which costs the forward 5 %)
__global__ void __launch_bounds__(NW * 32, (S == 2 && NW == 4) ? 0 : 6) rasterize_fwd2_kernel(const RasterArgs a) {
  constexpr int NT = NW * 32, SUB = 4 / NW;   // NW warps per CTA: a CTA covers NW of the tile's four 8x8 blocks
  static_assert(S == 2 || BATCH == NT, "ring: one slot per thread");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem<DP, BATCH, S>& s = *reinterpret_cast<Smem<DP, BATCH, S>*>(smem_raw);
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int tile_id = blockIdx.x / SUB;
  const int qd = (blockIdx.x - tile_id * SUB) * NW + warp;   // which 8x8 block of the tile this warp owns
  const int tiles_per_cam = a.tile_w * a.tile_h;
  const int cam = tile_id / tiles_per_cam;
  const int tl = tile_id - cam * tiles_per_cam;
  const int tyi = tl / a.tile_w, txi = tl - tyi * a.tile_w;
  const int start = __ldg(a.offsets + tile_id);
  const int end = list_end_of<EDEV>(a, tile_id, a.C * tiles_per_cam);
  const int x0 = txi * RS_TILE + (qd & 1) * 8, y0 = tyi * RS_TILE + (qd >> 1) * 8;
  const int pxi = x0 + (lane & 7);
  const int pyi[2] = {y0 + (lane >> 3), y0 + (lane >> 3) + 4};
  const bool inside[2] = {pxi < a.W && pyi[0] < a.H, pxi < a.W && pyi[1] < a.H};
  const float px = pxi + 0.5f;
  const float py[2] = {pyi[0] + 0.5f, pyi[1] + 0.5f};
  const float rcx = x0 + 4.0f, rcy = y0 + 4.0f;

  float T[2] = {inside[0] ? 1.f : 0.f, inside[1] ? 1.f : 0.f}, T_out[2] = {1.f, 1.f};
  float dsum[2] = {0.f, 0.f}, tmed[2] = {0.f, 0.f}, nx[2] = {0.f, 0.f}, ny[2] = {0.f, 0.f}, nz[2] = {0.f, 0.f};
  float acc[2][DP];
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int k = 0; k < DP; ++k) acc[h][k] = 0.f;
  int last_id[2] = {start - 1, start - 1}, med_id[2] = {-1, -1};
  int st_contrib[2] = {0, 0}, st_term[2] = {-1, -1}, st_evals = 0, st_blend = 0;  // STATS only

  const int nb = (end - start + BATCH - 1) / BATCH;
  bool warp_done = !__any_sync(RS_FULL_MASK, T[0] != 0.f || T[1] != 0.f);
  int ring_id = 0;   // ring: flatten id of this thread's slot in the next batch to gather
  if constexpr (S == 2) {
    if (nb > 0) {
      for (int i = t; i < BATCH; i += NT) { const int g = start + i; s.ids[0][i] = g < end ? __ldg(a.flatten_ids + g) : 0; }
      __syncthreads();
      issue_gather<DP, BATCH, NT>(s, 0, min(BATCH, end - start), a, t);
      if (nb > 1)
        for (int i = t; i < BATCH; i += NT) { const int g = start + BATCH + i; s.ids[1][i] = g < end ? __ldg(a.flatten_ids + g) : 0; }
    }
  } else {
    if (t == 0) {
#pragma unroll
      for (int i = 0; i < S; ++i) { mbar_init(&s.full[i], NT); mbar_init(&s.empty[i], NW); }
      s.done = 0;
    }
    __syncthreads();
    if (warp_done && lane == 0) atomicAdd(&s.done, 1);
    if (nb > 0) {
      ring_id = start + t < end ? __ldg(a.flatten_ids + start + t) : 0;
      ring_gather_share(s, 0, min(BATCH, end - start), ring_id, a, lane, warp);
      ring_id = start + BATCH + t < end ? __ldg(a.flatten_ids + start + BATCH + t) : 0;
    }
  }
  for (int b = 0; b < nb; ++b) {
    int next_id[BATCH / NT];
    if constexpr (S == 2) {
      rs::cp_async_wait_all();
      if (__syncthreads_count(T[0] != 0.f || T[1] != 0.f) == 0) break;
      if (b + 1 < nb) {
        issue_gather<DP, BATCH, NT>(s, (b + 1) & 1, min(BATCH, end - (start + (b + 1) * BATCH)), a, t);
        if (b + 2 < nb) {
#pragma unroll
          for (int r = 0; r < BATCH / NT; ++r) {
            const int g = start + (b + 2) * BATCH + r * NT + t;
            next_id[r] = g < end ? __ldg(a.flatten_ids + g) : 0;
          }
        }
      }
    } else {
      const volatile int* done = &s.done;
      if (*done >= NW) break;                                  // every warp of the tile is saturated
      if (b + 1 < nb) {                                        // refill the stage batch b + 1 - S lived in
        const int st1 = (b + 1) % S;
        if (b + 1 >= S && !mbar_wait_or_done(&s.empty[st1], (unsigned)(((b + 1 - S) / S) & 1), done, NW)) break;
        ring_gather_share(s, st1, min(BATCH, end - (start + (b + 1) * BATCH)), ring_id, a, lane, warp);
        const int g = start + (b + 2) * BATCH + t;
        ring_id = g < end ? __ldg(a.flatten_ids + g) : 0;
      }
      if (!mbar_wait_or_done(&s.full[b % S], (unsigned)((b / S) & 1), done, NW)) break;
    }
    if (!warp_done) {
      const int buf = S == 2 ? (b & 1) : (b % S);
      const int base_idx = start + b * BATCH;
      const int bcount = min(BATCH, end - base_idx);
      for (int g0 = 0; g0 < bcount; g0 += 32) {
        const int j = g0 + lane;
        bool hit = false;
        if (j < bcount) {
          hit = footprint_hit(s, buf, j, a.exact_cull != 0, rcx, rcy, 3.5f, 3.5f);
        }
        unsigned m = __ballot_sync(RS_FULL_MASK, hit);
        // survivors are taken two at a time: their alpha evaluations are independent (only the transmittance chain
        // is sequential), which gives the scheduler two dependency chains to interleave
        auto alpha_of = [&](int jj, float& dx, float (&dy)[2], float (&alpha)[2], bool (&ok)[2]) {
          const float4 q0 = s.q0[buf][jj], q1 = s.q1[buf][jj];
          dx = q0.x - px;
          const float A = __fmul_rn(__fmul_rn(q1.x, dx), dx), B = __fmul_rn(q1.y, dx);   // shared by both pixels of the column
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            dy[h] = q0.y - py[h];
            const float sig = sigma_col(A, B, q1.z, dy[h]);
            alpha[h] = fminf(RS_ALPHA_MAX, q1.w * rs::fast_exp2(-sig));
            ok[h] = sig >= 0.f && alpha[h] >= RS_ALPHA_MIN;
          }
        };
        auto blend = [&](int jj, float dx, const float (&dy)[2], const float (&alpha)[2], const bool (&ok)[2]) {
          if constexpr (STATS) {
            ++st_evals;
            st_blend += __any_sync(RS_FULL_MASK, (ok[0] && T[0] * (1.f - alpha[0]) > RS_T_STOP) ||
                                                     (ok[1] && T[1] * (1.f - alpha[1]) > RS_T_STOP));
          }
          if (ok[0] || ok[1]) {
            const float4 q2 = s.q2[buf][jj], q3 = s.q3[buf][jj];
            const float4* cp = reinterpret_cast<const float4*>(&s.col[buf][jj][0]);
            const float tb = q2.x + q2.y * dx;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              if (ok[h]) {
                const float nT = T[h] * (1.f - alpha[h]);
                if (!(nT > RS_T_STOP)) {
                  if (T[h] != 0.f) { T_out[h] = T[h]; T[h] = 0.f; if constexpr (STATS) st_term[h] = base_idx + jj; }
                } else {
                  if constexpr (STATS) ++st_contrib[h];
                  const float vis = alpha[h] * T[h];
                  const float tt = tb + q2.z * dy[h];
                  dsum[h] += vis * tt;
                  nx[h] += vis * q3.x; ny[h] += vis * q3.y; nz[h] += vis * q3.z;
#pragma unroll
                  for (int k = 0; k < DP / 4; ++k) {
                    const float4 cc = cp[k];
                    acc[h][4 * k] += vis * cc.x; acc[h][4 * k + 1] += vis * cc.y;
                    acc[h][4 * k + 2] += vis * cc.z; acc[h][4 * k + 3] += vis * cc.w;
                  }
#if RS_MEDIAN_INCLUSIVE
                  if (T[h] > 0.5f && nT <= 0.5f) { tmed[h] = tt; med_id[h] = base_idx + jj; }
#else
                  if (T[h] > 0.5f && nT < 0.5f) { tmed[h] = tt; med_id[h] = base_idx + jj; }
#endif
                  last_id[h] = base_idx + jj;
                  T[h] = nT;
                }
              }
            }
          }
        };
        while (m) {
          const int ja = g0 + __ffs(m) - 1;
          m &= m - 1;
          const bool two = m != 0;
          const int jb = two ? g0 + __ffs(m) - 1 : ja;
          m &= m - 1;                                  // (0 & anything stays 0)
          float dxa, dxb, dya[2], dyb[2], ala[2], alb[2];
          bool oka[2], okb[2];
          alpha_of(ja, dxa, dya, ala, oka);
          alpha_of(jb, dxb, dyb, alb, okb);
          blend(ja, dxa, dya, ala, oka);
          if (two) blend(jb, dxb, dyb, alb, okb);
        }
        if (!__any_sync(RS_FULL_MASK, T[0] != 0.f || T[1] != 0.f)) {
          warp_done = true;
          if constexpr (S > 2) { if (lane == 0) atomicAdd(&s.done, 1); }
          break;
        }
      }
    }
    if constexpr (S == 2) {
      if (b + 2 < nb) {
#pragma unroll
        for (int r = 0; r < BATCH / NT; ++r) s.ids[b & 1][r * NT + t] = next_id[r];
      }
    } else {
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.empty[b % S]);             // this warp is through with the stage
    }
  }
  rs::cp_async_wait_all();

  if constexpr (STATS) {   // same counters as the one-pixel kernel; evaluations are per (8x8 warp block, Gaussian)
    unsigned long long q = 0, qc = 0;
#pragma unroll
    for (int h = 0; h < 2; ++h)
      if (inside[h]) {
        q += (unsigned long long)((st_term[h] >= 0 ? st_term[h] + 1 : end) - start);
        qc += (unsigned long long)st_contrib[h];
      }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
      q += __shfl_xor_sync(RS_FULL_MASK, q, d);
      qc += __shfl_xor_sync(RS_FULL_MASK, qc, d);
    }
    if (lane == 0) {
      atomicAdd(a.stats + 0, q); atomicAdd(a.stats + 1, qc);
      atomicAdd(a.stats + 2, (unsigned long long)st_evals); atomicAdd(a.stats + 3, (unsigned long long)st_blend);
    }
  }

#pragma unroll
  for (int h = 0; h < 2; ++h) {
    if (!inside[h]) continue;
    float Tf = T[h] == 0.f ? T_out[h] : T[h];
    const size_t pix = ((size_t)cam * a.H + pyi[h]) * a.W + pxi;
    const float il = inv_ray_len(a, cam, px, py[h]);
    float* oc = a.out_colors + pix * a.D;
    const float* bg = a.backgrounds ? a.backgrounds + (size_t)cam * a.D : nullptr;
    const float ed_scale = 1.f / fmaxf(1.f - Tf, 1e-10f);
#pragma unroll
    for (int k = 0; k < DP; ++k)
      if (k < a.D) {
        float v = acc[h][k] + (bg ? Tf * __ldg(bg + k) : 0.f);
        v *= (k == a.ed_channel) ? ed_scale : 1.f;
        oc[k] = v;
      }
    a.out_alphas[pix] = 1.f - Tf;
    a.out_T[pix] = Tf;
    a.out_dexp[pix] = dsum[h] * il;
    a.out_dmed[pix] = tmed[h] * il;
    a.out_normals[pix * 3] = nx[h]; a.out_normals[pix * 3 + 1] = ny[h]; a.out_normals[pix * 3 + 2] = nz[h];
    a.last_ids[pix] = last_id[h];
    a.median_ids[pix] = med_id[h];
  }
}

// ------------------------------------------------------------------------------------------------ backward
// reduce-and-commit the DP colour gradients in power-of-two chunks
template <int DP, int OFF>
__device__ __forceinline__ void commit_color_grads(const float (&v_c)[DP], float vis, float* __restrict__ dst, int D,
                                                   int lane) {
  if constexpr (OFF < DP) {
    constexpr int REM = DP - OFF;
    constexpr int K = REM >= 32 ? 32 : REM >= 16 ? 16 : REM >= 8 ? 8 : 4;
    float cv[K];
#pragma unroll
    for (int k = 0; k < K; ++k) cv[k] = vis * v_c[OFF + k];
    rs::warp_reduce_scatter<K>(cv, lane);
    constexpr int GROUP = 32 / K;
    const int slot = OFF + lane / GROUP;
    if ((lane % GROUP) == 0 && slot < D) atomicAdd(dst + slot, cv[0]);
    commit_color_grads<DP, OFF + K>(v_c, vis, dst, D, lane);
  }
}

// ---- tensor-core reduction of the wide colour gradients (rade-features rows, DP >= 32)
// The per-Gaussian colour gradient is a small dense contraction over the warp's 32 pixels,
//     v_col[g][ch] = sum_p vis[p][g] * v_c[p][ch],
// the one piece of the compositing backward that is GEMM-shaped: M = channels (16-row tiles), N = 8 pending
// Gaussians, K = 32 pixels.  A warp parks the visibilities of the Gaussians it blends (8 x 32 floats) and every 8th
// one contracts them against its v_c rows (already in shared memory) with mma.sync m16n8k8 TF32 instructions.  Both
// operands are split into a TF32 head and a TF32 tail (x = hi + lo) and the three significant partial products are
// accumulated in fp32 ("3xTF32"), so the result keeps fp32 accuracy (the dropped lo*lo term is ~2^-22 relative).
// The C fragments land as (channel, Gaussian) pairs and are committed with one RED each, 8 consecutive channels per
// quarter-warp.  Replaces ~11 issue slots per contributing pixel and Gaussian of the shuffle/row-walk formulation.
// vc_w: the warp's v_c rows [32][DP]; vis_w: parked visibilities [CM_PEND][CM_VSTRIDE]; row_w: colour rows [CM_PEND]
template <int DP>
__device__ __forceinline__ void flush_color_mma(const float* __restrict__ vc_w, const float* __restrict__ vis_w,
                                                const int* __restrict__ row_w, int np, float* __restrict__ color_grad,
                                                int D, int lane) {
  constexpr int MT = (DP + 15) / 16;
  const int gid = lane >> 2, tig = lane & 3;
  float c[MT][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) c[mt][0] = c[mt][1] = c[mt][2] = c[mt][3] = 0.f;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    unsigned b0h, b0l, b1h, b1l;
    tf32_split(vis_w[gid * CM_VSTRIDE + tig + 8 * ks], b0h, b0l);
    tf32_split(vis_w[gid * CM_VSTRIDE + tig + 4 + 8 * ks], b1h, b1l);
    const float* r0 = vc_w + (tig + 8 * ks) * DP;
    const float* r1 = r0 + 4 * DP;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      // rows past DP only feed C rows that are never committed; clamp them into the row to stay inside the array
      const int ch0 = min(gid + 16 * mt, DP - 1), ch1 = min(gid + 8 + 16 * mt, DP - 1);
      unsigned ah[4], al[4];
      tf32_split(r0[ch0], ah[0], al[0]);
      tf32_split(r0[ch1], ah[1], al[1]);
      tf32_split(r1[ch0], ah[2], al[2]);
      tf32_split(r1[ch1], ah[3], al[3]);
      mma_tf32_16x8x8(c[mt], al, b0h, b1h);
      mma_tf32_16x8x8(c[mt], ah, b0l, b1l);
      mma_tf32_16x8x8(c[mt], ah, b0h, b1h);
    }
  }
  const int g0 = 2 * tig, g1 = g0 + 1;
  float* d0 = color_grad + (size_t)row_w[g0] * DP;
  float* d1 = color_grad + (size_t)row_w[g1] * DP;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    const int ch = gid + 16 * mt;
    if (ch < D) {
      if (g0 < np) atomicAdd(d0 + ch, c[mt][0]);
      if (g1 < np) atomicAdd(d1 + ch, c[mt][1]);
    }
    if (ch + 8 < D) {
      if (g0 < np) atomicAdd(d0 + ch + 8, c[mt][2]);
      if (g1 < np) atomicAdd(d1 + ch + 8, c[mt][3]);
    }
  }
}

// The median depth of a pixel is the ray distance t = ray_t + ray_plane . d of ONE Gaussian (median_ids): its
// gradient is three scalar atomics per pixel, issued once here instead of a compare + select per blended pair in
// the compositing loop.
__device__ __forceinline__ void commit_median_grad(const RasterArgs& a, int med_id, float v_dmed, float px, float py) {
  if (med_id < 0 || v_dmed == 0.f) return;
  const int gid = __ldg(a.flatten_ids + med_id);
  const float4 q0 = __ldg(a.geom + (size_t)gid * 4);
  float* rec = a.geom_grad + (size_t)gid * 16;
  atomicAdd(rec + 6, v_dmed);
  atomicAdd(rec + 7, v_dmed * (q0.x - px));
  atomicAdd(rec + 8, v_dmed * (q0.y - py));
}

template <int DP, int BATCH, bool ABSGRAD, bool CMMA>
__global__ void __launch_bounds__(RT, (CMMA ? 2 : 0)) rasterize_bwd_kernel(const RasterArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem<DP, BATCH>& s = *reinterpret_cast<Smem<DP, BATCH>*>(smem_raw);
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const TileCtx c = tile_ctx(a, lane, warp);
  const int start = c.start, end = c.end;
  const float px = c.px, py = c.py;
  const bool inside = c.inside;
  const size_t pix = inside ? ((size_t)c.cam * a.H + c.pyi) * a.W + c.pxi : 0;

  float v_c[DP];
#pragma unroll
  for (int k = 0; k < DP; ++k) v_c[k] = (inside && k < a.D) ? __ldg(a.v_colors + pix * a.D + k) : 0.f;
  const float T_final = inside ? a.out_T[pix] : 1.f;
  // "ED" channel: out = S / max(alpha, 1e-10)  ->  v_S = v_out / max(alpha,..),  v_alpha += -out * v_out / alpha
  float v_alpha_ed = 0.f;
  if (a.ed_channel >= 0 && inside) {   // select form: a dynamic index would push v_c into local memory
    const float alpha_out = 1.f - T_final;
    const float sc = 1.f / fmaxf(alpha_out, 1e-10f);
    const float oc = __ldg(a.out_colors + pix * a.D + a.ed_channel);
    const float vce = __ldg(a.v_colors + pix * a.D + a.ed_channel);   // re-read: v_c[ed_channel] would be a dynamic index
    if (alpha_out > 1e-10f) v_alpha_ed = -oc * vce * sc;
#pragma unroll
    for (int k = 0; k < DP; ++k) v_c[k] *= (k == a.ed_channel) ? sc : 1.f;
  }
  // Wide colour rows (>= 32 channels): the per-Gaussian colour gradient sum_p vis[p] * v_c[p][k] is formed with
  // LANES = CHANNELS (each lane owns channels lane, lane + 32, ...), walking the warp's contributing pixels: vis[p]
  // is broadcast by one shuffle, v_c[p][.] is read from shared memory (conflict free) and the totals land one channel
  // per lane, i.e. as coalesced REDs -- instead of 31 shuffle exchanges per 32 channels with lanes = pixels.
  constexpr bool CHLANE = DP >= 32;
  constexpr bool STAGED = CHLANE || CMMA;   // the tile's v_c rows are also kept in shared memory
  static_assert(!CMMA || DP >= 16, "the tensor-core colour reduction needs at least one 16-channel tile");
  float* s_vc = reinterpret_cast<float*>(smem_raw + sizeof(Smem<DP, BATCH>));   // [RT][DP], CHLANE only
  // CMMA: per-warp parking lot of the tensor-core colour reduction (flush_color_mma)
  float* cm_vis = s_vc + RT * DP + warp * (CM_PEND * CM_VSTRIDE);               // [CM_PEND][CM_VSTRIDE]
  int* cm_row = reinterpret_cast<int*>(s_vc + RT * DP + (RT / 32) * (CM_PEND * CM_VSTRIDE)) + warp * CM_PEND;
  int cm_n = 0;
  if constexpr (STAGED) {
#pragma unroll
    for (int k = 0; k < DP; ++k) s_vc[t * DP + k] = v_c[k];   // (after the ED rescale above; made visible by the barrier below)
  }
  const int last_id = inside ? a.last_ids[pix] : start - 1;
  const float il = inside ? inv_ray_len(a, c.cam, px, py) : 0.f;
#if RS_NORMALIZE_EXPECTED_DEPTH
#error "alpha-normalised expected depth (Q1) needs the extra dDexp/dalpha term in the backward"
#endif
  const float v_dsum = inside ? __ldg(a.v_dexp + pix) * il : 0.f;
  if (inside) commit_median_grad(a, a.median_ids[pix], __ldg(a.v_dmed + pix) * il, px, py);
  const float v_n0 = inside ? __ldg(a.v_normals + pix * 3) : 0.f;
  const float v_n1 = inside ? __ldg(a.v_normals + pix * 3 + 1) : 0.f;
  const float v_n2 = inside ? __ldg(a.v_normals + pix * 3 + 2) : 0.f;
  float bgdot = 0.f;
  if (a.backgrounds) {
#pragma unroll
    for (int k = 0; k < DP; ++k)
      if (k < a.D) bgdot += __ldg(a.backgrounds + (size_t)c.cam * a.D + k) * v_c[k];
  }
  const float tfin_term = inside ? T_final * (__ldg(a.v_alphas + pix) + v_alpha_ed - bgdot) : 0.f;
  float T = T_final, R = 0.f;

  // warp / CTA extent of the lists
  int warp_last = last_id;
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) warp_last = max(warp_last, __shfl_xor_sync(RS_FULL_MASK, warp_last, d));
  if (lane == 0) s.red[warp] = warp_last;
  __syncthreads();
  int blk_last = start - 1;
#pragma unroll
  for (int w = 0; w < 8; ++w) blk_last = max(blk_last, s.red[w]);
  const int nb = blk_last >= start ? (blk_last - start) / BATCH + 1 : 0;

  if (nb > 0) {
    const int bl = nb - 1;
    if (t < BATCH) { const int i = start + bl * BATCH + t; s.ids[bl & 1][t] = i < end ? __ldg(a.flatten_ids + i) : 0; }
    __syncthreads();
    issue_gather<DP, BATCH>(s, bl & 1, min(BATCH, end - (start + bl * BATCH)), a, t);
    if (bl >= 1 && t < BATCH) s.ids[(bl - 1) & 1][t] = __ldg(a.flatten_ids + start + (bl - 1) * BATCH + t);
  }
  for (int b = nb - 1; b >= 0; --b) {
    rs::cp_async_wait_all();
    __syncthreads();
    int next_id = 0;
    if (b >= 1) {
      issue_gather<DP, BATCH>(s, (b - 1) & 1, BATCH, a, t);
      if (b >= 2 && t < BATCH) next_id = __ldg(a.flatten_ids + start + (b - 2) * BATCH + t);
    }
    const int buf = b & 1;
    const int base_idx = start + b * BATCH;
    const int hi = min(min(BATCH, end - base_idx) - 1, warp_last - base_idx);
    for (int g0 = hi >= 0 ? (hi & ~31) : -32; g0 >= 0; g0 -= 32) {
      const int j = g0 + lane;
      bool hit = false;
      if (j <= hi) {
        hit = footprint_hit(s, buf, j, a.exact_cull != 0, c.rcx, c.rcy, 3.5f, 1.5f);
      }
      unsigned m = __ballot_sync(RS_FULL_MASK, hit);
      while (m) {
        const int bit = 31 - __clz(m);
        m &= ~(1u << bit);
        const int jj = g0 + bit;
        const int idx = base_idx + jj;
        const float4 q0 = s.q0[buf][jj], q1 = s.q1[buf][jj];
        const float dx = q0.x - px, dy = q0.y - py;
        const float sig = q1.x * dx * dx + q1.z * dy * dy + q1.y * dx * dy;
        const float ex = rs::fast_exp2(-sig);
        const float oe = q1.w * ex;
        const float alpha = fminf(RS_ALPHA_MAX, oe);
        const bool valid = inside && idx <= last_id && sig >= 0.f && alpha >= RS_ALPHA_MIN;
        if (!__any_sync(RS_FULL_MASK, valid)) continue;
        const float4 q2 = s.q2[buf][jj], q3 = s.q3[buf][jj];
        // Branch-free: invalid lanes behave as a Gaussian with alpha = 0 (ra = 1, vis = 0), so every gradient
        // term below is an exact 0 for them and no zero-initialisation / divergent region is needed.
        const float am = valid ? alpha : 0.f;
        const float ra = rs::fast_rcp(1.f - am);
        T *= ra;  // transmittance in front of this Gaussian
        const float vis = am * T;
        const float tt = q2.x + q2.y * dx + q2.z * dy;
        float w = v_dsum * tt + v_n0 * q3.x + v_n1 * q3.y + v_n2 * q3.z;
        {
          // <v_c, colour>: four independent partial sums -- one accumulator would be a chain of DP dependent FMAs
          // (4 cycles each), which the 2-4 resident warps per scheduler of the wide variants cannot hide
          const float4* cp = reinterpret_cast<const float4*>(&s.col[buf][jj][0]);
          float w1 = 0.f, w2 = 0.f, w3 = 0.f;
#pragma unroll
          for (int k = 0; k < DP / 4; ++k) {
            const float4 cc = cp[k];
            w = fmaf(v_c[4 * k], cc.x, w); w1 = fmaf(v_c[4 * k + 1], cc.y, w1);
            w2 = fmaf(v_c[4 * k + 2], cc.z, w2); w3 = fmaf(v_c[4 * k + 3], cc.w, w3);
          }
          w = (w + w1) + (w2 + w3);
        }
        const float v_alpha = T * w - R * ra + tfin_term * ra;
        R += vis * w;
        const float v_t = vis * v_dsum;   // (the median Gaussian's extra term: commit_median_grad)
        const bool unclamped = valid && oe <= RS_ALPHA_MAX;
        const float v_sig = unclamped ? -am * v_alpha : 0.f;
        // the record carries the moments S v_sigma*(dx, dy, 1); unpack_geom_grad_kernel turns them into the
        // means2d / opacity gradients (per-Gaussian linear maps)
        float gq[16];
        gq[0] = v_sig * dx; gq[1] = v_sig * dy;
        gq[2] = 0.5f * dx * gq[0]; gq[3] = dy * gq[0]; gq[4] = 0.5f * dy * gq[1]; gq[5] = v_sig;
        gq[6] = v_t; gq[7] = v_t * dx; gq[8] = v_t * dy;
        gq[9] = vis * v_n0; gq[10] = vis * v_n1; gq[11] = vis * v_n2;
        if constexpr (DP == 4) {
          gq[12] = vis * v_c[0]; gq[13] = vis * v_c[1]; gq[14] = vis * v_c[2]; gq[15] = vis * v_c[3];
        } else {
          gq[12] = gq[13] = gq[14] = gq[15] = 0.f;
        }
        float ax = 0.f, ay = 0.f;
        if constexpr (ABSGRAD) {   // |d L / d means2d| summed per pixel: needs the per-pixel gradient itself
          const float vs2 = v_sig * RS_LN2;
          ax = fabsf(vs2 * (2.f * q1.x * dx + q1.y * dy) + v_t * q2.y);
          ay = fabsf(vs2 * (q1.y * dx + 2.f * q1.z * dy) + v_t * q2.z);
        }
        const int id = __float_as_int(q2.w);  // flatten id carried by the record (s.ids is recycled concurrently)
        rs::warp_reduce_scatter<16>(gq, lane);
        {
          const int slot = lane >> 1;
          if ((lane & 1) == 0 && (DP == 4 || slot < 12)) atomicAdd(a.geom_grad + (size_t)id * 16 + slot, gq[0]);
        }
        if constexpr (ABSGRAD) {
          float ab[2] = {ax, ay};
          rs::warp_reduce_scatter<2>(ab, lane);
          if ((lane & 15) == 0) atomicAdd(a.abs_grad + (size_t)id * 2 + (lane >> 4), ab[0]);
        }
        if constexpr (DP > 4) {
          const int row = a.color_per_cam ? id : id % a.N;
          if constexpr (CMMA) {
            cm_vis[cm_n * CM_VSTRIDE + lane] = vis;
            if (lane == 0) cm_row[cm_n] = row;
            if (++cm_n == CM_PEND) {
              __syncwarp();
              flush_color_mma<DP>(s_vc + (size_t)(warp * 32) * DP, cm_vis, cm_row, CM_PEND, a.color_grad, a.D, lane);
              __syncwarp();
              cm_n = 0;
            }
          } else if constexpr (CHLANE) {
            constexpr int NCH = (DP + 31) / 32;
            float acc[NCH];
#pragma unroll
            for (int cch = 0; cch < NCH; ++cch) acc[cch] = 0.f;
            unsigned nz = __ballot_sync(RS_FULL_MASK, vis != 0.f);
            const float* rows = s_vc + (size_t)(warp * 32) * DP + lane;
            constexpr int UN = 4;   // contributing pixels per round (8 measured slower, 1 latency-bound): their shuffles and row loads are independent
            while (nz) {
              int pl[UN];
              float vp[UN];
#pragma unroll
              for (int u = 0; u < UN; ++u) {
                pl[u] = nz ? __ffs(nz) - 1 : 0;
                const bool on = nz != 0;
                nz &= nz - 1;
                vp[u] = __shfl_sync(RS_FULL_MASK, vis, pl[u]);
                vp[u] = on ? vp[u] : 0.f;
              }
              float rv[UN][NCH];
#pragma unroll
              for (int u = 0; u < UN; ++u)
#pragma unroll
                for (int cch = 0; cch < NCH; ++cch)
                  rv[u][cch] = (cch * 32 + 32 <= DP || cch * 32 + lane < DP) ? rows[pl[u] * DP + cch * 32] : 0.f;
#pragma unroll
              for (int u = 0; u < UN; ++u)
#pragma unroll
                for (int cch = 0; cch < NCH; ++cch) acc[cch] = fmaf(vp[u], rv[u][cch], acc[cch]);
            }
            float* dst = a.color_grad + (size_t)row * DP + lane;
#pragma unroll
            for (int cch = 0; cch < NCH; ++cch)
              if (cch * 32 + lane < a.D) atomicAdd(dst + cch * 32, acc[cch]);
          } else {
            commit_color_grads<DP, 0>(v_c, vis, a.color_grad + (size_t)row * DP, a.D, lane);
          }
        }
      }
    }
    if (b >= 2 && t < BATCH) s.ids[b & 1][t] = next_id;
  }
  rs::cp_async_wait_all();
  if constexpr (CMMA) {
    if (cm_n > 0) {   // contract what is still parked (warp-uniform); unused slots count as zero visibility
      for (int g = cm_n; g < CM_PEND; ++g) {
        cm_vis[g * CM_VSTRIDE + lane] = 0.f;
        if (lane == 0) cm_row[g] = 0;
      }
      __syncwarp();
      flush_color_mma<DP>(s_vc + (size_t)(warp * 32) * DP, cm_vis, cm_row, cm_n, a.color_grad, a.D, lane);
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward, 2 px/lane
// Backward counterpart of rasterize_fwd2_kernel (<= 4 colour channels): a warp owns an 8x8 pixel block, every
// lane differentiates the two pixels (x, y) and (x, y+4).  The two pixels share the Gaussian record loads, the
// x-dependent terms and -- the point of the variant -- ONE 16-value warp reduction and ONE 64-byte RED per
// (warp, Gaussian): their contributions are summed in registers first.  Because dx is common to the two pixels
// the record's moments factor as dx * (sum over the two pixels), which removes most per-pixel multiplies.
template <int BATCH, bool ABSGRAD, int MINB, int NW, int S = 2, bool EDEV = false>
__global__ void __launch_bounds__(NW * 32, MINB * (4 / NW)) rasterize_bwd2_kernel(const RasterArgs a) {
  constexpr int DP = 4;
  constexpr int NT = NW * 32, SUB = 4 / NW, IPT = BATCH / NT;   // NW warps per CTA; IPT ids per thread and batch
  static_assert(BATCH % NT == 0, "whole ids per thread");
  static_assert(S == 2 || BATCH == NT, "ring: one slot per thread");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem<DP, BATCH, S>& s = *reinterpret_cast<Smem<DP, BATCH, S>*>(smem_raw);
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int tile_id = blockIdx.x / SUB;
  const int qd = (blockIdx.x - tile_id * SUB) * NW + warp;   // which 8x8 block of the tile this warp owns
  const int tiles_per_cam = a.tile_w * a.tile_h;
  const int cam = tile_id / tiles_per_cam;
  const int tl = tile_id - cam * tiles_per_cam;
  const int tyi = tl / a.tile_w, txi = tl - tyi * a.tile_w;
  const int start = __ldg(a.offsets + tile_id);
  const int end = list_end_of<EDEV>(a, tile_id, a.C * tiles_per_cam);
  const int x0 = txi * RS_TILE + (qd & 1) * 8, y0 = tyi * RS_TILE + (qd >> 1) * 8;
  const int pxi = x0 + (lane & 7);
  const float px = pxi + 0.5f;
  const float rcx = x0 + 4.0f, rcy = y0 + 4.0f;

  float py[2], v_c[2][4], v_ds[2], v_n[2][3], tfin[2], T[2], R[2];
  int last_id[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int pyi = y0 + (lane >> 3) + 4 * h;
    const bool inside = pxi < a.W && pyi < a.H;
    const size_t pix = inside ? ((size_t)cam * a.H + pyi) * a.W + pxi : 0;
    py[h] = pyi + 0.5f;
#pragma unroll
    for (int k = 0; k < DP; ++k) v_c[h][k] = (inside && k < a.D) ? __ldg(a.v_colors + pix * a.D + k) : 0.f;
    const float T_final = inside ? a.out_T[pix] : 1.f;
    float v_alpha_ed = 0.f;
    if (a.ed_channel >= 0 && inside) {   // select form: a dynamic index would push v_c into local memory
      const float alpha_out = 1.f - T_final;
      const float sc = 1.f / fmaxf(alpha_out, 1e-10f);
      const float oc = __ldg(a.out_colors + pix * a.D + a.ed_channel);
      const float vce = __ldg(a.v_colors + pix * a.D + a.ed_channel);   // re-read: v_c[ed_channel] would be a dynamic index
      if (alpha_out > 1e-10f) v_alpha_ed = -oc * vce * sc;
#pragma unroll
      for (int k = 0; k < DP; ++k) v_c[h][k] *= (k == a.ed_channel) ? sc : 1.f;
    }
    last_id[h] = inside ? a.last_ids[pix] : start - 1;   // an outside pixel owns no list entry: never "valid"
    const float il = inside ? inv_ray_len(a, cam, px, py[h]) : 0.f;
    v_ds[h] = inside ? __ldg(a.v_dexp + pix) * il : 0.f;
    if (inside) commit_median_grad(a, a.median_ids[pix], __ldg(a.v_dmed + pix) * il, px, py[h]);
#pragma unroll
    for (int k = 0; k < 3; ++k) v_n[h][k] = inside ? __ldg(a.v_normals + pix * 3 + k) : 0.f;
    float bgdot = 0.f;
    if (a.backgrounds) {
#pragma unroll
      for (int k = 0; k < DP; ++k)
        if (k < a.D) bgdot += __ldg(a.backgrounds + (size_t)cam * a.D + k) * v_c[h][k];
    }
    tfin[h] = inside ? T_final * (__ldg(a.v_alphas + pix) + v_alpha_ed - bgdot) : 0.f;
    T[h] = T_final;
    R[h] = 0.f;
  }

  // Select-free reduction of the eight "visibility x per-pixel constant" sums (normals 3, colours 4, ray_t): lane L
  // keeps them in registers permuted by xa(L) = lane bits (4,3,2), i.e. register r holds slot r ^ xa, so that at
  // every recursive-halving step "send the upper half of the registers, keep the lower half" is right for every
  // lane.  The permutation is applied once, here, to the per-pixel constants.
  const int xa = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
  float ca[2][8];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const float src[8] = {v_n[h][0], v_n[h][1], v_n[h][2], v_c[h][0], v_c[h][1], v_c[h][2], v_c[h][3], v_ds[h]};
#pragma unroll
    for (int r = 0; r < 8; ++r) ca[h][r] = src[r];
#pragma unroll
    for (int bit = 4; bit >= 1; bit >>= 1) {   // r -> r ^ xa as three conditional swaps
      const bool sw = (xa & bit) != 0;
#pragma unroll
      for (int r = 0; r < 8; ++r)
        if ((r & bit) == 0) {
          const float lo = ca[h][r], up = ca[h][r | bit];
          ca[h][r] = sw ? up : lo;
          ca[h][r | bit] = sw ? lo : up;
        }
    }
  }
  // record slot this lane commits: lanes with bit 1 clear end up with slot xa of group A (gn 9..11, colours 12..15,
  // g_ray_t 6), the others with slot xa of group B (S10 0, S01 1, ga 2, gc 4, g_rpx 7, g_rpy 8, S00 5, gb 3)
  const int rec_slot = (int)(((lane & 2) ? 0x35874210u : 0x6FEDCBA9u) >> (4 * xa)) & 15;

  int warp_last = max(last_id[0], last_id[1]);
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) warp_last = max(warp_last, __shfl_xor_sync(RS_FULL_MASK, warp_last, d));
  if (lane == 0) s.red[warp] = warp_last;
  __syncthreads();
  int blk_last = start - 1;
#pragma unroll
  for (int w = 0; w < NW; ++w) blk_last = max(blk_last, s.red[w]);
  const int nb = blk_last >= start ? (blk_last - start) / BATCH + 1 : 0;

  int ring_id = 0;   // ring: flatten id of this thread's slot in the next batch to gather
  if constexpr (S == 2) {
    if (nb > 0) {
      const int bl = nb - 1;
#pragma unroll
      for (int r = 0; r < IPT; ++r) {
        const int i = start + bl * BATCH + r * NT + t;
        s.ids[bl & 1][r * NT + t] = i < end ? __ldg(a.flatten_ids + i) : 0;
      }
      __syncthreads();
      issue_gather<DP, BATCH, NT>(s, bl & 1, min(BATCH, end - (start + bl * BATCH)), a, t);
      if (bl >= 1) {
#pragma unroll
        for (int r = 0; r < IPT; ++r) s.ids[(bl - 1) & 1][r * NT + t] = __ldg(a.flatten_ids + start + (bl - 1) * BATCH + r * NT + t);
      }
    }
  } else {
    // ring: iteration it = nb - 1 - b walks the batches back to front; stage = it % S (see ring_gather_share)
    if (t == 0) {
#pragma unroll
      for (int i = 0; i < S; ++i) { mbar_init(&s.full[i], NT); mbar_init(&s.empty[i], NW); }
      s.done = 0;
    }
    __syncthreads();
    if (nb > 0) {
      const int bl = nb - 1;
      const int i = start + bl * BATCH + t;
      ring_id = i < end ? __ldg(a.flatten_ids + i) : 0;
      ring_gather_share(s, 0, min(BATCH, end - (start + bl * BATCH)), ring_id, a, lane, warp);
      if (bl >= 1) ring_id = __ldg(a.flatten_ids + start + (bl - 1) * BATCH + t);
    }
  }
  for (int b = nb - 1; b >= 0; --b) {
    int next_id[IPT];
    const int it = nb - 1 - b;
    if constexpr (S == 2) {
      rs::cp_async_wait_all();
      __syncthreads();
      if (b >= 1) {
        issue_gather<DP, BATCH, NT>(s, (b - 1) & 1, BATCH, a, t);
        if (b >= 2) {
#pragma unroll
          for (int r = 0; r < IPT; ++r) next_id[r] = __ldg(a.flatten_ids + start + (b - 2) * BATCH + r * NT + t);
        }
      }
    } else {
      if (b >= 1) {                                            // refill the stage iteration it + 1 - S lived in
        const int st1 = (it + 1) % S;
        if (it + 1 >= S) mbar_wait_or_done(&s.empty[st1], (unsigned)(((it + 1 - S) / S) & 1), &s.done, 1 << 30);
        ring_gather_share(s, st1, BATCH, ring_id, a, lane, warp);
        if (b >= 2) ring_id = __ldg(a.flatten_ids + start + (b - 2) * BATCH + t);
      }
      mbar_wait_or_done(&s.full[it % S], (unsigned)((it / S) & 1), &s.done, 1 << 30);
    }
    const int buf = S == 2 ? (b & 1) : (it % S);
    const int base_idx = start + b * BATCH;
    const int hi = min(min(BATCH, end - base_idx) - 1, warp_last - base_idx);
    for (int g0 = hi >= 0 ? (hi & ~31) : -32; g0 >= 0; g0 -= 32) {
      const int j = g0 + lane;
      bool hit = false;
      if (j <= hi) {
        hit = footprint_hit(s, buf, j, a.exact_cull != 0, rcx, rcy, 3.5f, 3.5f);
      }
      unsigned m = __ballot_sync(RS_FULL_MASK, hit);
      while (m) {
        const int bit = 31 - __clz(m);   // back to front
        m &= ~(1u << bit);
        const int jj = g0 + bit;
        const int idx = base_idx + jj;
        const float2 xy = *reinterpret_cast<const float2*>(&s.q0[buf][jj]);
        const float4 q1 = s.q1[buf][jj];
        const float dx = xy.x - px;
        const float adx2 = __fmul_rn(__fmul_rn(q1.x, dx), dx), bdx = __fmul_rn(q1.y, dx);
        float dy[2], am[2];
        bool valid[2], unc[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          dy[h] = xy.y - py[h];
          const float sig = sigma_col(adx2, bdx, q1.z, dy[h]);
          const float oe = q1.w * rs::fast_exp2(-sig);
          const float alpha = fminf(RS_ALPHA_MAX, oe);
          valid[h] = idx <= last_id[h] && sig >= 0.f && alpha >= RS_ALPHA_MIN;
          am[h] = valid[h] ? alpha : 0.f;   // an invalid pair acts as alpha = 0: every term below is an exact 0
          unc[h] = valid[h] && oe <= RS_ALPHA_MAX;
        }
        if (!__any_sync(RS_FULL_MASK, valid[0] || valid[1])) continue;
        const float4 q2 = s.q2[buf][jj], q3 = s.q3[buf][jj];
        const float4 cc = *reinterpret_cast<const float4*>(&s.col[buf][jj][0]);
        const float tb = fmaf(q2.y, dx, q2.x);
        float vis[2], v_t[2], v_sig[2], vsy[2], ax = 0.f, ay = 0.f;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float ra = rs::fast_rcp(1.f - am[h]);
          T[h] *= ra;   // transmittance in front of this Gaussian
          vis[h] = am[h] * T[h];
          const float tt = fmaf(q2.z, dy[h], tb);
          float w = v_ds[h] * tt;
          w = fmaf(v_n[h][0], q3.x, w); w = fmaf(v_n[h][1], q3.y, w); w = fmaf(v_n[h][2], q3.z, w);
          w = fmaf(v_c[h][0], cc.x, w); w = fmaf(v_c[h][1], cc.y, w);
          w = fmaf(v_c[h][2], cc.z, w); w = fmaf(v_c[h][3], cc.w, w);
          const float v_alpha = fmaf(T[h], w, ra * (tfin[h] - R[h]));
          R[h] = fmaf(vis[h], w, R[h]);
          v_t[h] = vis[h] * v_ds[h];   // (the median Gaussian's extra term: commit_median_grad)
          v_sig[h] = unc[h] ? -am[h] * v_alpha : 0.f;
          vsy[h] = v_sig[h] * dy[h];
          if constexpr (ABSGRAD) {
            const float vs2 = v_sig[h] * RS_LN2;
            ax += fabsf(vs2 * (2.f * q1.x * dx + q1.y * dy[h]) + v_t[h] * q2.y);
            ay += fabsf(vs2 * (q1.y * dx + 2.f * q1.z * dy[h]) + v_t[h] * q2.z);
          }
        }
        // the two pixels share dx: the record's moments are dx * (sums over the pair)
        const float s0 = v_sig[0] + v_sig[1], sy = vsy[0] + vsy[1], t0 = v_t[0] + v_t[1];
        float ga[8], gb[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) ga[r] = fmaf(vis[1], ca[1][r], vis[0] * ca[0][r]);
        gb[0] = dx * s0; gb[1] = sy; gb[2] = 0.5f * dx * gb[0]; gb[3] = 0.5f * fmaf(vsy[1], dy[1], vsy[0] * dy[0]);
        gb[4] = dx * t0; gb[5] = fmaf(v_t[1], dy[1], v_t[0] * dy[0]); gb[6] = s0; gb[7] = dx * sy;
        float* const rec = a.geom_grad + (size_t)__float_as_int(q2.w) * 16;
        // A: 7 shuffles, no selects
#pragma unroll
        for (int r = 0; r < 4; ++r) ga[r] += __shfl_xor_sync(RS_FULL_MASK, ga[r + 4], 16);
#pragma unroll
        for (int r = 0; r < 2; ++r) ga[r] += __shfl_xor_sync(RS_FULL_MASK, ga[r + 2], 8);
        ga[0] += __shfl_xor_sync(RS_FULL_MASK, ga[1], 4);
        // B: recursive halving with selects, down to the 4-lane classes
        {
          int dist = 16;
#pragma unroll
          for (int half = 4; half >= 1; half /= 2) {
            const bool upper = (lane & dist) != 0;
#pragma unroll
            for (int r = 0; r < half; ++r) {
              const float send = upper ? gb[r] : gb[r + half];
              const float keep = upper ? gb[r + half] : gb[r];
              gb[r] = keep + __shfl_xor_sync(RS_FULL_MASK, send, dist);
            }
            dist >>= 1;
          }
        }
        // lanes with bit 1 clear take the A slot, the others the B slot; then add the two halves
        {
          const bool upper = (lane & 2) != 0;
          const float send = upper ? ga[0] : gb[0];
          float keep = upper ? gb[0] : ga[0];
          keep += __shfl_xor_sync(RS_FULL_MASK, send, 2);
          keep += __shfl_xor_sync(RS_FULL_MASK, keep, 1);
          if ((lane & 1) == 0) atomicAdd(rec + rec_slot, keep);
        }
        const int id = __float_as_int(q2.w);
        if constexpr (ABSGRAD) {
          float ab[2] = {ax, ay};
          rs::warp_reduce_scatter<2>(ab, lane);
          if ((lane & 15) == 0) atomicAdd(a.abs_grad + (size_t)id * 2 + (lane >> 4), ab[0]);
        }
      }
    }
    if constexpr (S == 2) {
      if (b >= 2) {
#pragma unroll
        for (int r = 0; r < IPT; ++r) s.ids[b & 1][r * NT + t] = next_id[r];
      }
    } else {
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.empty[it % S]);            // this warp is through with the stage
    }
  }
  rs::cp_async_wait_all();
}

// ------------------------------------------------------------------------------------------------ backward, 2 px/lane, MMA reduction
// Same per-pixel arithmetic as rasterize_bwd2_kernel, but the per-Gaussian reduction over the warp's 64 pixels is a
// tensor-core contraction instead of a shuffle tree.  Every one of the 16 sums of a (warp, Gaussian) pair is linear in
// one of two per-pixel factors with PER-PIXEL CONSTANT coefficients:
//   vis[p]     x { v_n(3), v_c(4), v_ds, v_ds*u, v_ds*v }            (10 columns; u, v = pixel offset from the warp
//   v_sigma[p] x { 1, u, v, u^2, u v, v^2 }                          ( 6 columns   rectangle's centre: +-0.5 .. +-3.5)
// so a lane only PARKS (vis, v_sigma) of its two pixels (one 16-byte shared store) and every B3_PEND Gaussians the warp
// contracts the parked block with mma.sync m16n8k8 (TF32): A = the constants, transposed (rows = output columns,
// k = 8 pixels per step, 8 steps), B = parked values (k = pixels, n = 8 Gaussians).  The parked values are split
// x = hi + lo on the fly and the arbitrary constants once per tile (3 products per step; the monomials are exact in
// TF32: 2 products), so the sums keep fp32-level accuracy (|error| <= 2^-20 of each term).  The raw moments about the
// rectangle centre are shifted to the Gaussian's centre (dx = X - u) and committed with one 16-byte vector RED per
// lane: 8 Gaussians x 64 bytes per instruction.  Replaces 16 SHFL + ~60 ALU/FMA issue slots per (warp, Gaussian) by
// ~25.
constexpr int B3_PEND = 8;       // Gaussians parked per warp before a contraction (= the MMA's N)
constexpr int B3_PSTRIDE = 36;   // float4 row stride of the parking lot: fragment loads hit 8 distinct 16-byte bank groups
constexpr int B3_MSTRIDE = 20;   // float row stride of the moment rows (conflict-free transposing stores)

// A fragments (m16n8k8: a0 = (row gid, k tig), a1 = (row gid + 8, k tig), a2 = (row gid, k tig + 4), a3 = (row gid + 8,
// k tig + 4)) are stored as ready-made register quads, one 16-byte load each.  The 10 "vis" columns sit in rows 0..4 and
// 8..12, the 6 monomials in rows 0..2 and 8..10, so that only lanes with gid < 5 (gid < 3) hold anything and the others
// keep a quad of zeros in registers for the whole contraction.
// The monomial quads: entry [ks][lane < 12] = rows gid, gid + 3 of {1, u, v | u^2, u v, v^2} at pixel lane 4 ks + tig,
// pixels h0 / h1 (u = (l & 7) - 3.5, v = (l >> 3) + 4 h - 3.5) -- the same for every warp of every tile.
struct B3SigTab { float4 v[8 * 12]; };
constexpr float b3_mono(int row, float u, float v) {
  return row == 0 ? 1.f : row == 1 ? u : row == 2 ? v : row == 3 ? u * u : row == 4 ? u * v : v * v;
}
constexpr B3SigTab b3_make_sig() {
  B3SigTab t{};
  for (int ks = 0; ks < 8; ++ks)
    for (int lane = 0; lane < 12; ++lane) {
      const int row = lane >> 2, src = 4 * ks + (lane & 3);
      const float u = (float)(src & 7) - 3.5f, v0 = (float)(src >> 3) - 3.5f, v1 = v0 + 4.f;
      t.v[ks * 12 + lane] = float4{b3_mono(row, u, v0), b3_mono(row + 3, u, v0), b3_mono(row, u, v1), b3_mono(row + 3, u, v1)};
    }
  return t;
}
__device__ const B3SigTab g_b3_sig = b3_make_sig();

struct Bwd3Warp {                          // per-warp shared memory.  Everything is laid out in 8-byte pairs: measured on
                                           // B200, LDS.64 / STS.64 move 128 B/clk/SM but LDS.128 only 64 (scripts/ubench)
  float2 pvis[B3_PEND * B3_PSTRIDE];       // [g][pixel lane] = vis of pixels (h0, h1)
  float2 psig[B3_PEND * B3_PSTRIDE];       // [g][pixel lane] = v_sigma of pixels (h0, h1)
  float4 meta[B3_PEND];                    // (x_g - rcx, y_g - rcy, flatten id bits, -)
  float mom[B3_PEND * B3_MSTRIDE];         // [g][column] raw moments, transposed out of the C fragments
  float2 cah[2][8 * 20];                   // TF32 heads of the "vis" constants: [a0 a1 | a2 a3][k step][lane < 20]
  float2 cal[2][8 * 20];                   // their tails
  float2 zero;                             // what the lanes without constant rows read
};
struct Bwd3Cta {                           // per-CTA shared memory
  float2 sig[2][8 * 12];                   // g_b3_sig as pairs: [a0 a1 | a2 a3][k step][lane < 12]
};

__device__ __forceinline__ void b3_split(float x, unsigned& hi, unsigned& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;        // TF32 head by truncation (exactly representable)
  lo = __float_as_uint(x - __uint_as_float(hi)); // exact remainder; the tensor core reads its leading 11 bits
}

__device__ __forceinline__ void bwd3_flush(Bwd3Warp& w, const Bwd3Cta& wc, int n, float* __restrict__ geom_grad,
                                           int lane) {
  const int gid = lane >> 2, tig = lane & 3;
  __syncwarp();   // the parked rows are complete
  // five independent accumulator chains (one per product), 8 k steps deep each, summed at the end: the HMMAs of one
  // chain are ~30 cycles apart, so a single chain per group would serialise 24 of them
  float c1[4] = {0.f, 0.f, 0.f, 0.f}, c2[4] = {0.f, 0.f, 0.f, 0.f}, c3[4] = {0.f, 0.f, 0.f, 0.f};
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  const float2* pv = w.pvis + gid * B3_PSTRIDE + tig;
  const float2* ps = w.psig + gid * B3_PSTRIDE + tig;
  // lanes without rows (gid >= 5, gid >= 3) read a pair of zeros instead (stride 0): the loop body stays branch free
  const bool hv = lane < 20, hs = lane < 12;
  const float2* ph0 = hv ? w.cah[0] + lane : &w.zero;
  const float2* ph1 = hv ? w.cah[1] + lane : &w.zero;
  const float2* pl0 = hv ? w.cal[0] + lane : &w.zero;
  const float2* pl1 = hv ? w.cal[1] + lane : &w.zero;
  const float2* pg0 = hs ? wc.sig[0] + lane : &w.zero;
  const float2* pg1 = hs ? wc.sig[1] + lane : &w.zero;
  const int sv = hv ? 20 : 0, sg = hs ? 12 : 0;
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    // Gaussian gid, pixel lane 4 ks + tig: k = tig is its pixel h0, k = tig + 4 its pixel h1
    const float2 v = pv[4 * ks], sgm = ps[4 * ks];
    const float2 h01 = ph0[ks * sv], h23 = ph1[ks * sv], l01 = pl0[ks * sv], l23 = pl1[ks * sv];
    const float2 g01 = pg0[ks * sg], g23 = pg1[ks * sg];
    unsigned vh0, vl0, vh1, vl1, sh0, sl0, sh1, sl1;
    b3_split(v.x, vh0, vl0); b3_split(v.y, vh1, vl1);
    b3_split(sgm.x, sh0, sl0); b3_split(sgm.y, sh1, sl1);
    const unsigned ah[4] = {__float_as_uint(h01.x), __float_as_uint(h01.y), __float_as_uint(h23.x), __float_as_uint(h23.y)};
    const unsigned al[4] = {__float_as_uint(l01.x), __float_as_uint(l01.y), __float_as_uint(l23.x), __float_as_uint(l23.y)};
    const unsigned as[4] = {__float_as_uint(g01.x), __float_as_uint(g01.y), __float_as_uint(g23.x), __float_as_uint(g23.y)};
    mma_tf32_16x8x8(c1, ah, vh0, vh1);
    mma_tf32_16x8x8(s1, as, sh0, sh1);
    mma_tf32_16x8x8(c2, al, vh0, vh1);
    mma_tf32_16x8x8(s2, as, sl0, sl1);
    mma_tf32_16x8x8(c3, ah, vl0, vl1);
  }
  float cv[4], cs[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { cv[i] = c1[i] + (c2[i] + c3[i]); cs[i] = s1[i] + s2[i]; }
  // C fragments -> mom[g][column]: (c0, c1) = row gid, (c2, c3) = row gid + 8 of Gaussians 2 tig, 2 tig + 1
  float* const m = w.mom;
  if (gid < 5) {
    m[(2 * tig) * B3_MSTRIDE + gid] = cv[0];
    m[(2 * tig + 1) * B3_MSTRIDE + gid] = cv[1];
    m[(2 * tig) * B3_MSTRIDE + 5 + gid] = cv[2];
    m[(2 * tig + 1) * B3_MSTRIDE + 5 + gid] = cv[3];
  }
  if (gid < 3) {
    m[(2 * tig) * B3_MSTRIDE + 10 + gid] = cs[0];
    m[(2 * tig + 1) * B3_MSTRIDE + 10 + gid] = cs[1];
    m[(2 * tig) * B3_MSTRIDE + 13 + gid] = cs[2];
    m[(2 * tig + 1) * B3_MSTRIDE + 13 + gid] = cs[3];
  }
  __syncwarp();
  // lane (g, q): quarter q of Gaussian g's 16-float gradient record.  Moments about the rectangle centre -> sums over
  // dx = X - u, dy = Y - v:  S v dx = X M00 - M10,  S v dx^2 = X (S v dx) - (X M10 - M20),  S v dx dy = X (S v dy) - (Y M10 - M11) ...
  const int g = gid, q = tig;
  if (g < n) {
    const float4* r = reinterpret_cast<const float4*>(m + g * B3_MSTRIDE);
    const float4 m0 = r[0], m1 = r[1], m2 = r[2], m3 = r[3], mt = w.meta[g];
    // m0 = (n0 n1 n2 c0)  m1 = (c1 c2 c3 Vt0)  m2 = (Vtu Vtv M00 M10)  m3 = (M01 M20 M11 M02)
    const float X = mt.x, Y = mt.y;
    float4 out;
    if (q == 0) {
      const float s10 = fmaf(X, m2.z, -m2.w), s01 = fmaf(Y, m2.z, -m3.x);
      out.x = s10;
      out.y = s01;
      out.z = 0.5f * fmaf(X, s10, -fmaf(X, m2.w, -m3.y));
      out.w = fmaf(X, s01, -fmaf(Y, m2.w, -m3.z));
    } else if (q == 1) {
      const float s01 = fmaf(Y, m2.z, -m3.x);
      out.x = 0.5f * fmaf(Y, s01, -fmaf(Y, m3.x, -m3.w));
      out.y = m2.z;
      out.z = m1.w;
      out.w = fmaf(X, m1.w, -m2.x);
    } else if (q == 2) {
      out = make_float4(fmaf(Y, m1.w, -m2.y), m0.x, m0.y, m0.z);
    } else {
      out = make_float4(m0.w, m1.x, m1.y, m1.z);
    }
    atomicAdd(reinterpret_cast<float4*>(geom_grad + (size_t)__float_as_int(mt.z) * 16 + 4 * q), out);
  }
}

template <int BATCH, int MINB, bool EDEV = false>
__global__ void __launch_bounds__(RT2, MINB) rasterize_bwd3_kernel(const RasterArgs a) {
  constexpr int DP = 4;
  static_assert(BATCH <= RT2, "at most one id per thread");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem<DP, BATCH>& s = *reinterpret_cast<Smem<DP, BATCH>*>(smem_raw);
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  Bwd3Warp& w = reinterpret_cast<Bwd3Warp*>(smem_raw + sizeof(Smem<DP, BATCH>))[warp];
  Bwd3Cta& wc = *reinterpret_cast<Bwd3Cta*>(smem_raw + sizeof(Smem<DP, BATCH>) + (RT2 / 32) * sizeof(Bwd3Warp));
  const int tile_id = blockIdx.x;
  const int tiles_per_cam = a.tile_w * a.tile_h;
  const int cam = tile_id / tiles_per_cam;
  const int tl = tile_id - cam * tiles_per_cam;
  const int tyi = tl / a.tile_w, txi = tl - tyi * a.tile_w;
  const int start = __ldg(a.offsets + tile_id);
  const int end = list_end_of<EDEV>(a, tile_id, a.C * tiles_per_cam);
  const int x0 = txi * RS_TILE + (warp & 1) * 8, y0 = tyi * RS_TILE + (warp >> 1) * 8;
  const int pxi = x0 + (lane & 7);
  const float px = pxi + 0.5f;
  const float rcx = x0 + 4.0f, rcy = y0 + 4.0f;

  float py[2], v_c[2][4], v_ds[2], v_n[2][3], tfin[2], T[2], R[2];
  int last_id[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int pyi = y0 + (lane >> 3) + 4 * h;
    const bool inside = pxi < a.W && pyi < a.H;
    const size_t pix = inside ? ((size_t)cam * a.H + pyi) * a.W + pxi : 0;
    py[h] = pyi + 0.5f;
#pragma unroll
    for (int k = 0; k < DP; ++k) v_c[h][k] = (inside && k < a.D) ? __ldg(a.v_colors + pix * a.D + k) : 0.f;
    const float T_final = inside ? a.out_T[pix] : 1.f;
    float v_alpha_ed = 0.f;
    if (a.ed_channel >= 0 && inside) {   // select form: a dynamic index would push v_c into local memory
      const float alpha_out = 1.f - T_final;
      const float sc = 1.f / fmaxf(alpha_out, 1e-10f);
      const float oc = __ldg(a.out_colors + pix * a.D + a.ed_channel);
      const float vce = __ldg(a.v_colors + pix * a.D + a.ed_channel);
      if (alpha_out > 1e-10f) v_alpha_ed = -oc * vce * sc;
#pragma unroll
      for (int k = 0; k < DP; ++k) v_c[h][k] *= (k == a.ed_channel) ? sc : 1.f;
    }
    last_id[h] = inside ? a.last_ids[pix] : start - 1;   // an outside pixel owns no list entry: never "valid"
    const float il = inside ? inv_ray_len(a, cam, px, py[h]) : 0.f;
    v_ds[h] = inside ? __ldg(a.v_dexp + pix) * il : 0.f;
    if (inside) commit_median_grad(a, a.median_ids[pix], __ldg(a.v_dmed + pix) * il, px, py[h]);
#pragma unroll
    for (int k = 0; k < 3; ++k) v_n[h][k] = inside ? __ldg(a.v_normals + pix * 3 + k) : 0.f;
    float bgdot = 0.f;
    if (a.backgrounds) {
#pragma unroll
      for (int k = 0; k < DP; ++k)
        if (k < a.D) bgdot += __ldg(a.backgrounds + (size_t)cam * a.D + k) * v_c[h][k];
    }
    tfin[h] = inside ? T_final * (__ldg(a.v_alphas + pix) + v_alpha_ed - bgdot) : 0.f;
    T[h] = T_final;
    R[h] = 0.f;
  }

  // ---- A fragments of the constants (once per tile).  k step ks covers pixel lanes 4 ks .. 4 ks + 3: k = j is pixel
  // h0 of lane 4 ks + j, k = j + 4 its pixel h1.  Columns: 0..2 v_n, 3..6 v_c, 7 v_ds, 8 v_ds*u, 9 v_ds*v; lane
  // (gid < 5, tig) holds columns gid (row gid) and 5 + gid (row gid + 8).  Heads by truncation, tails exact (the
  // tensor core reads their leading 11 bits).
  {
    const int gid = lane >> 2, tig = lane & 3;
    const float u_own = (float)(lane & 7) - 3.5f;
    float4* sc = reinterpret_cast<float4*>(w.pvis);   // scratch [lane][5 float4]: 10 columns x 2 pixels (pvis + psig)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float v_own = (float)((lane >> 3) + 4 * h) - 3.5f;
      float* d = reinterpret_cast<float*>(sc + lane * 5) + h * 10;
      d[0] = v_n[h][0]; d[1] = v_n[h][1]; d[2] = v_n[h][2];
      d[3] = v_c[h][0]; d[4] = v_c[h][1]; d[5] = v_c[h][2]; d[6] = v_c[h][3];
      d[7] = v_ds[h]; d[8] = v_ds[h] * u_own; d[9] = v_ds[h] * v_own;
    }
    if (t < 8 * 12) {                            // (visible to the other warps after the barrier below)
      const float4 q = g_b3_sig.v[t];
      wc.sig[0][t] = make_float2(q.x, q.y);
      wc.sig[1][t] = make_float2(q.z, q.w);
    }
    if (lane == 0) w.zero = make_float2(0.f, 0.f);
    __syncwarp();
    if (lane < 20) {
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const float* r = reinterpret_cast<const float*>(sc + (4 * ks + tig) * 5);
        const float c[4] = {r[gid], r[5 + gid], r[10 + gid], r[15 + gid]};
        float hi[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) hi[i] = __uint_as_float(__float_as_uint(c[i]) & 0xffffe000u);
        w.cah[0][ks * 20 + lane] = make_float2(hi[0], hi[1]);
        w.cah[1][ks * 20 + lane] = make_float2(hi[2], hi[3]);
        w.cal[0][ks * 20 + lane] = make_float2(c[0] - hi[0], c[1] - hi[1]);
        w.cal[1][ks * 20 + lane] = make_float2(c[2] - hi[2], c[3] - hi[3]);
      }
    }
    __syncwarp();   // the scratch rows (parking lot) may be overwritten from here on
  }

  int warp_last = max(last_id[0], last_id[1]);
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) warp_last = max(warp_last, __shfl_xor_sync(RS_FULL_MASK, warp_last, d));
  if (lane == 0) s.red[warp] = warp_last;
  __syncthreads();
  int blk_last = start - 1;
#pragma unroll
  for (int wi = 0; wi < RT2 / 32; ++wi) blk_last = max(blk_last, s.red[wi]);
  const int nb = blk_last >= start ? (blk_last - start) / BATCH + 1 : 0;

  if (nb > 0) {
    const int bl = nb - 1;
    if (t < BATCH) { const int i = start + bl * BATCH + t; s.ids[bl & 1][t] = i < end ? __ldg(a.flatten_ids + i) : 0; }
    __syncthreads();
    issue_gather<DP, BATCH, RT2>(s, bl & 1, min(BATCH, end - (start + bl * BATCH)), a, t);
    if (bl >= 1 && t < BATCH) s.ids[(bl - 1) & 1][t] = __ldg(a.flatten_ids + start + (bl - 1) * BATCH + t);
  }
  int np = 0;   // Gaussians parked
  for (int b = nb - 1; b >= 0; --b) {
    rs::cp_async_wait_all();
    __syncthreads();
    int next_id = 0;
    if (b >= 1) {
      issue_gather<DP, BATCH, RT2>(s, (b - 1) & 1, BATCH, a, t);
      if (b >= 2 && t < BATCH) next_id = __ldg(a.flatten_ids + start + (b - 2) * BATCH + t);
    }
    const int buf = b & 1;
    const int base_idx = start + b * BATCH;
    const int hi = min(min(BATCH, end - base_idx) - 1, warp_last - base_idx);
    for (int g0 = hi >= 0 ? (hi & ~31) : -32; g0 >= 0; g0 -= 32) {
      const int j = g0 + lane;
      bool hit = false;
      if (j <= hi) hit = footprint_hit(s, buf, j, a.exact_cull != 0, rcx, rcy, 3.5f, 3.5f);
      unsigned m = __ballot_sync(RS_FULL_MASK, hit);
      while (m) {
        const int bit = 31 - __clz(m);   // back to front
        m &= ~(1u << bit);
        const int jj = g0 + bit;
        const int idx = base_idx + jj;
        const float2 xy = *reinterpret_cast<const float2*>(&s.q0[buf][jj]);
        const float4 q1 = s.q1[buf][jj];
        const float dx = xy.x - px;
        const float adx2 = __fmul_rn(__fmul_rn(q1.x, dx), dx), bdx = __fmul_rn(q1.y, dx);
        float dy[2], am[2];
        bool valid[2], unc[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          dy[h] = xy.y - py[h];
          const float sig = sigma_col(adx2, bdx, q1.z, dy[h]);
          const float oe = q1.w * rs::fast_exp2(-sig);
          const float alpha = fminf(RS_ALPHA_MAX, oe);
          valid[h] = idx <= last_id[h] && sig >= 0.f && alpha >= RS_ALPHA_MIN;
          am[h] = valid[h] ? alpha : 0.f;   // an invalid pair acts as alpha = 0: every term below is an exact 0
          unc[h] = valid[h] && oe <= RS_ALPHA_MAX;
        }
        if (!__any_sync(RS_FULL_MASK, valid[0] || valid[1])) continue;
        const float4 q2 = s.q2[buf][jj], q3 = s.q3[buf][jj];
        const float4 cc = *reinterpret_cast<const float4*>(&s.col[buf][jj][0]);
        const float tb = fmaf(q2.y, dx, q2.x);
        float pv[4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float ra = rs::fast_rcp(1.f - am[h]);
          T[h] *= ra;   // transmittance in front of this Gaussian
          const float vis = am[h] * T[h];
          const float tt = fmaf(q2.z, dy[h], tb);
          float wv = v_ds[h] * tt;
          wv = fmaf(v_n[h][0], q3.x, wv); wv = fmaf(v_n[h][1], q3.y, wv); wv = fmaf(v_n[h][2], q3.z, wv);
          wv = fmaf(v_c[h][0], cc.x, wv); wv = fmaf(v_c[h][1], cc.y, wv);
          wv = fmaf(v_c[h][2], cc.z, wv); wv = fmaf(v_c[h][3], cc.w, wv);
          const float v_alpha = fmaf(T[h], wv, ra * (tfin[h] - R[h]));
          R[h] = fmaf(vis, wv, R[h]);
          pv[h] = vis;
          pv[2 + h] = unc[h] ? -am[h] * v_alpha : 0.f;   // v_sigma
        }
        w.pvis[np * B3_PSTRIDE + lane] = make_float2(pv[0], pv[1]);
        w.psig[np * B3_PSTRIDE + lane] = make_float2(pv[2], pv[3]);
        if (lane == 0) w.meta[np] = make_float4(xy.x - rcx, xy.y - rcy, q2.w, 0.f);
        if (++np == B3_PEND) {
          bwd3_flush(w, wc, B3_PEND, a.geom_grad, lane);
          np = 0;
        }
      }
    }
    if (b >= 2 && t < BATCH) s.ids[b & 1][t] = next_id;
  }
  rs::cp_async_wait_all();
  if (np > 0) bwd3_flush(w, wc, np, a.geom_grad, lane);   // rows >= np hold stale values: their columns are never committed
}

// ------------------------------------------------------------------------------------------------ launch
template <int DP, bool STATS> int launch_fwd2(const RasterArgs& a, cudaStream_t st) {
  constexpr int B = Batch<DP>::value;
  const size_t smem = sizeof(Smem<DP, B>);
  cudaError_t e = cudaFuncSetAttribute(rasterize_fwd_kernel<DP, B, STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) { rs_set_last_cuda_error((int)e); return RS_ERR_LAUNCH; }
  rasterize_fwd_kernel<DP, B, STATS><<<a.C * a.tile_w * a.tile_h, RT, smem, st>>>(a);
  RS_RETURN_LAST_ERROR();
}
// per-call options (include/rade_b200.h); the values must match the header's
constexpr int F_CULL_BBOX = 0x1, F_ONE_PIXEL = 0x2, F_NO_COLOR_MMA = 0x4, F_BWD_MMA = 0x8, F_FWD_RING = 0x10, F_BWD_BARRIER = 0x20;
static inline int bwd_tune(int flags) { return (flags >> 8) & 0xf; }

template <int DP> int launch_fwd_mma(const RasterArgs& a, cudaStream_t st) {
  constexpr int B = Batch<DP>::value;
  const size_t smem = sizeof(Smem<DP, B>) + (RT / 32) * (sizeof(float) * CM_PEND * CM_VSTRIDE + sizeof(int) * CM_PEND);
  cudaError_t e = cudaFuncSetAttribute(rasterize_fwd_mma_kernel<DP, B>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) { rs_set_last_cuda_error((int)e); return RS_ERR_LAUNCH; }
  rasterize_fwd_mma_kernel<DP, B><<<a.C * a.tile_w * a.tile_h, RT, smem, st>>>(a);
  RS_RETURN_LAST_ERROR();
}

template <int DP> int launch_fwd(const RasterArgs& a, cudaStream_t st) {
  if constexpr (DP >= 64) {   // narrower rows: the C fragments cost more registers than the blend saves
    if (!(a.flags & F_NO_COLOR_MMA) && !a.stats) return launch_fwd_mma<DP>(a, st);
  }
  if constexpr (DP == 4) {
    if (!(a.flags & F_ONE_PIXEL)) {
      constexpr int B2 = 128;
      const int tiles = a.C * a.tile_w * a.tile_h;
      // (CTAs of 2 or 1 warps -- half / quarter tiles, no or less barrier coupling -- were measured slower: every CTA
      // gathers the whole tile list, profiles/r02_warps_per_cta_ab.txt)
      // (every variant exists for a host-side and a device-side intersection count, see list_end_of)
#define RS_FWD2(STATS_, S_)                                                                                          \
  do {                                                                                                               \
    if (a.end_dev) rasterize_fwd2_kernel<DP, B2, STATS_, 4, S_, true><<<tiles, 128, sizeof(Smem<DP, B2, S_>), st>>>(a); \
    else rasterize_fwd2_kernel<DP, B2, STATS_, 4, S_, false><<<tiles, 128, sizeof(Smem<DP, B2, S_>), st>>>(a);       \
  } while (0)
      if (a.stats) RS_FWD2(true, 2);  // counting only
      else if (a.flags & F_FWD_RING) RS_FWD2(false, 3);
      else RS_FWD2(false, 2);
#undef RS_FWD2
      RS_RETURN_LAST_ERROR();
    }
    if (a.stats) return launch_fwd2<DP, true>(a, st);  // instrumented variant (counting only, never timed)
  }
  return launch_fwd2<DP, false>(a, st);
}

// backward batch: wide rows also keep the tile's v_c rows (RT x DP floats) in shared memory, so the staged batches
// are halved there to keep two CTAs per SM
template <int DP> struct BwdBatch { static constexpr int value = DP >= 64 ? 32 : Batch<DP>::value; };

template <int DP, bool ABSGRAD, bool CMMA> int launch_bwd3(const RasterArgs& a, cudaStream_t st) {
  constexpr int B = BwdBatch<DP>::value;
  const size_t smem = sizeof(Smem<DP, B>) + ((DP >= 32 || CMMA) ? sizeof(float) * RT * DP : 0) +
                      (CMMA ? (RT / 32) * (sizeof(float) * CM_PEND * CM_VSTRIDE + sizeof(int) * CM_PEND) : 0);
  cudaError_t e = cudaFuncSetAttribute(rasterize_bwd_kernel<DP, B, ABSGRAD, CMMA>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { rs_set_last_cuda_error((int)e); return RS_ERR_LAUNCH; }
  rasterize_bwd_kernel<DP, B, ABSGRAD, CMMA><<<a.C * a.tile_w * a.tile_h, RT, smem, st>>>(a);
  RS_RETURN_LAST_ERROR();
}
template <int DP, bool ABSGRAD> int launch_bwd2(const RasterArgs& a, cudaStream_t st) {
  if constexpr (DP >= 20) {   // measured: -3 % at 20 channels, -24..-29 % at 32-36, +5 % (worse) at 16
    if (!(a.flags & F_NO_COLOR_MMA)) return launch_bwd3<DP, ABSGRAD, true>(a, st);
  }
  return launch_bwd3<DP, ABSGRAD, false>(a, st);
}
template <int DP> int launch_bwd(const RasterArgs& a, cudaStream_t st) {
  if constexpr (DP == 4) {
    if (!(a.flags & F_ONE_PIXEL)) {   // must match the forward variant: the two share sigma_col()'s rounding
      constexpr int B2 = RT2;
      const size_t smem = sizeof(Smem<DP, B2>);
      const int grid = a.C * a.tile_w * a.tile_h;
      // MINB = CTAs per SM the register allocation is capped for: 4 -> 94 registers, 6 -> 80, 7 -> 72 (no spills)
      const int tune = bwd_tune(a.flags);
      if (!a.abs_grad && (a.flags & F_BWD_MMA)) {   // opt-in (absgrad sums |per-pixel gradient|: not linear, shuffle tree only)
        // RS_RASTER_BWD_TUNE: 0 (default) = 64-Gaussian batches, 4 CTAs/SM; 1 = 128-Gaussian batches, 3 CTAs/SM
#define RS_LAUNCH_BWD3(BT, MINB)                                                                                    \
  do {                                                                                                              \
    const size_t smem3 = sizeof(Smem<DP, BT>) + (RT2 / 32) * sizeof(Bwd3Warp) + sizeof(Bwd3Cta);                    \
    cudaError_t e = cudaFuncSetAttribute(rasterize_bwd3_kernel<BT, MINB, false>,                                    \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3);                  \
    if (e == cudaSuccess)                                                                                           \
      e = cudaFuncSetAttribute(rasterize_bwd3_kernel<BT, MINB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                               (int)smem3);                                                                         \
    if (e != cudaSuccess) { rs_set_last_cuda_error((int)e); return RS_ERR_LAUNCH; }                                 \
    if (a.end_dev) rasterize_bwd3_kernel<BT, MINB, true><<<grid, RT2, smem3, st>>>(a);                              \
    else rasterize_bwd3_kernel<BT, MINB, false><<<grid, RT2, smem3, st>>>(a);                                       \
  } while (0)
        if (tune == 1) RS_LAUNCH_BWD3(128, 3);
        else RS_LAUNCH_BWD3(64, 4);
#undef RS_LAUNCH_BWD3
        RS_RETURN_LAST_ERROR();
      }
#define RS_BWD2(ABS_, MINB_, S_)                                                                                       \
  do {                                                                                                                 \
    if (a.end_dev) rasterize_bwd2_kernel<B2, ABS_, MINB_, 4, S_, true><<<grid, RT2, sizeof(Smem<DP, B2, S_>), st>>>(a); \
    else rasterize_bwd2_kernel<B2, ABS_, MINB_, 4, S_, false><<<grid, RT2, sizeof(Smem<DP, B2, S_>), st>>>(a);         \
  } while (0)
      if (a.abs_grad) RS_BWD2(true, 4, 2);
      else if (tune == 7) RS_BWD2(false, 7, 2);
      else if (tune == 6) RS_BWD2(false, 6, 2);
      else if (a.flags & F_BWD_BARRIER) RS_BWD2(false, 4, 2);
      else if (tune == 2) RS_BWD2(false, 5, 4);
      else RS_BWD2(false, 5, 3);   // default: ring
#undef RS_BWD2
      RS_RETURN_LAST_ERROR();
    }
  }
  return a.abs_grad ? launch_bwd2<DP, true>(a, st) : launch_bwd2<DP, false>(a, st);
}

#define RS_DP_LIST(X) X(4) X(8) X(16) X(20) X(32) X(36) X(64) X(68) X(72)

int padded_channels(int D) {
  const int list[] = {4, 8, 16, 20, 32, 36, 64, 68, 72};
  for (int v : list)
    if (D <= v) return v;
  return -1;
}

bool check_common(const RasterArgs& a) {
  return a.C > 0 && a.N > 0 && a.W > 0 && a.H > 0 && a.tile_w > 0 && a.tile_h > 0 && a.D > 0 && a.M >= 0 && a.geom &&
         a.colors && a.Ks && a.offsets && (a.M == 0 || a.flatten_ids) && a.out_T && a.last_ids && a.median_ids;
}

}  // namespace

extern "C" int rs_raster_padded_channels(int D) { return padded_channels(D); }

extern "C" int rs_pack_geom(const float* means2d, const float* conics, const float* opacities, int opac_per_cam,
                            const float* compensations, int C, int N, const float* ray_ts, const float* ray_planes,
                            const float* normals, const int32_t* radii, float* geom, int flags, void* stream) {
  RsSpan span__("rs_pack_geom", stream);
  if (C < 0 || N < 0) return RS_ERR_BAD_ARG;
  const long long n_elems = (long long)C * N;
  if (n_elems == 0) return RS_OK;
  if (!means2d || !conics || !opacities || !ray_ts || !ray_planes || !normals || !geom) return RS_ERR_BAD_ARG;
  pack_geom_kernel<<<rs_div_up(n_elems, 256), 256, 0, (cudaStream_t)stream>>>(
      (const float2*)means2d, conics, opacities, opac_per_cam, compensations, N, ray_ts, (const float2*)ray_planes,
      normals, (const int2*)radii, n_elems, (flags & F_CULL_BBOX) ? 0 : 1, (float4*)geom);
  RS_RETURN_LAST_ERROR();
}

extern "C" int rs_pack_colors(const float* colors, long long rows, int D, int DP, float* out, void* stream) {
  RsSpan span__("rs_pack_colors", stream);
  if (rows < 0 || D <= 0 || DP < D) return RS_ERR_BAD_ARG;
  if (rows == 0) return RS_OK;
  if (!colors || !out) return RS_ERR_BAD_ARG;
  pack_colors_kernel<<<rs_div_up(rows * DP, 256), 256, 0, (cudaStream_t)stream>>>(colors, rows, D, DP, out);
  RS_RETURN_LAST_ERROR();
}

extern "C" int rs_unpack_geom_grad(const float* geom_grad, const float* geom, const float* abs_grad, int C, int N,
                                   const float* opacities, int opac_per_cam, const float* compensations,
                                   float* v_means2d, float* v_means2d_abs, float* v_conics, float* v_opacities,
                                   float* v_compensations, float* v_ray_ts, float* v_ray_planes, float* v_normals,
                                   float* v_colors4, int color_per_cam, int D, void* stream) {
  RsSpan span__("rs_unpack_geom_grad", stream);
  if (C < 0 || N < 0) return RS_ERR_BAD_ARG;
  if ((long long)C * N == 0) return RS_OK;
  if (!geom_grad || !geom || !v_means2d || !v_conics || !v_opacities || !v_ray_ts || !v_ray_planes || !v_normals)
    return RS_ERR_BAD_ARG;
  if (v_means2d_abs && !abs_grad) return RS_ERR_BAD_ARG;
  if (compensations && (!opacities || !v_compensations)) return RS_ERR_BAD_ARG;
  if (v_colors4 && (D < 1 || D > 4)) return RS_ERR_BAD_ARG;
  unpack_geom_grad_kernel<<<rs_div_up(N, 256), 256, 0, (cudaStream_t)stream>>>(
      (const float4*)geom_grad, (const float4*)geom, (const float2*)abs_grad, C, N, opacities, opac_per_cam, compensations,
      (float2*)v_means2d, (float2*)v_means2d_abs, v_conics, v_opacities, v_compensations, v_ray_ts,
      (float2*)v_ray_planes, v_normals, v_colors4, color_per_cam, D);
  RS_RETURN_LAST_ERROR();
}

extern "C" int rs_unpack_colors_grad(const float* color_grad, long long rows, int D, int DP, float* out, void* stream) {
  RsSpan span__("rs_unpack_colors_grad", stream);
  if (rows < 0 || D <= 0 || DP < D) return RS_ERR_BAD_ARG;
  if (rows == 0) return RS_OK;
  if (!color_grad || !out) return RS_ERR_BAD_ARG;
  unpack_colors_grad_kernel<<<rs_div_up(rows * D, 256), 256, 0, (cudaStream_t)stream>>>(color_grad, rows, D, DP, out);
  RS_RETURN_LAST_ERROR();
}

extern "C" int rs_rasterize_fwd(const float* geom, const float* colors_padded, int color_per_cam, int D,
                                int ed_channel, const float* backgrounds, const float* Ks, int C, int N, int width,
                                int height, int tile_w, int tile_h, const int32_t* tile_offsets,
                                const int32_t* flatten_ids, long long M, float* out_colors, float* out_alphas,
                                float* out_expected_depths, float* out_median_depths, float* out_normals,
                                float* out_transmittance, int32_t* last_ids, int32_t* median_ids, int flags,
                                unsigned long long* stats, const int32_t* n_isects_dev, void* stream) {
  RsSpan span__("rs_rasterize_fwd", stream);
  if (M >= (1ll << 31)) return RS_ERR_UNSUPPORTED;
  RasterArgs a{};
  a.geom = (const float4*)geom; a.colors = colors_padded; a.backgrounds = backgrounds; a.Ks = Ks;
  a.C = C; a.N = N; a.W = width; a.H = height; a.tile_w = tile_w; a.tile_h = tile_h; a.D = D;
  a.color_per_cam = (color_per_cam || C == 1) ? 1 : 0;
  a.ed_channel = ed_channel;
  a.exact_cull = (flags & F_CULL_BBOX) ? 0 : 1;
  a.flags = flags;
  a.offsets = tile_offsets; a.flatten_ids = flatten_ids; a.M = (int)M;
  a.out_colors = out_colors; a.out_alphas = out_alphas; a.out_dexp = out_expected_depths; a.out_dmed = out_median_depths;
  a.out_normals = out_normals; a.out_T = out_transmittance; a.last_ids = last_ids; a.median_ids = median_ids;
  a.stats = stats;
  a.end_dev = n_isects_dev;
  if (!check_common(a) || !out_colors || !out_alphas || !out_expected_depths || !out_median_depths || !out_normals)
    return RS_ERR_BAD_ARG;
  if (ed_channel >= D) return RS_ERR_BAD_ARG;
  if (tile_w != rs_div_up(width, RS_TILE) || tile_h != rs_div_up(height, RS_TILE)) return RS_ERR_BAD_ARG;
  const int DP = padded_channels(D);
  cudaStream_t st = (cudaStream_t)stream;
  switch (DP) {
#define X(v) case v: return launch_fwd<v>(a, st);
    RS_DP_LIST(X)
#undef X
    default: return RS_ERR_UNSUPPORTED;
  }
}

extern "C" int rs_rasterize_bwd(const float* geom, const float* colors_padded, int color_per_cam, int D,
                                int ed_channel, const float* backgrounds, const float* Ks, int C, int N, int width,
                                int height, int tile_w, int tile_h, const int32_t* tile_offsets,
                                const int32_t* flatten_ids, long long M, const float* out_colors,
                                const float* transmittance, const int32_t* last_ids, const int32_t* median_ids,
                                const float* v_colors, const float* v_alphas, const float* v_expected_depths,
                                const float* v_median_depths, const float* v_normals, float* geom_grad,
                                float* color_grad, float* abs_grad, int flags, const int32_t* n_isects_dev,
                                void* stream) {
  RsSpan span__("rs_rasterize_bwd", stream);
  if (M >= (1ll << 31)) return RS_ERR_UNSUPPORTED;
  RasterArgs a{};
  a.geom = (const float4*)geom; a.colors = colors_padded; a.backgrounds = backgrounds; a.Ks = Ks;
  a.C = C; a.N = N; a.W = width; a.H = height; a.tile_w = tile_w; a.tile_h = tile_h; a.D = D;
  a.color_per_cam = (color_per_cam || C == 1) ? 1 : 0;
  a.ed_channel = ed_channel;
  a.exact_cull = (flags & F_CULL_BBOX) ? 0 : 1;
  a.flags = flags;
  a.offsets = tile_offsets; a.flatten_ids = flatten_ids; a.M = (int)M;
  a.out_colors = (float*)out_colors;
  a.out_T = (float*)transmittance; a.last_ids = (int*)last_ids; a.median_ids = (int*)median_ids;
  a.v_colors = v_colors; a.v_alphas = v_alphas; a.v_dexp = v_expected_depths; a.v_dmed = v_median_depths;
  a.v_normals = v_normals; a.geom_grad = geom_grad; a.color_grad = color_grad; a.abs_grad = abs_grad;
  a.end_dev = n_isects_dev;
  const int DP = padded_channels(D);
  if (!check_common(a) || !v_colors || !v_alphas || !v_expected_depths || !v_median_depths || !v_normals ||
      !geom_grad || (DP > 4 && !color_grad) || (ed_channel >= 0 && !out_colors) || ed_channel >= D)
    return RS_ERR_BAD_ARG;
  if (tile_w != rs_div_up(width, RS_TILE) || tile_h != rs_div_up(height, RS_TILE)) return RS_ERR_BAD_ARG;
  if (M == 0) return RS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  switch (DP) {
#define X(v) case v: return launch_bwd<v>(a, st);
    RS_DP_LIST(X)
#undef X
    default: return RS_ERR_UNSUPPORTED;
  }
}
