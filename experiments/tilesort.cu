// Tile-partitioned intersection lists: the fast path behind `rasterization()` for
// isect_tiles(sort=True) + isect_offset_encode.
//
// The reference emits (tile|depth, id) pairs in Gaussian order and then runs a global 46-bit radix sort
// (152 B per intersection).  The same result -- keys ascending by (camera, tile, depth), ties in
// emission order, i.e. by flatten id -- is reached here with far less traffic:
//   1. count:  per-Gaussian tile counts AND a per-tile histogram (one RED per intersection);
//   2. scan:   exclusive scan over the C*tiles histogram = the tile offsets themselves (one small block);
//   3. emit:   every intersection claims a slot inside its tile's segment with an atomic cursor and stores
//              (depth bits << 32 | flatten id) -- 8 B, already partitioned by tile, order inside a tile arbitrary;
//   4. sort:   one CTA per tile sorts its segment in shared memory (bitonic network on the 64-bit
//              (depth, id) keys, which are unique, so the order is deterministic and equals the stable sort's)
//              and writes the final isect_ids / flatten_ids.
// Traffic: 8 B written + 8 B read + 12 B written per intersection (28 B vs 152 B + 12 B emit + 8 B offsets).
// Segments longer than TS_MAX_SEG fall back to the radix path (the host decides from the max tile count).
#include "common.cuh"

namespace {

typedef unsigned long long u64;
constexpr int TB = 256;
constexpr int TS_MAX_SEG = 8192;  // 64 KB of shared memory per CTA

// same arithmetic as isect.cu (and the oracle): tile bbox of one projected Gaussian
__device__ __forceinline__ bool tile_bbox(float2 m, int2 r, int tile_w, int tile_h, int& xmin, int& ymin, int& xmax,
                                          int& ymax) {
  if (r.x <= 0 || r.y <= 0) return false;
  const float ts = (float)RS_TILE;
  float tx = __fdiv_rn(m.x, ts), ty = __fdiv_rn(m.y, ts);
  float rx = __fdiv_rn((float)r.x, ts), ry = __fdiv_rn((float)r.y, ts);
  float fx0 = floorf(__fsub_rn(tx, rx)), fy0 = floorf(__fsub_rn(ty, ry));
  float fx1 = ceilf(__fadd_rn(tx, rx)), fy1 = ceilf(__fadd_rn(ty, ry));
  xmin = (int)fminf(fmaxf(fx0, 0.f), (float)tile_w);
  ymin = (int)fminf(fmaxf(fy0, 0.f), (float)tile_h);
  xmax = (int)fminf(fmaxf(fx1, 0.f), (float)tile_w);
  ymax = (int)fminf(fmaxf(fy1, 0.f), (float)tile_h);
  return true;
}

__global__ void __launch_bounds__(TB)
tile_count_kernel(const float2* __restrict__ means2d, const int2* __restrict__ radii, int C, int N, int tile_w,
                  int tile_h, int32_t* __restrict__ tiles_per_gauss, int32_t* __restrict__ tile_counts) {
  const long long e = (long long)blockIdx.x * TB + threadIdx.x;
  if (e >= (long long)C * N) return;
  int xmin, ymin, xmax, ymax, cnt = 0;
  if (tile_bbox(__ldg(means2d + e), __ldg(radii + e), tile_w, tile_h, xmin, ymin, xmax, ymax)) {
    cnt = (xmax - xmin) * (ymax - ymin);
    int32_t* tc = tile_counts + (e / N) * (long long)tile_w * tile_h;
    for (int y = ymin; y < ymax; ++y)
      for (int x = xmin; x < xmax; ++x) atomicAdd(tc + y * tile_w + x, 1);
  }
  tiles_per_gauss[e] = cnt;
}

// one block: exclusive scan of `n` tile counts -> offsets and cursors; totals = {M, max count}
__global__ void __launch_bounds__(1024)
tile_scan_kernel(const int32_t* __restrict__ counts, int n, int32_t* __restrict__ offsets,
                 int32_t* __restrict__ cursors, long long* __restrict__ totals) {
  __shared__ int s_warp[32];
  __shared__ int s_carry, s_max;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (t == 0) { s_carry = 0; s_max = 0; }
  __syncthreads();
  int vmax = 0;
  for (int base = 0; base < n; base += 1024) {
    const int i = base + t;
    const int v = i < n ? __ldg(counts + i) : 0;
    vmax = max(vmax, v);
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int o = __shfl_up_sync(RS_FULL_MASK, inc, d);
      if (lane >= d) inc += o;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      int w = s_warp[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        int o = __shfl_up_sync(RS_FULL_MASK, w, d);
        if (lane >= d) w += o;
      }
      s_warp[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const int excl = s_carry + (warp > 0 ? s_warp[warp - 1] : 0) + inc - v;
    if (i < n) { offsets[i] = excl; cursors[i] = excl; }
    __syncthreads();
    if (t == 1023) s_carry = excl + v;
    __syncthreads();
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) vmax = max(vmax, __shfl_xor_sync(RS_FULL_MASK, vmax, d));
  if (lane == 0) atomicMax(&s_max, vmax);
  __syncthreads();
  if (t == 0) { totals[0] = s_carry; totals[1] = s_max; }
}

__global__ void __launch_bounds__(TB)
tile_emit_kernel(const float2* __restrict__ means2d, const int2* __restrict__ radii, const float* __restrict__ depths,
                 int C, int N, int tile_w, int tile_h, int32_t* __restrict__ cursors, u64* __restrict__ pairs) {
  const long long e = (long long)blockIdx.x * TB + threadIdx.x;
  if (e >= (long long)C * N) return;
  int xmin, ymin, xmax, ymax;
  if (!tile_bbox(__ldg(means2d + e), __ldg(radii + e), tile_w, tile_h, xmin, ymin, xmax, ymax)) return;
  const u64 v = ((u64)__float_as_uint(__ldg(depths + e)) << 32) | (u64)(unsigned)e;
  int32_t* cur = cursors + (e / N) * (long long)tile_w * tile_h;
  for (int y = ymin; y < ymax; ++y)
    for (int x = xmin; x < xmax; ++x) {
      const int slot = atomicAdd(cur + y * tile_w + x, 1);
      pairs[slot] = v;
    }
}

// one CTA per (camera, tile): bitonic sort of the segment's (depth, id) keys in shared memory
__global__ void __launch_bounds__(TB)
tile_sort_kernel(const u64* __restrict__ pairs, const int32_t* __restrict__ offsets, int n_total_tiles, int M,
                 int n_tiles, int tile_bits, long long* __restrict__ isect_ids, int32_t* __restrict__ flatten_ids) {
  extern __shared__ __align__(16) u64 s_key[];
  const int tile_id = blockIdx.x;
  const int start = __ldg(offsets + tile_id);
  const int end = tile_id + 1 < n_total_tiles ? __ldg(offsets + tile_id + 1) : M;
  const int n = end - start;
  if (n <= 0) return;
  int p2 = 1;
  while (p2 < n) p2 <<= 1;
  const int t = threadIdx.x;
  for (int i = t; i < p2; i += TB) s_key[i] = i < n ? __ldg(pairs + start + i) : ~0ull;
  __syncthreads();
  for (int k = 2; k <= p2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int q = t; q < (p2 >> 1); q += TB) {
        const int i = ((q & ~(j - 1)) << 1) | (q & (j - 1));  // index with bit j clear
        const int l = i | j;
        const u64 a = s_key[i], b = s_key[l];
        const bool up = (i & k) == 0;
        if ((a > b) == up) { s_key[i] = b; s_key[l] = a; }
      }
      __syncthreads();
    }
  }
  const u64 cam = (u64)(tile_id / n_tiles), tile = (u64)(tile_id % n_tiles);
  const u64 hi = (cam << (32 + tile_bits)) | (tile << 32);
  for (int i = t; i < n; i += TB) {
    const u64 v = s_key[i];
    isect_ids[start + i] = (long long)(hi | (v >> 32));
    flatten_ids[start + i] = (int32_t)(unsigned)(v & 0xffffffffull);
  }
}

int tile_bits_for(long long n_tiles) {
  int b = 0;
  while (n_tiles > 0) { ++b; n_tiles >>= 1; }
  return b;
}

}  // namespace

extern "C" int rs_tile_sort_max_segment(void) { return TS_MAX_SEG; }

// tiles_per_gauss[C*N] is written; tile_counts[C*tile_h*tile_w] must be zero-filled by the caller.
extern "C" int rs_isect_tile_count(const float* means2d, const int32_t* radii, int C, int N, int tile_w, int tile_h,
                                   int32_t* tiles_per_gauss, int32_t* tile_counts, void* stream) {
  RsSpan span__("rs_isect_tile_count", stream);
  if (C < 0 || N < 0 || tile_w <= 0 || tile_h <= 0) return RS_ERR_BAD_ARG;
  if ((long long)C * N >= (1ll << 31)) return RS_ERR_UNSUPPORTED;
  if (C == 0 || N == 0) return RS_OK;
  if (!means2d || !radii || !tiles_per_gauss || !tile_counts) return RS_ERR_BAD_ARG;
  tile_count_kernel<<<rs_div_up((long long)C * N, TB), TB, 0, (cudaStream_t)stream>>>(
      (const float2*)means2d, (const int2*)radii, C, N, tile_w, tile_h, tiles_per_gauss, tile_counts);
  RS_RETURN_LAST_ERROR();
}

// offsets (= isect_offsets, exclusive) and cursors (a working copy) get n_total_tiles entries each;
// totals_dev[0] = M (number of intersections), totals_dev[1] = longest tile segment.
extern "C" int rs_isect_tile_scan(const int32_t* tile_counts, int n_total_tiles, int32_t* offsets, int32_t* cursors,
                                  long long* totals_dev, void* stream) {
  RsSpan span__("rs_isect_tile_scan", stream);
  if (n_total_tiles <= 0 || !tile_counts || !offsets || !cursors || !totals_dev) return RS_ERR_BAD_ARG;
  tile_scan_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(tile_counts, n_total_tiles, offsets, cursors, totals_dev);
  RS_RETURN_LAST_ERROR();
}

extern "C" int rs_isect_tile_emit(const float* means2d, const int32_t* radii, const float* depths, int C, int N,
                                  int tile_w, int tile_h, int32_t* cursors, unsigned long long* pairs, void* stream) {
  RsSpan span__("rs_isect_tile_emit", stream);
  if (C < 0 || N < 0 || tile_w <= 0 || tile_h <= 0) return RS_ERR_BAD_ARG;
  if (C == 0 || N == 0) return RS_OK;
  if (!means2d || !radii || !depths || !cursors || !pairs) return RS_ERR_BAD_ARG;
  tile_emit_kernel<<<rs_div_up((long long)C * N, TB), TB, 0, (cudaStream_t)stream>>>(
      (const float2*)means2d, (const int2*)radii, depths, C, N, tile_w, tile_h, cursors, pairs);
  RS_RETURN_LAST_ERROR();
}

// max_segment (host value of totals[1]) must be <= rs_tile_sort_max_segment(); otherwise use rs_sort_pairs.
extern "C" int rs_isect_tile_sort(const unsigned long long* pairs, const int32_t* offsets, int C, int tile_w,
                                  int tile_h, long long M, int max_segment, long long* isect_ids,
                                  int32_t* flatten_ids, void* stream) {
  RsSpan span__("rs_isect_tile_sort", stream);
  if (C <= 0 || tile_w <= 0 || tile_h <= 0 || M < 0 || max_segment < 0) return RS_ERR_BAD_ARG;
  if (max_segment > TS_MAX_SEG || M >= (1ll << 31)) return RS_ERR_UNSUPPORTED;
  if (M == 0) return RS_OK;
  if (!pairs || !offsets || !isect_ids || !flatten_ids) return RS_ERR_BAD_ARG;
  int p2 = 1;
  while (p2 < max_segment) p2 <<= 1;
  const size_t smem = sizeof(u64) * (size_t)p2;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(tile_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { rs_set_last_cuda_error((int)e); return RS_ERR_LAUNCH; }
  }
  const int n_tiles = tile_w * tile_h;
  tile_sort_kernel<<<C * n_tiles, TB, smem, (cudaStream_t)stream>>>(pairs, offsets, C * n_tiles, (int)M, n_tiles,
                                                                   tile_bits_for(n_tiles), isect_ids, flatten_ids);
  RS_RETURN_LAST_ERROR();
}
