// Tile intersection: per-Gaussian tile counts, an exclusive-scan, 64-bit tile|depth key emission and
// per-tile offset encoding.  Replaces gsplat `isect_tiles` (count + emit passes) and
// `isect_offset_encode`: SURVEY.md rows a7/a9, Appendix A7.  Integer outputs are bit-exact with the
// oracle given the same means2d / radii / depths.
//
// All kernels are HBM-bound streams: count 16 B in + 4 B out per (c,n); scan 4 B in + 8 B out;
// emit 28 B per (c,n) + 12 B per intersection; offsets 8 B per intersection + 4 B per tile.
#include "common.cuh"

namespace {

constexpr int IB = 256;

// tile bounding box [xmin,xmax) x [ymin,ymax) of one projected Gaussian (same op order as the oracle:
// tile_xy = mean/tile, tile_r = radius/tile, floor(tile_xy - tile_r), ceil(tile_xy + tile_r), clamp)
__device__ __forceinline__ bool tile_bbox(float2 m, int2 r, int tile_w, int tile_h, int& xmin, int& ymin, int& xmax,
                                          int& ymax) {
  if (r.x <= 0 || r.y <= 0) return false;
  static_assert((RS_TILE & (RS_TILE - 1)) == 0, "tile size is a power of two: x / tile == x * (1 / tile) bit for bit");
  const float its = 1.0f / (float)RS_TILE;
  float tx = __fmul_rn(m.x, its), ty = __fmul_rn(m.y, its);
  float rx = __fmul_rn((float)r.x, its), ry = __fmul_rn((float)r.y, its);
  float fx0 = floorf(__fsub_rn(tx, rx)), fy0 = floorf(__fsub_rn(ty, ry));
  float fx1 = ceilf(__fadd_rn(tx, rx)), fy1 = ceilf(__fadd_rn(ty, ry));
  xmin = (int)fminf(fmaxf(fx0, 0.f), (float)tile_w);
  ymin = (int)fminf(fmaxf(fy0, 0.f), (float)tile_h);
  xmax = (int)fminf(fmaxf(fx1, 0.f), (float)tile_w);
  ymax = (int)fminf(fmaxf(fy1, 0.f), (float)tile_h);
  return true;
}

__global__ void __launch_bounds__(IB)
isect_count_kernel(const float2* __restrict__ means2d, const int2* __restrict__ radii, long long n_elems, int tile_w,
                   int tile_h, int32_t* __restrict__ tiles_per_gauss) {
  const long long e = (long long)blockIdx.x * IB + threadIdx.x;
  if (e >= n_elems) return;
  int xmin, ymin, xmax, ymax, cnt = 0;
  if (tile_bbox(__ldg(means2d + e), __ldg(radii + e), tile_w, tile_h, xmin, ymin, xmax, ymax))
    cnt = (xmax - xmin) * (ymax - ymin);
  tiles_per_gauss[e] = cnt;
}

// ------------------------------------------------------------------------------------------------
// Single-pass inclusive scan int32 -> int64 with decoupled look-back.  Tile order is the order in
// which blocks draw tickets, so every predecessor a block waits on is already resident.
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = IB * SCAN_ITEMS;
#define ST_AGG (1ull << 62)
#define ST_PREFIX (2ull << 62)
#define ST_MASK (3ull << 62)

// `order` (optional): the scanned sequence is in[order[i]] (the presorted path scans tile counts in depth order)
__global__ void __launch_bounds__(IB)
scan_kernel(const int32_t* __restrict__ in, const int32_t* __restrict__ order, long long* __restrict__ out,
            long long n, volatile unsigned long long* status, unsigned int* ticket) {
  __shared__ unsigned int s_tile;
  __shared__ long long s_warp[IB / 32];
  __shared__ long long s_excl;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (t == 0) s_tile = atomicAdd(ticket, 1u);
  __syncthreads();
  const long long tile = s_tile;
  const long long base = tile * SCAN_TILE + (long long)t * SCAN_ITEMS;
  int v[SCAN_ITEMS];
  if (order) {
    int idx[SCAN_ITEMS];
    if (base + SCAN_ITEMS <= n) {
      int4 a = __ldg(reinterpret_cast<const int4*>(order + base));
      int4 b = __ldg(reinterpret_cast<const int4*>(order + base) + 1);
      idx[0] = a.x; idx[1] = a.y; idx[2] = a.z; idx[3] = a.w; idx[4] = b.x; idx[5] = b.y; idx[6] = b.z; idx[7] = b.w;
    } else {
#pragma unroll
      for (int i = 0; i < SCAN_ITEMS; ++i) idx[i] = (base + i < n) ? __ldg(order + base + i) : -1;
    }
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) v[i] = idx[i] >= 0 ? __ldg(in + idx[i]) : 0;
  } else if (base + SCAN_ITEMS <= n) {
    int4 a = __ldg(reinterpret_cast<const int4*>(in + base));
    int4 b = __ldg(reinterpret_cast<const int4*>(in + base) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) v[i] = (base + i < n) ? __ldg(in + base + i) : 0;
  }
  long long tsum = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) tsum += v[i];
  long long inc = tsum;  // inclusive scan of thread sums inside the warp
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    long long o = __shfl_up_sync(RS_FULL_MASK, inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  long long warp_off = 0, tile_total = 0;
#pragma unroll
  for (int w = 0; w < IB / 32; ++w) {
    long long s = s_warp[w];
    if (w < warp) warp_off += s;
    tile_total += s;
  }
  if (warp == 0) {
    if (lane == 0) status[tile] = (tile == 0 ? ST_PREFIX : ST_AGG) | (unsigned long long)tile_total;
    long long excl = 0;
    long long p = tile - 1;  // look back 32 predecessors at a time
    while (p >= 0) {
      long long q = p - lane;
      unsigned long long s;
      unsigned ready;
      do {
        s = (q >= 0) ? status[q] : ST_PREFIX;  // out-of-range lanes act as a zero prefix
        ready = __ballot_sync(RS_FULL_MASK, (s & ST_MASK) != 0ull);
      } while (ready != RS_FULL_MASK);
      unsigned is_prefix = __ballot_sync(RS_FULL_MASK, (s & ST_MASK) == ST_PREFIX);
      // nearest predecessor holding an inclusive prefix; none in this window -> take all 32 aggregates
      int first = is_prefix ? (__ffs(is_prefix) - 1) : 31;
      long long val = (lane <= first) ? (long long)(s & ~ST_MASK) : 0ll;
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) val += __shfl_xor_sync(RS_FULL_MASK, val, d);
      excl += val;
      if (is_prefix) break;
      p -= 32;
    }
    if (lane == 0) {
      if (tile != 0) status[tile] = ST_PREFIX | (unsigned long long)(excl + tile_total);
      s_excl = excl;
    }
  }
  __syncthreads();
  long long run = s_excl + warp_off + (inc - tsum);
  long long o[SCAN_ITEMS];
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) { run += v[i]; o[i] = run; }
  if (base + SCAN_ITEMS <= n) {
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i += 2)
      *reinterpret_cast<longlong2*>(out + base + i) = make_longlong2(o[i], o[i + 1]);
  } else {
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i)
      if (base + i < n) out[base + i] = o[i];
  }
}

__global__ void __launch_bounds__(IB)
isect_emit_kernel(const float2* __restrict__ means2d, const int2* __restrict__ radii,
                  const float* __restrict__ depths, const long long* __restrict__ cum, int C, int N, int tile_w,
                  int tile_h, int tile_bits, long long* __restrict__ isect_ids, int32_t* __restrict__ flatten_ids) {
  const long long e = (long long)blockIdx.x * IB + threadIdx.x;
  if (e >= (long long)C * N) return;
  int xmin, ymin, xmax, ymax;
  if (!tile_bbox(__ldg(means2d + e), __ldg(radii + e), tile_w, tile_h, xmin, ymin, xmax, ymax)) return;
  const int cnt = (xmax - xmin) * (ymax - ymin);
  if (cnt <= 0) return;
  long long pos = __ldg(cum + e) - cnt;
  const long long cam = e / N;
  const unsigned long long hi_cam = (unsigned long long)cam << (32 + tile_bits);
  const unsigned long long dbits = (unsigned long long)__float_as_uint(__ldg(depths + e));
  for (int y = ymin; y < ymax; ++y)
    for (int x = xmin; x < xmax; ++x) {
      unsigned long long tile = (unsigned long long)(y * tile_w + x);
      isect_ids[pos] = (long long)(hi_cam | (tile << 32) | dbits);
      flatten_ids[pos] = (int32_t)e;
      ++pos;
    }
}

// Presorted path: thread p handles the p-th (camera, Gaussian) entry in DEPTH order (`order` = stable argsort of
// the depth bits) and writes its tile keys at cum[p] - count, so the emitted stream is already sorted by
// (depth, flatten id, tile y, tile x) -- exactly what the first four 8-bit passes of the LSD sort over the 64-bit
// keys produce from the (flatten id, y, x) emission order.  Only the (camera | tile) bits remain to be sorted.
// Warp-cooperative: lane l describes entry p = warp_first + l (tile box, key bits, output offset); the warp's
// intersections form ONE contiguous output range, and the lanes then write it item by item (lane k handles items
// k, k + 32, ...), so the 8-byte key and 4-byte id stores are fully coalesced however uneven the tile counts are.
__global__ void __launch_bounds__(IB)
isect_emit_ordered_kernel(const float2* __restrict__ means2d, const int2* __restrict__ radii,
                          const float* __restrict__ depths, const int32_t* __restrict__ order,
                          const long long* __restrict__ cum, int C, int N, int tile_w, int tile_h, int tile_bits,
                          long long* __restrict__ isect_ids, int32_t* __restrict__ flatten_ids, long long capacity,
                          int* __restrict__ overflow, long long* __restrict__ count_mirror = nullptr) {
  constexpr int WARPS = IB / 32;
  __shared__ int s_pref[WARPS][33];
  __shared__ int s_xmin[WARPS][32], s_ymin[WARPS][32], s_w[WARPS][32], s_e[WARPS][32];
  __shared__ unsigned long long s_key[WARPS][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long p = (long long)blockIdx.x * IB + threadIdx.x;
  const long long total = (long long)C * N;
  int cnt = 0, xmin = 0, ymin = 0, xmax = 0, ymax = 0, e = 0;
  long long cum_p = 0;
  if (p < total) {
    e = __ldg(order + p);
    cum_p = __ldg(cum + p);
    if (tile_bbox(__ldg(means2d + e), __ldg(radii + e), tile_w, tile_h, xmin, ymin, xmax, ymax))
      cnt = max((xmax - xmin) * (ymax - ymin), 0);
    // bounded output (capacity-sized buffers, device-side count): the last entry knows the total
    if (p == total - 1) {
      if (overflow && cum_p > capacity) *overflow = 1;
      if (count_mirror) *count_mirror = cum_p;   // (may be mapped host memory: read by the host a step later, no sync)
    }
  }
  int incl = cnt;   // inclusive prefix of the counts inside the warp
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(RS_FULL_MASK, incl, d);
    if (lane >= d) incl += o;
  }
  const int warp_total = __shfl_sync(RS_FULL_MASK, incl, 31);
  if (warp_total == 0) return;
  // output offset of the warp's first item: every lane knows cum_p - incl (they are all equal where p < total)
  const long long base = __shfl_sync(RS_FULL_MASK, cum_p - incl, 0);
  s_pref[warp][lane + 1] = incl;
  if (lane == 0) s_pref[warp][0] = 0;
  s_xmin[warp][lane] = xmin; s_ymin[warp][lane] = ymin; s_w[warp][lane] = xmax - xmin; s_e[warp][lane] = e;
  if (cnt > 0) {
    const unsigned long long cam = (unsigned long long)(e / N);
    s_key[warp][lane] = (cam << (32 + tile_bits)) | (unsigned long long)__float_as_uint(__ldg(depths + e));
  }
  __syncwarp();
  for (int k = lane; k < warp_total; k += 32) {
    int lo = 0, hi = 32;   // owner j: pref[j] <= k < pref[j+1]
#pragma unroll
    for (int it = 0; it < 5; ++it) {
      const int mid = (lo + hi) >> 1;
      if (s_pref[warp][mid] <= k) lo = mid; else hi = mid;
    }
    const int r = k - s_pref[warp][lo];
    const int w = s_w[warp][lo];
    const int ry = r / w, rx = r - ry * w;
    const unsigned long long tile = (unsigned long long)((s_ymin[warp][lo] + ry) * tile_w + s_xmin[warp][lo] + rx);
    if (base + k < capacity) {
      isect_ids[base + k] = (long long)(s_key[warp][lo] | (tile << 32));
      flatten_ids[base + k] = s_e[warp][lo];
    }
  }
}

// Compact variant of the presorted path: the pair that goes through the radix passes is (camera|tile as u32, flatten id)
// -- 8 bytes instead of 12 -- because the depth bits play no part in the remaining sort; the 64-bit keys the API
// exposes are rebuilt after the sort by isect_finish32_kernel.
__global__ void __launch_bounds__(IB)
isect_emit_ordered32_kernel(const float2* __restrict__ means2d, const int2* __restrict__ radii,
                            const int32_t* __restrict__ order, const long long* __restrict__ cum, int C, int N,
                            int tile_w, int tile_h, int tile_bits, unsigned int* __restrict__ keys32,
                            int32_t* __restrict__ flatten_ids) {
  constexpr int WARPS = IB / 32;
  __shared__ int s_pref[WARPS][33];
  __shared__ int s_xmin[WARPS][32], s_ymin[WARPS][32], s_w[WARPS][32], s_e[WARPS][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long p = (long long)blockIdx.x * IB + threadIdx.x;
  const long long total = (long long)C * N;
  int cnt = 0, xmin = 0, ymin = 0, xmax = 0, ymax = 0, e = 0;
  long long cum_p = 0;
  if (p < total) {
    e = __ldg(order + p);
    cum_p = __ldg(cum + p);
    if (tile_bbox(__ldg(means2d + e), __ldg(radii + e), tile_w, tile_h, xmin, ymin, xmax, ymax))
      cnt = max((xmax - xmin) * (ymax - ymin), 0);
  }
  int incl = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(RS_FULL_MASK, incl, d);
    if (lane >= d) incl += o;
  }
  const int warp_total = __shfl_sync(RS_FULL_MASK, incl, 31);
  if (warp_total == 0) return;
  const long long base = __shfl_sync(RS_FULL_MASK, cum_p - incl, 0);
  s_pref[warp][lane + 1] = incl;
  if (lane == 0) s_pref[warp][0] = 0;
  s_xmin[warp][lane] = xmin; s_ymin[warp][lane] = ymin; s_w[warp][lane] = xmax - xmin; s_e[warp][lane] = e;
  __syncwarp();
  for (int k = lane; k < warp_total; k += 32) {
    int lo = 0, hi = 32;   // owner j: pref[j] <= k < pref[j+1]
#pragma unroll
    for (int it = 0; it < 5; ++it) {
      const int mid = (lo + hi) >> 1;
      if (s_pref[warp][mid] <= k) lo = mid; else hi = mid;
    }
    const int r = k - s_pref[warp][lo];
    const int w = s_w[warp][lo];
    const int ry = r / w, rx = r - ry * w;
    const int ee = s_e[warp][lo];
    const unsigned int tile = (unsigned int)((s_ymin[warp][lo] + ry) * tile_w + s_xmin[warp][lo] + rx);
    keys32[base + k] = ((unsigned int)(ee / N) << tile_bits) | tile;
    flatten_ids[base + k] = ee;
  }
}

// sorted (camera|tile, flatten id) pairs -> the API's 64-bit keys (camera|tile << 32 | depth bits) and the tile offsets
__global__ void __launch_bounds__(IB)
isect_finish32_kernel(const unsigned int* __restrict__ keys32, const int32_t* __restrict__ flatten_ids,
                      const float* __restrict__ depths, long long M, int n_tiles, int tile_bits, int total,
                      long long* __restrict__ isect_ids, int32_t* __restrict__ offsets) {
  const long long i = (long long)blockIdx.x * IB + threadIdx.x;
  if (i >= M) return;
  const unsigned int tmask = (1u << tile_bits) - 1u;
  const unsigned int k = __ldg(keys32 + i);
  const unsigned int dbits = __float_as_uint(__ldg(depths + __ldg(flatten_ids + i)));
  isect_ids[i] = (long long)(((unsigned long long)k << 32) | (unsigned long long)dbits);
  const long long cur = (long long)(k >> tile_bits) * n_tiles + (long long)(k & tmask);
  if (i == 0) {
    for (long long t = 0; t <= cur; ++t) offsets[t] = 0;
  } else {
    const unsigned int kp = __ldg(keys32 + i - 1);
    const long long prev = (long long)(kp >> tile_bits) * n_tiles + (long long)(kp & tmask);
    for (long long t = prev + 1; t <= cur; ++t) offsets[t] = (int32_t)i;
  }
  if (i == M - 1) {
    for (long long t = cur + 1; t < total; ++t) offsets[t] = (int32_t)M;
  }
}

__global__ void __launch_bounds__(IB)
offset_encode_kernel(const long long* __restrict__ isect_ids, long long M, const long long* __restrict__ n_dev,
                     int n_tiles, int tile_bits, int total, int32_t* __restrict__ offsets,
                     int32_t* __restrict__ n_clamped) {
  // four consecutive keys per thread (two 16-byte loads in flight) + the key before them: the kernel only reads, so
  // its speed is the number of bytes each thread keeps in flight
  constexpr int KPT = 4;
  const long long th = (long long)blockIdx.x * IB + threadIdx.x;
  if (n_dev) {   // device-side count: M is the capacity the grid was sized for
    M = min(M, __ldg(n_dev));
    if (th == 0 && n_clamped) *n_clamped = (int32_t)M;   // what the compositing kernels read as the end of the last list
    if (M == 0) {
      if (th < total) offsets[th] = 0;   // (the grid covers at least `total` threads in this mode)
      return;
    }
  }
  const long long i0 = th * KPT;
  if (i0 >= M) return;
  const unsigned long long tmask = (1ull << tile_bits) - 1ull;
  unsigned long long key[KPT];
  if (i0 + KPT <= M && (reinterpret_cast<uintptr_t>(isect_ids) & 15) == 0) {
    const ulonglong2 a = __ldg(reinterpret_cast<const ulonglong2*>(isect_ids + i0));
    const ulonglong2 b = __ldg(reinterpret_cast<const ulonglong2*>(isect_ids + i0) + 1);
    key[0] = a.x; key[1] = a.y; key[2] = b.x; key[3] = b.y;
  } else {
#pragma unroll
    for (int j = 0; j < KPT; ++j) key[j] = i0 + j < M ? (unsigned long long)__ldg(isect_ids + i0 + j) : 0ull;
  }
  long long prev = -1;   // linear tile index of the entry before: "-1" makes entry 0 fill offsets[0..cur] with 0
  if (i0 > 0) {
    const unsigned long long kp = (unsigned long long)__ldg(isect_ids + i0 - 1) >> 32;
    prev = (long long)(kp >> tile_bits) * n_tiles + (long long)(kp & tmask);
  }
#pragma unroll
  for (int j = 0; j < KPT; ++j) {
    const long long i = i0 + j;
    if (i >= M) break;
    const unsigned long long k = key[j] >> 32;
    const long long cur = (long long)(k >> tile_bits) * n_tiles + (long long)(k & tmask);
    for (long long t = prev + 1; t <= cur; ++t) offsets[t] = (int32_t)i;
    prev = cur;
    if (i == M - 1) {
      for (long long t = cur + 1; t < total; ++t) offsets[t] = (int32_t)M;
    }
  }
}

}  // namespace

static int tile_bits_for(long long n_tiles) {
  int b = 0;
  while (n_tiles > 0) { ++b; n_tiles >>= 1; }
  return b;  // == Python int.bit_length(), == floor(log2(n))+1
}

extern "C" int rs_tile_bits(int tile_w, int tile_h) { return tile_bits_for((long long)tile_w * tile_h); }

extern "C" int rs_isect_count(const float* means2d, const int32_t* radii, long long n_elems, int tile_w, int tile_h,
                              int32_t* tiles_per_gauss, void* stream) {
  RsSpan span__("rs_isect_count", stream);
  if (n_elems < 0 || tile_w <= 0 || tile_h <= 0) return RS_ERR_BAD_ARG;
  if (n_elems == 0) return RS_OK;
  if (!means2d || !radii || !tiles_per_gauss) return RS_ERR_BAD_ARG;
  isect_count_kernel<<<rs_div_up(n_elems, IB), IB, 0, (cudaStream_t)stream>>>(
      (const float2*)means2d, (const int2*)radii, n_elems, tile_w, tile_h, tiles_per_gauss);
  RS_RETURN_LAST_ERROR();
}

extern "C" long long rs_cumsum_temp_bytes(long long n) {
  long long tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  return 16 + 8 * (tiles > 0 ? tiles : 1);
}

// inclusive prefix sum int32 -> int64; temp is caller-owned scratch of rs_cumsum_temp_bytes(n) bytes
extern "C" int rs_cumsum_i32_i64(const int32_t* in, long long* out, long long n, void* temp, long long temp_bytes,
                                 void* stream) {
  RsSpan span__("rs_cumsum_i32_i64", stream);
  if (n < 0) return RS_ERR_BAD_ARG;
  if (n == 0) return RS_OK;
  if (!in || !out || !temp || temp_bytes < rs_cumsum_temp_bytes(n)) return RS_ERR_BAD_ARG;
  long long tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  cudaError_t e = cudaMemsetAsync(temp, 0, (size_t)rs_cumsum_temp_bytes(n), (cudaStream_t)stream);
  if (e != cudaSuccess) { rs_set_last_cuda_error((int)e); return RS_ERR_LAUNCH; }
  unsigned int* ticket = (unsigned int*)temp;
  unsigned long long* status = (unsigned long long*)((char*)temp + 16);
  scan_kernel<<<(unsigned)tiles, IB, 0, (cudaStream_t)stream>>>(in, nullptr, out, n, status, ticket);
  RS_RETURN_LAST_ERROR();
}

// out[i] = sum_{j <= i} in[order[j]]  (order: a permutation of 0..n-1)
extern "C" int rs_cumsum_gather_i32_i64(const int32_t* in, const int32_t* order, long long* out, long long n,
                                        void* temp, long long temp_bytes, void* stream) {
  RsSpan span__("rs_cumsum_gather_i32_i64", stream);
  if (n < 0) return RS_ERR_BAD_ARG;
  if (n == 0) return RS_OK;
  if (!in || !order || !out || !temp || temp_bytes < rs_cumsum_temp_bytes(n)) return RS_ERR_BAD_ARG;
  long long tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  cudaError_t e = cudaMemsetAsync(temp, 0, (size_t)rs_cumsum_temp_bytes(n), (cudaStream_t)stream);
  if (e != cudaSuccess) { rs_set_last_cuda_error((int)e); return RS_ERR_LAUNCH; }
  unsigned int* ticket = (unsigned int*)temp;
  unsigned long long* status = (unsigned long long*)((char*)temp + 16);
  scan_kernel<<<(unsigned)tiles, IB, 0, (cudaStream_t)stream>>>(in, order, out, n, status, ticket);
  RS_RETURN_LAST_ERROR();
}

extern "C" int rs_isect_emit(const float* means2d, const int32_t* radii, const float* depths,
                             const long long* cum_tiles, int C, int N, int tile_w, int tile_h, long long* isect_ids,
                             int32_t* flatten_ids, void* stream) {
  RsSpan span__("rs_isect_emit", stream);
  if (C < 0 || N < 0 || tile_w <= 0 || tile_h <= 0) return RS_ERR_BAD_ARG;
  if ((long long)C * N >= (1ll << 31)) return RS_ERR_UNSUPPORTED;
  if (C == 0 || N == 0) return RS_OK;
  if (!means2d || !radii || !depths || !cum_tiles || !isect_ids || !flatten_ids) return RS_ERR_BAD_ARG;
  int tile_bits = tile_bits_for((long long)tile_w * tile_h);
  isect_emit_kernel<<<rs_div_up((long long)C * N, IB), IB, 0, (cudaStream_t)stream>>>(
      (const float2*)means2d, (const int2*)radii, depths, cum_tiles, C, N, tile_w, tile_h, tile_bits, isect_ids,
      flatten_ids);
  RS_RETURN_LAST_ERROR();
}

// order[C*N]: stable argsort of the depth bits (rs_argsort_u32); cum_tiles[p]: inclusive sum of the tile counts in
// that order (rs_cumsum_gather_i32_i64).  The emitted pairs only need sorting on key bits [32, end_bit).
extern "C" int rs_isect_emit_ordered(const float* means2d, const int32_t* radii, const float* depths,
                                     const int32_t* order, const long long* cum_tiles, int C, int N, int tile_w,
                                     int tile_h, long long* isect_ids, int32_t* flatten_ids, void* stream) {
  RsSpan span__("rs_isect_emit_ordered", stream);
  if (C < 0 || N < 0 || tile_w <= 0 || tile_h <= 0) return RS_ERR_BAD_ARG;
  if ((long long)C * N >= (1ll << 31)) return RS_ERR_UNSUPPORTED;
  if (C == 0 || N == 0) return RS_OK;
  if (!means2d || !radii || !depths || !order || !cum_tiles || !isect_ids || !flatten_ids) return RS_ERR_BAD_ARG;
  int tile_bits = tile_bits_for((long long)tile_w * tile_h);
  isect_emit_ordered_kernel<<<rs_div_up((long long)C * N, IB), IB, 0, (cudaStream_t)stream>>>(
      (const float2*)means2d, (const int2*)radii, depths, order, cum_tiles, C, N, tile_w, tile_h, tile_bits, isect_ids,
      flatten_ids, 0x7fffffffffffffffll, nullptr);
  RS_RETURN_LAST_ERROR();
}

// Sync-free form: the output buffers hold `capacity` entries (sized from an earlier step); entries past the capacity
// are dropped and *overflow (device int, zeroed by the caller once) is raised instead.  The count itself stays on the
// device: it is the last element of cum_tiles.
extern "C" int rs_isect_emit_ordered_bounded(const float* means2d, const int32_t* radii, const float* depths,
                                             const int32_t* order, const long long* cum_tiles, int C, int N, int tile_w,
                                             int tile_h, long long* isect_ids, int32_t* flatten_ids, long long capacity,
                                             int32_t* overflow, long long* count_mirror, void* stream) {
  RsSpan span__("rs_isect_emit_ordered", stream);
  if (C < 0 || N < 0 || tile_w <= 0 || tile_h <= 0 || capacity <= 0) return RS_ERR_BAD_ARG;
  if ((long long)C * N >= (1ll << 31) || capacity >= (1ll << 31)) return RS_ERR_UNSUPPORTED;
  if (C == 0 || N == 0) return RS_OK;
  if (!means2d || !radii || !depths || !order || !cum_tiles || !isect_ids || !flatten_ids || !overflow)
    return RS_ERR_BAD_ARG;
  int tile_bits = tile_bits_for((long long)tile_w * tile_h);
  isect_emit_ordered_kernel<<<rs_div_up((long long)C * N, IB), IB, 0, (cudaStream_t)stream>>>(
      (const float2*)means2d, (const int2*)radii, depths, order, cum_tiles, C, N, tile_w, tile_h, tile_bits, isect_ids,
      flatten_ids, capacity, overflow, count_mirror);
  RS_RETURN_LAST_ERROR();
}

extern "C" int rs_offset_encode(const long long* isect_ids, long long M, int C, int tile_w, int tile_h,
                                int32_t* offsets, void* stream) {
  RsSpan span__("rs_offset_encode", stream);
  if (M < 0 || C <= 0 || tile_w <= 0 || tile_h <= 0 || !offsets) return RS_ERR_BAD_ARG;
  if (M >= (1ll << 31)) return RS_ERR_UNSUPPORTED;
  long long n_tiles = (long long)tile_w * tile_h;
  long long total = n_tiles * C;
  if (M == 0) {
    cudaError_t e = cudaMemsetAsync(offsets, 0, sizeof(int32_t) * (size_t)total, (cudaStream_t)stream);
    if (e != cudaSuccess) { rs_set_last_cuda_error((int)e); return RS_ERR_LAUNCH; }
    return RS_OK;
  }
  if (!isect_ids) return RS_ERR_BAD_ARG;
  offset_encode_kernel<<<rs_div_up(rs_div_up(M, 4), IB), IB, 0, (cudaStream_t)stream>>>(isect_ids, M, nullptr, (int)n_tiles,
                                                                        tile_bits_for(n_tiles), (int)total, offsets, nullptr);
  RS_RETURN_LAST_ERROR();
}

// Sync-free form: the number of sorted keys is min(*n_isects_dev, capacity), read on the device.
extern "C" int rs_offset_encode_dev(const long long* isect_ids, long long capacity, const long long* n_isects_dev, int C,
                                    int tile_w, int tile_h, int32_t* offsets, int32_t* n_valid, void* stream) {
  RsSpan span__("rs_offset_encode", stream);
  if (capacity <= 0 || C <= 0 || tile_w <= 0 || tile_h <= 0 || !offsets || !isect_ids || !n_isects_dev || !n_valid)
    return RS_ERR_BAD_ARG;
  if (capacity >= (1ll << 31)) return RS_ERR_UNSUPPORTED;
  const long long n_tiles = (long long)tile_w * tile_h, total = n_tiles * C;
  const long long threads = (capacity + 3) / 4 > total ? (capacity + 3) / 4 : total;
  offset_encode_kernel<<<rs_div_up(threads, IB), IB, 0, (cudaStream_t)stream>>>(
      isect_ids, capacity, n_isects_dev, (int)n_tiles, tile_bits_for(n_tiles), (int)total, offsets, n_valid);
  RS_RETURN_LAST_ERROR();
}

// Compact presorted path (see isect_emit_ordered32_kernel): emits (camera|tile u32, flatten id) pairs in depth order.
extern "C" int rs_isect_emit_ordered32(const float* means2d, const int32_t* radii, const int32_t* order,
                                       const long long* cum_tiles, int C, int N, int tile_w, int tile_h,
                                       uint32_t* keys32, int32_t* flatten_ids, void* stream) {
  RsSpan span__("rs_isect_emit_ordered32", stream);
  if (C < 0 || N < 0 || tile_w <= 0 || tile_h <= 0) return RS_ERR_BAD_ARG;
  if ((long long)C * N >= (1ll << 31)) return RS_ERR_UNSUPPORTED;
  if (C == 0 || N == 0) return RS_OK;
  if (!means2d || !radii || !order || !cum_tiles || !keys32 || !flatten_ids) return RS_ERR_BAD_ARG;
  const int tile_bits = tile_bits_for((long long)tile_w * tile_h);
  int cam_bits = 0;
  while ((1 << cam_bits) <= C) ++cam_bits;                 // floor(log2 C) + 1
  if (tile_bits + cam_bits > 32) return RS_ERR_UNSUPPORTED;  // the caller falls back to the 64-bit path
  isect_emit_ordered32_kernel<<<rs_div_up((long long)C * N, IB), IB, 0, (cudaStream_t)stream>>>(
      (const float2*)means2d, (const int2*)radii, order, cum_tiles, C, N, tile_w, tile_h, tile_bits, keys32,
      flatten_ids);
  RS_RETURN_LAST_ERROR();
}

// After the pairs are sorted on the key bits: isect_ids[M] (64-bit keys) and offsets[C*tile_h*tile_w] in one pass.
extern "C" int rs_isect_finish32(const uint32_t* keys32, const int32_t* flatten_ids, const float* depths, long long M,
                                 int C, int tile_w, int tile_h, long long* isect_ids, int32_t* offsets, void* stream) {
  RsSpan span__("rs_isect_finish32", stream);
  if (M < 0 || C <= 0 || tile_w <= 0 || tile_h <= 0 || !offsets) return RS_ERR_BAD_ARG;
  if (M >= (1ll << 31)) return RS_ERR_UNSUPPORTED;
  const long long n_tiles = (long long)tile_w * tile_h, total = n_tiles * C;
  if (M == 0) {
    cudaError_t e = cudaMemsetAsync(offsets, 0, sizeof(int32_t) * (size_t)total, (cudaStream_t)stream);
    if (e != cudaSuccess) { rs_set_last_cuda_error((int)e); return RS_ERR_LAUNCH; }
    return RS_OK;
  }
  if (!keys32 || !flatten_ids || !depths || !isect_ids) return RS_ERR_BAD_ARG;
  isect_finish32_kernel<<<rs_div_up(M, IB), IB, 0, (cudaStream_t)stream>>>(
      keys32, flatten_ids, depths, M, (int)n_tiles, tile_bits_for(n_tiles), (int)total, isect_ids, offsets);
  RS_RETURN_LAST_ERROR();
}
