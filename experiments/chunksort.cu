// Chunked counting sort of the tile intersections: a third, sort-free way to produce what gsplat's
// `isect_tiles(sort=True)` + `isect_offset_encode` produce, bit-identical to the other two
// paths (isect.cu + radix_sort.cu, tilesort.cu).
//
// With the (camera, Gaussian) entries of each camera already in depth order (`order`, one stable u32 argsort per
// camera), the sorted position of the intersection (entry p, tile t) is
//     offsets[cam][t]  +  #{ entries p' < p of the same camera that touch t },
// a per-tile running count.  The entries of a camera are cut into chunks of G consecutive entries (in depth order):
//   chunk_count_kernel   H[chunk][t]  = entries of the chunk touching t          (shared-memory counters, u16 out)
//   chunk_colsum_kernel  tot[cam][t]  = sum over the camera's chunks             -> inclusive scan (scan_kernel) -> M
//   chunk_base_kernel    base[chunk][t] = offsets[cam][t] + sum over earlier chunks; also writes isect_offsets
//   chunk_emit_kernel    every warp of a chunk owns a contiguous sub-range of its entries; per-warp private u16
//                        counters (prefix over the warps first) give each (entry, tile) its stable rank, and the
//                        64-bit key + flatten id are written straight to their final position.
// No global atomics, no key/value ping-pong: 32 B gathered per entry (twice) + 12 B written per intersection +
// 6 B per (chunk, tile), against (8 + 24 * passes + 12 + 8) B per intersection for emit + radix sort + offset encode.
// MEASURED SLOWER than the radix path on B200 (profiles/r01_chunk_ab.txt: config 2, 6.9 M intersections: 0.62 ms
// vs 0.38 ms of kernels): the 8-byte + 4-byte stores go to 32 different sectors per warp instruction (13.8 M
// partial-sector writes, 0.2 ms on their own), which costs more than the two coalesced radix passes it removes.
// Kept as an opt-in pipeline (gsplat.cuda._wrapper.ISECT_PIPELINE = "chunk") with its bit-exactness tests.
#include "common.cuh"

namespace {

typedef unsigned int u32;
typedef unsigned short u16;
typedef unsigned long long u64;

__device__ __forceinline__ bool cs_tile_bbox(float2 m, int2 r, int tile_w, int tile_h, int& xmin, int& ymin, int& xmax,
                                             int& ymax) {   // same op order as isect.cu / the oracle
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

struct ChunkArgs {
  const float2* means2d;   // [C*N]
  const int2* radii;       // [C*N]
  const float* depths;     // [C*N]
  const int* order;        // [C][N]: index n of the i-th entry of camera c in depth order
  int C, N, tile_w, tile_h, tile_bits;
  int G;                   // entries per chunk
  int chunks_per_cam;
};

// per-warp staging of one group of 32 entries
struct WarpStage {
  int pref[33];
  int xmin[32], ymin[32], w[32], e[32];
};

// Walks the intersections of entries [i0, i1) of camera `cam` (positions in depth order) in (entry, tile row-major)
// order, 32 at a time: lane l of a window holds item k = window_first + l.  body(active, active_mask, tile, e, depth_bits)
template <typename Body>
__device__ __forceinline__ void walk_items(const ChunkArgs& a, int cam, int i0, int i1, WarpStage& s, int lane,
                                           Body body) {
  const int T_w = a.tile_w;
  for (int g = i0; g < i1; g += 32) {
    const int i = g + lane;
    int cnt = 0, xmin = 0, ymin = 0, xmax = 0, ymax = 0, e = 0;
    if (i < i1) {
      e = cam * a.N + __ldg(a.order + (size_t)cam * a.N + i);
      if (cs_tile_bbox(__ldg(a.means2d + e), __ldg(a.radii + e), a.tile_w, a.tile_h, xmin, ymin, xmax, ymax))
        cnt = max((xmax - xmin) * (ymax - ymin), 0);
    }
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(RS_FULL_MASK, incl, d);
      if (lane >= d) incl += o;
    }
    const int total = __shfl_sync(RS_FULL_MASK, incl, 31);
    if (total == 0) continue;
    __syncwarp();
    s.pref[lane + 1] = incl;
    if (lane == 0) s.pref[0] = 0;
    s.xmin[lane] = xmin; s.ymin[lane] = ymin; s.w[lane] = xmax - xmin; s.e[lane] = e;
    __syncwarp();
    for (int k0 = 0; k0 < total; k0 += 32) {
      const int k = k0 + lane;
      const bool active = k < total;
      const unsigned am = __ballot_sync(RS_FULL_MASK, active);
      int tile = 0, ee = 0;
      if (active) {
        int lo = 0, hi = 32;   // owner j: pref[j] <= k < pref[j+1]
#pragma unroll
        for (int it = 0; it < 5; ++it) {
          const int mid = (lo + hi) >> 1;
          if (s.pref[mid] <= k) lo = mid; else hi = mid;
        }
        const int r = k - s.pref[lo];
        const int w = s.w[lo];
        const int ry = r / w, rx = r - ry * w;
        tile = (s.ymin[lo] + ry) * T_w + s.xmin[lo] + rx;
        ee = s.e[lo];
      }
      body(active, am, tile, ee);
    }
  }
}

// ---- H[chunk][t]: one CTA per chunk, CTA-wide packed u16 counters
__global__ void __launch_bounds__(256)
chunk_count_kernel(ChunkArgs a, u16* __restrict__ H) {
  extern __shared__ u32 cs_sm[];
  __shared__ WarpStage stage[8];
  const int T = a.tile_w * a.tile_h, words = (T + 1) >> 1;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  for (int i = t; i < words; i += 256) cs_sm[i] = 0;
  __syncthreads();
  const int chunk = blockIdx.x, cam = chunk / a.chunks_per_cam, cc = chunk - cam * a.chunks_per_cam;
  const int g0 = cc * a.G, g1 = min(a.N, g0 + a.G), sub = a.G / 8;
  walk_items(a, cam, min(g1, g0 + warp * sub), min(g1, g0 + (warp + 1) * sub), stage[warp], lane,
             [&](bool active, unsigned, int tile, int) {
               if (active) atomicAdd(cs_sm + (tile >> 1), 1u << (16 * (tile & 1)));
             });
  __syncthreads();
  u32* out = reinterpret_cast<u32*>(H + (size_t)chunk * (2 * words));   // rows padded to an even tile count
  for (int i = t; i < words; i += 256) out[i] = cs_sm[i];
}

// ---- tot[cam*T + t] = sum over the camera's chunks
__global__ void __launch_bounds__(256)
chunk_colsum_kernel(const u16* __restrict__ H, int C, int T, int Tpad, int chunks_per_cam, int* __restrict__ tot) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= (long long)C * T) return;
  const int cam = (int)(i / T), t = (int)(i - (long long)cam * T);
  const u16* col = H + (size_t)cam * chunks_per_cam * Tpad + t;
  int s = 0;
#pragma unroll 8
  for (int c = 0; c < chunks_per_cam; ++c) s += col[(size_t)c * Tpad];
  tot[i] = s;
}

// ---- base[chunk][t] = offsets[cam][t] + sum over earlier chunks of the camera; offsets = exclusive scan of tot
__global__ void __launch_bounds__(256)
chunk_base_kernel(const u16* __restrict__ H, const int* __restrict__ tot, const long long* __restrict__ incl, int C,
                  int T, int Tpad, int chunks_per_cam, u32* __restrict__ base, int* __restrict__ offsets) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= (long long)C * T) return;
  const int cam = (int)(i / T), t = (int)(i - (long long)cam * T);
  u32 run = (u32)(incl[i] - tot[i]);
  offsets[i] = (int)run;
  const u16* col = H + (size_t)cam * chunks_per_cam * Tpad + t;
  u32* bcol = base + (size_t)cam * chunks_per_cam * Tpad + t;
#pragma unroll 8
  for (int c = 0; c < chunks_per_cam; ++c) {
    bcol[(size_t)c * Tpad] = run;
    run += col[(size_t)c * Tpad];
  }
}

// ---- emit: W warps per chunk, warp-private packed u16 counters [W][words]
template <int W>
__global__ void __launch_bounds__(32 * W)
chunk_emit_kernel(ChunkArgs a, const u32* __restrict__ base, long long* __restrict__ isect_ids,
                  int* __restrict__ flatten_ids) {
  extern __shared__ u32 cs_sm[];
  __shared__ WarpStage stage[W];
  const int T = a.tile_w * a.tile_h, words = (T + 1) >> 1, Tpad = 2 * words;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  for (int i = t; i < W * words; i += 32 * W) cs_sm[i] = 0;
  __syncthreads();
  const int chunk = blockIdx.x, cam = chunk / a.chunks_per_cam, cc = chunk - cam * a.chunks_per_cam;
  const int g0 = cc * a.G, g1 = min(a.N, g0 + a.G), sub = a.G / W;
  const int i0 = min(g1, g0 + warp * sub), i1 = min(g1, g0 + (warp + 1) * sub);
  u32* mine = cs_sm + warp * words;
  walk_items(a, cam, i0, i1, stage[warp], lane, [&](bool active, unsigned, int tile, int) {
    if (active) atomicAdd(mine + (tile >> 1), 1u << (16 * (tile & 1)));
  });
  __syncthreads();
  // exclusive prefix over the warps, both packed halves at once (no carry: a chunk holds < 65536 entries)
  for (int i = t; i < words; i += 32 * W) {
    u32 run = 0;
#pragma unroll
    for (int w = 0; w < W; ++w) {
      const u32 c = cs_sm[w * words + i];
      cs_sm[w * words + i] = run;
      run += c;
    }
  }
  __syncthreads();
  const u32* brow = base + (size_t)chunk * Tpad;
  const u64 hi_cam = (u64)cam << (32 + a.tile_bits);
  const u32 lt = rs::lanemask_lt();
  walk_items(a, cam, i0, i1, stage[warp], lane, [&](bool active, unsigned am, int tile, int e) {
    if (!active) return;
    const u32 peers = __match_any_sync(am, tile);
    const int leader = __ffs(peers) - 1;
    const int sh = 16 * (tile & 1);
    u32 old = 0;
    if (lane == leader) old = (atomicAdd(mine + (tile >> 1), (u32)__popc(peers) << sh) >> sh) & 0xffffu;
    old = __shfl_sync(peers, old, leader);
    const u32 pos = __ldg(brow + tile) + old + __popc(peers & lt);
    isect_ids[pos] = (long long)(hi_cam | ((u64)tile << 32) | (u64)__float_as_uint(__ldg(a.depths + e)));
    flatten_ids[pos] = e;
  });
}

static int cs_tile_bits_for(long long n_tiles) {
  int b = 0;
  while (n_tiles > 0) { ++b; n_tiles >>= 1; }
  return b;
}

}  // namespace

// Entries per chunk for a (C, N) problem: a power of two in [2048, 16384] that keeps about a thousand chunks.
extern "C" int rs_isect_chunk_size(int C, int N) {
  long long want = ((long long)C * N + 1023) / 1024;
  int G = 2048;
  while (G < want && G < 16384) G *= 2;
  return G;
}

// Largest tile count the emit kernel handles (warp-private u16 counters in shared memory): above it use the radix path.
extern "C" int rs_isect_chunk_max_tiles(void) { return 56000; }

// H: u16 [C*chunks_per_cam][Tpad] (Tpad = tiles rounded up to even), tot: i32 [C*T].
extern "C" int rs_isect_chunk_count(const float* means2d, const int32_t* radii, const int32_t* order, int C, int N,
                                    int tile_w, int tile_h, int G, unsigned short* H, int32_t* tot, void* stream) {
  RsSpan span__("rs_isect_chunk_count", stream);
  if (C <= 0 || N <= 0 || tile_w <= 0 || tile_h <= 0 || G < 256 || G > 32768 || (G % 256) != 0) return RS_ERR_BAD_ARG;
  if (!means2d || !radii || !order || !H || !tot) return RS_ERR_BAD_ARG;
  const long long T = (long long)tile_w * tile_h;
  if (T > rs_isect_chunk_max_tiles() || (long long)C * N >= (1ll << 31)) return RS_ERR_UNSUPPORTED;
  const int words = (int)((T + 1) >> 1), cpc = (N + G - 1) / G;
  ChunkArgs a{(const float2*)means2d, (const int2*)radii, nullptr, order, C, N, tile_w, tile_h,
              cs_tile_bits_for(T), G, cpc};
  const size_t smem = sizeof(u32) * (size_t)words;
  if (smem > 32 * 1024) {   // static staging + dynamic counters may pass the 48 KB default together
    cudaError_t e = cudaFuncSetAttribute(chunk_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { rs_set_last_cuda_error((int)e); return RS_ERR_LAUNCH; }
  }
  cudaStream_t st = (cudaStream_t)stream;
  chunk_count_kernel<<<C * cpc, 256, smem, st>>>(a, H);
  chunk_colsum_kernel<<<rs_div_up((long long)C * T, 256), 256, 0, st>>>(H, C, (int)T, 2 * words, cpc, tot);
  rs_count_launches(1);
  RS_RETURN_LAST_ERROR();
}

// incl: inclusive i64 scan of tot (rs_cumsum_i32_i64) -> base u32 [C*chunks_per_cam][Tpad], offsets i32 [C*T].
extern "C" int rs_isect_chunk_base(const unsigned short* H, const int32_t* tot, const long long* incl, int C, int N,
                                   int tile_w, int tile_h, int G, unsigned int* base, int32_t* offsets, void* stream) {
  RsSpan span__("rs_isect_chunk_base", stream);
  if (C <= 0 || N <= 0 || tile_w <= 0 || tile_h <= 0 || G < 256) return RS_ERR_BAD_ARG;
  if (!H || !tot || !incl || !base || !offsets) return RS_ERR_BAD_ARG;
  const long long T = (long long)tile_w * tile_h;
  const int words = (int)((T + 1) >> 1), cpc = (N + G - 1) / G;
  chunk_base_kernel<<<rs_div_up((long long)C * T, 256), 256, 0, (cudaStream_t)stream>>>(
      H, tot, incl, C, (int)T, 2 * words, cpc, base, offsets);
  RS_RETURN_LAST_ERROR();
}

// Writes the M = incl[C*T-1] sorted (key, flatten id) pairs.
extern "C" int rs_isect_chunk_emit(const float* means2d, const int32_t* radii, const float* depths,
                                   const int32_t* order, int C, int N, int tile_w, int tile_h, int G,
                                   const unsigned int* base, long long* isect_ids, int32_t* flatten_ids, void* stream) {
  RsSpan span__("rs_isect_chunk_emit", stream);
  if (C <= 0 || N <= 0 || tile_w <= 0 || tile_h <= 0 || G < 256 || G > 32768 || (G % 256) != 0) return RS_ERR_BAD_ARG;
  if (!means2d || !radii || !depths || !order || !base || !isect_ids || !flatten_ids) return RS_ERR_BAD_ARG;
  const long long T = (long long)tile_w * tile_h;
  if (T > rs_isect_chunk_max_tiles()) return RS_ERR_UNSUPPORTED;
  const int words = (int)((T + 1) >> 1), cpc = (N + G - 1) / G;
  ChunkArgs a{(const float2*)means2d, (const int2*)radii, depths, order, C, N, tile_w, tile_h,
              cs_tile_bits_for(T), G, cpc};
  cudaStream_t st = (cudaStream_t)stream;
  const int W = T <= 14000 ? 8 : (T <= 28000 ? 4 : 2);
  const size_t smem = sizeof(u32) * (size_t)words * W;
#define RS_LAUNCH_EMIT(WW)                                                                                      \
  do {                                                                                                          \
    if (smem > 32 * 1024) {                                                                                     \
      cudaError_t e = cudaFuncSetAttribute(chunk_emit_kernel<WW>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                           (int)smem);                                                          \
      if (e != cudaSuccess) { rs_set_last_cuda_error((int)e); return RS_ERR_LAUNCH; }                           \
    }                                                                                                           \
    chunk_emit_kernel<WW><<<C * cpc, 32 * WW, smem, st>>>(a, base, (long long*)isect_ids, flatten_ids);         \
  } while (0)
  if (W == 8) RS_LAUNCH_EMIT(8); else if (W == 4) RS_LAUNCH_EMIT(4); else RS_LAUNCH_EMIT(2);
#undef RS_LAUNCH_EMIT
  RS_RETURN_LAST_ERROR();
}
