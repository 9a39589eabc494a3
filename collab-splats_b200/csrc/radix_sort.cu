// Stable LSD radix sort of (uint64 key, uint32 value) pairs on a bit range, "onesweep" style:
// one histogram pass over the keys for all digit places, then one read + one write of the pairs per
// 8-bit digit with a decoupled look-back across blocks for the global digit offsets.
// Replaces the cub::DeviceRadixSort::SortPairs call inside gsplat `isect_tiles(sort=True)`
// (SURVEY.md row a8).  Stability + ascending order make the result identical, element for element,
// to a stable argsort of the keys, which is what the oracle does.
//
// HBM-bound: (8 + 24 * ceil(bits/8)) B per pair.  Block = 256 threads x 8 keys.  In-block ranking uses
// warp match-any (one counter row per warp, no shared-memory atomics); pairs are reordered through
// shared memory so that global writes are coalesced runs per digit.
#include "common.cuh"

namespace {

typedef unsigned long long u64;
typedef unsigned int u32;

constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int ST = 256;            // threads per block
constexpr int SITEMS = 8;          // pairs per thread: 64 registers, 4 CTAs/SM, 2048-pair tiles (16: 98 registers, slower)
constexpr int LBW = 8;             // look-back polling window: predecessors polled per round trip
constexpr int SWARPS = ST / 32;
constexpr int MAX_PASSES = 8;
#define LB_AGG (1u << 30)
#define LB_PREFIX (2u << 30)
#define LB_FLAGS (3u << 30)
#define LB_VALUE (~LB_FLAGS)

struct PassInfo {
  int shift[MAX_PASSES];
  u32 mask[MAX_PASSES];
  int n;
};

template <typename K>
__global__ void __launch_bounds__(ST)
radix_hist_kernel(const K* __restrict__ keys, long long M, const long long* __restrict__ n_dev, PassInfo pi,
                  u32* __restrict__ ghist) {
  __shared__ u32 sh[MAX_PASSES * RADIX];
  if (n_dev) M = min(M, __ldg(n_dev));   // device-side count: M is the capacity the grid was sized for
  for (int i = threadIdx.x; i < MAX_PASSES * RADIX; i += ST) sh[i] = 0;
  __syncthreads();
  const long long stride = (long long)gridDim.x * ST;
  constexpr int UN = 4;   // keys in flight per thread: the pass only reads, its speed is its memory-level parallelism
  long long i = (long long)blockIdx.x * ST + threadIdx.x;
  for (; i + (UN - 1) * stride < M; i += UN * stride) {
    K k[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) k[u] = __ldg(keys + i + u * stride);
#pragma unroll
    for (int u = 0; u < UN; ++u)
#pragma unroll
      for (int p = 0; p < MAX_PASSES; ++p)
        if (p < pi.n) atomicAdd(&sh[p * RADIX + (u32)((k[u] >> pi.shift[p]) & pi.mask[p])], 1u);
  }
  for (; i < M; i += stride) {
    K k = __ldg(keys + i);
#pragma unroll
    for (int p = 0; p < MAX_PASSES; ++p)
      if (p < pi.n) atomicAdd(&sh[p * RADIX + (u32)((k >> pi.shift[p]) & pi.mask[p])], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < pi.n * RADIX; i += ST)
    if (sh[i]) atomicAdd(ghist + i, sh[i]);
}

// inclusive scan of one u32 per thread across the 256-thread block
__device__ __forceinline__ u32 block_inclusive_scan(u32 v, u32* s_warp, int lane, int warp) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    u32 o = __shfl_up_sync(RS_FULL_MASK, v, d);
    if (lane >= d) v += o;
  }
  if (lane == 31) s_warp[warp] = v;
  __syncthreads();
  u32 off = 0;
#pragma unroll
  for (int w = 0; w < SWARPS; ++w)
    if (w < warp) off += s_warp[w];
  __syncthreads();
  return v + off;
}

// K = u64 (tile|depth keys) or u32 (depth keys of the presorted path); vin == nullptr: the value of pair i is i.
// Round 2 shortened the per-tile dependency chain of the round-1 kernel (-4 %, profiles/r02_sort_ab.txt):
//  * ranking: all match.any of a thread's items are issued back to back, the digit leaders then bump the warp's counter
//    row with one shared-memory atomic each (program order within the warp keeps the rounds ordered) and the bases come
//    back with one shuffle per item: three pipelined groups instead of 8 dependent match -> LDS -> STS -> SHFL chains;
//  * the global digit bases ride in the look-back: tile 0 alone scans the global histogram and publishes
//    base + count as its inclusive prefix, so every other tile gets base + predecessors from the look-back sum
//    (one block scan and two barriers fewer per tile);
//  * keys and values are reordered through separate shared arrays in the same phase (one barrier pair fewer, key and
//    value stores of a row issued together).
template <typename K, int SI, int LB_WINDOW>
__global__ void __launch_bounds__(ST, 4)
radix_scatter_kernel(const K* __restrict__ kin, const u32* __restrict__ vin, K* __restrict__ kout,
                      u32* __restrict__ vout, int M, const long long* __restrict__ n_dev, int shift, u32 mask,
                      const u32* __restrict__ ghist, volatile u32* status, u32* ticket) {
  __shared__ u32 s_cnt[SWARPS][RADIX];
  __shared__ u32 s_bin_start[RADIX];
  __shared__ int s_gbase[RADIX];
  __shared__ u32 s_warp[SWARPS];
  __shared__ u32 s_tile;
  constexpr int STILE = ST * SI;
  __shared__ __align__(16) K s_keys[STILE];
  __shared__ __align__(16) u32 s_vals[STILE];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (t == 0) s_tile = atomicAdd(ticket, 1u);
#pragma unroll
  for (int w = 0; w < SWARPS; ++w) s_cnt[w][t] = 0;
  __syncthreads();
  const int tile = (int)s_tile;
  const int tile_base = tile * STILE;
  if (n_dev) {
    M = (int)min((long long)M, __ldg(n_dev));
    if (tile_base >= M) return;
  }
  const int n_valid = min(STILE, M - tile_base);

  K key[SI];
  u32 val[SI];
#pragma unroll
  for (int r = 0; r < SI; ++r) {
    const int i = warp * (32 * SI) + r * 32 + lane;
    const bool ok = i < n_valid;
    key[r] = ok ? __ldg(kin + tile_base + i) : (K)~(K)0;
    val[r] = ok ? (vin ? __ldg(vin + tile_base + i) : (u32)(tile_base + i)) : 0u;
  }
  // ---- warp-local stable ranks
  u32 pos[SI], peers[SI];
  const u32 lt = rs::lanemask_lt();
#pragma unroll
  for (int r = 0; r < SI; ++r) peers[r] = __match_any_sync(RS_FULL_MASK, (u32)((key[r] >> shift) & mask));
#pragma unroll
  for (int r = 0; r < SI; ++r) {
    pos[r] = 0;
    if ((peers[r] & lt) == 0u)   // lowest lane of the digit group
      pos[r] = atomicAdd(&s_cnt[warp][(u32)((key[r] >> shift) & mask)], (u32)__popc(peers[r]));
  }
#pragma unroll
  for (int r = 0; r < SI; ++r)
    pos[r] = __shfl_sync(RS_FULL_MASK, pos[r], __ffs(peers[r]) - 1) + __popc(peers[r] & lt);
  __syncthreads();
  // ---- per digit (thread t == digit): prefix over warps, block count
  u32 bcount = 0;
#pragma unroll
  for (int w = 0; w < SWARPS; ++w) {
    u32 c = s_cnt[w][t];
    s_cnt[w][t] = bcount;
    bcount += c;
  }
  volatile u32* my_status = status + (size_t)tile * RADIX + t;
  u32 excl_prev = 0;
  if (tile == 0) {   // block-uniform: the global base of every digit, published as part of the first inclusive prefix
    const u32 gh = __ldg(ghist + t);
    excl_prev = block_inclusive_scan(gh, s_warp, lane, warp) - gh;
    *my_status = LB_PREFIX | (excl_prev + bcount);
  } else {
    *my_status = LB_AGG | bcount;
  }
  const u32 bin_incl = block_inclusive_scan(bcount, s_warp, lane, warp);
  s_bin_start[t] = bin_incl - bcount;
  __syncthreads();
#pragma unroll
  for (int r = 0; r < SI; ++r) {
    const u32 d = (u32)((key[r] >> shift) & mask);
    pos[r] += s_bin_start[d] + s_cnt[warp][d];
    s_keys[pos[r]] = key[r];
    s_vals[pos[r]] = val[r];
  }
  if (tile != 0) {
    int p = tile - 1;
    bool found = false;
    while (!found) {
      u32 sv[LB_WINDOW];
#pragma unroll
      for (int i = 0; i < LB_WINDOW; ++i) {
        sv[i] = LB_PREFIX;
        if (p - i >= 0) {
          const u32* ps = const_cast<const u32*>(status) + (size_t)(p - i) * RADIX + t;
          asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(sv[i]) : "l"(ps) : "memory");
        }
      }
#pragma unroll
      for (int i = 0; i < LB_WINDOW; ++i) {
        if (found) break;
        while ((sv[i] & LB_FLAGS) == 0u) {
          const u32* ps = const_cast<const u32*>(status) + (size_t)(p - i) * RADIX + t;
          asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(sv[i]) : "l"(ps) : "memory");
        }
        excl_prev += sv[i] & LB_VALUE;
        found = (sv[i] & LB_FLAGS) == LB_PREFIX;
      }
      p -= LB_WINDOW;
    }
    *my_status = LB_PREFIX | (excl_prev + bcount);
  }
  s_gbase[t] = (int)excl_prev - (int)(bin_incl - bcount);
  __syncthreads();
#pragma unroll
  for (int i = 0; i < SI; ++i) {
    const int p = i * ST + t;
    const K k = s_keys[p];
    const u32 v = s_vals[p];
    const int o = s_gbase[(u32)((k >> shift) & mask)] + p;
    if (p < n_valid) { kout[o] = k; vout[o] = v; }
  }
}

}  // namespace

static long long sort_blocks(long long M, int items) { return (M + ST * items - 1) / (ST * items); }

extern "C" long long rs_sort_pairs_temp_bytes(long long M, int begin_bit, int end_bit) {
  int npass = end_bit > begin_bit ? (end_bit - begin_bit + RADIX_BITS - 1) / RADIX_BITS : 0;
  if (npass > MAX_PASSES) npass = MAX_PASSES;
  long long nb = sort_blocks(M > 0 ? M : 1, SITEMS);
  return (long long)MAX_PASSES * RADIX * 4 + 256 + (long long)npass * nb * RADIX * 4;
}

template <typename K>
static int sort_pairs_impl(K* ka, u32* va, K* kb, u32* vb, bool iota_vals, long long M, int begin_bit, int end_bit,
                           void* temp, long long temp_bytes, void* stream, const long long* n_dev = nullptr) {
  if (M < 0 || begin_bit < 0 || end_bit > (int)(8 * sizeof(K))) return RS_ERR_BAD_ARG;
  if (M >= (1ll << 30)) return RS_ERR_UNSUPPORTED;  // look-back words carry 30-bit counts
  if (M == 0 || end_bit <= begin_bit) return 1;
  const int npass = (end_bit - begin_bit + RADIX_BITS - 1) / RADIX_BITS;
  if (npass > MAX_PASSES) return RS_ERR_BAD_ARG;
  if (!ka || (!va && !iota_vals) || !kb || !vb || !temp || temp_bytes < rs_sort_pairs_temp_bytes(M, begin_bit, end_bit))
    return RS_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  PassInfo pi;
  pi.n = npass;
  // the key bits are split EVENLY over the passes (13 tile bits: 7 + 6, not 8 + 5): fewer bins in the wide pass mean
  // longer runs per bin in every block's output, i.e. fuller sectors on the scattered writes
  {
    int sh = begin_bit;
    for (int p = 0; p < MAX_PASSES; ++p) {
      int nb = 0;
      if (p < npass) nb = (end_bit - sh + (npass - p) - 1) / (npass - p);
      pi.shift[p] = p < npass ? sh : 0;
      pi.mask[p] = p < npass ? ((1u << nb) - 1u) : 0u;
      sh += nb;
    }
  }
  const long long nblocks = sort_blocks(M, SITEMS);
  cudaError_t e = cudaMemsetAsync(temp, 0, (size_t)rs_sort_pairs_temp_bytes(M, begin_bit, end_bit), st);
  if (e != cudaSuccess) { rs_set_last_cuda_error((int)e); return RS_ERR_LAUNCH; }
  u32* ghist = (u32*)temp;
  u32* tickets = (u32*)((char*)temp + MAX_PASSES * RADIX * 4);
  u32* status = (u32*)((char*)temp + MAX_PASSES * RADIX * 4 + 256);
  int hist_blocks = (int)(nblocks < 148 * 8 ? nblocks : 148 * 8);
  const long long status_stride = nblocks * RADIX;
  radix_hist_kernel<K><<<hist_blocks, ST, 0, st>>>(ka, M, n_dev, pi, ghist);
  for (int p = 0; p < npass; ++p) {
    const u32* vin = (p == 0 && iota_vals) ? nullptr : va;
    radix_scatter_kernel<K, SITEMS, LBW><<<(unsigned)nblocks, ST, 0, st>>>(
        ka, vin, kb, vb, (int)M, n_dev, pi.shift[p], pi.mask[p], ghist + p * RADIX,
        status + (size_t)p * status_stride, tickets + p);
    K* tk = ka; ka = kb; kb = tk;
    u32* tv = va; va = vb; vb = tv;
  }
  rs_count_launches(1 + npass);
  e = cudaPeekAtLastError();
  if (e != cudaSuccess) { rs_set_last_cuda_error((int)e); return RS_ERR_LAUNCH; }
  return (npass & 1) ? 0 : 1;
}

// Sorts M pairs by key bits [begin_bit, end_bit), stable, ascending.  Both buffer pairs are clobbered.
// Returns 0 if the sorted pairs are in (keys_b, vals_b), 1 if they are in (keys_a, vals_a), < 0 on error.
extern "C" int rs_sort_pairs(long long* keys_a, int32_t* vals_a, long long* keys_b, int32_t* vals_b, long long M,
                             int begin_bit, int end_bit, void* temp, long long temp_bytes, void* stream) {
  RsSpan span__("rs_sort_pairs", stream);
  return sort_pairs_impl<u64>((u64*)keys_a, (u32*)vals_a, (u64*)keys_b, (u32*)vals_b, false, M, begin_bit, end_bit,
                              temp, temp_bytes, stream);
}

// Sync-free form: sorts the first min(*n_pairs_dev, capacity) pairs; buffers, temp and grids are sized for `capacity`.
extern "C" int rs_sort_pairs_dev(long long* keys_a, int32_t* vals_a, long long* keys_b, int32_t* vals_b,
                                 long long capacity, const long long* n_pairs_dev, int begin_bit, int end_bit, void* temp,
                                 long long temp_bytes, void* stream) {
  RsSpan span__("rs_sort_pairs", stream);
  if (!n_pairs_dev || capacity <= 0) return RS_ERR_BAD_ARG;
  return sort_pairs_impl<u64>((u64*)keys_a, (u32*)vals_a, (u64*)keys_b, (u32*)vals_b, false, capacity, begin_bit,
                              end_bit, temp, temp_bytes, stream, n_pairs_dev);
}

// Same for 32-bit keys (the depth keys of the presorted intersection path).  vals_a's CONTENT is not read: the
// value of pair i is i (an argsort); vals_a is still needed as the second ping-pong buffer.
extern "C" int rs_argsort_u32(uint32_t* keys_a, int32_t* vals_a, uint32_t* keys_b, int32_t* vals_b, long long M,
                              int begin_bit, int end_bit, void* temp, long long temp_bytes, void* stream) {
  RsSpan span__("rs_argsort_u32", stream);
  if (!vals_a) return RS_ERR_BAD_ARG;
  return sort_pairs_impl<u32>((u32*)keys_a, (u32*)vals_a, (u32*)keys_b, (u32*)vals_b, true, M, begin_bit, end_bit,
                              temp, temp_bytes, stream);
}

// (uint32 key, int32 value) pairs, values read from vals_a (the compact intersection pairs, see isect.cu).
extern "C" int rs_sort_pairs_u32(uint32_t* keys_a, int32_t* vals_a, uint32_t* keys_b, int32_t* vals_b, long long M,
                                 int begin_bit, int end_bit, void* temp, long long temp_bytes, void* stream) {
  RsSpan span__("rs_sort_pairs_u32", stream);
  return sort_pairs_impl<u32>((u32*)keys_a, (u32*)vals_a, (u32*)keys_b, (u32*)vals_b, false, M, begin_bit, end_bit,
                              temp, temp_bytes, stream);
}
