/* rade_b200.h -- C ABI of librade_b200.so, the B200 (sm_100a) RaDe-GS rasterizer hot path.
 *
 * This is the drop-in boundary for the path collab-splats reaches through gsplat-rade:
 *   collab_splats/models/rade_gs_model.py:15,20   (imports rasterization, fully_fused_projection)
 *   collab_splats/models/rade_gs_model.py:373-389 (direct fully_fused_projection call, 8-tuple :392-394)
 *   collab_splats/models/rade_gs_model.py:439-465 (rasterization(..., return_depth_normal=True))
 *   collab_splats/models/rade_features_model.py:20,430-434 (spherical_harmonics), :450-476 (rasterization)
 * gsplat-rade binds its CUDA through a torch C++ extension (gsplat/cuda/_wrapper.py -> csrc/); each entry
 * point below replaces one of those torch ops with plain pointers and sizes.  The Python mirror of the
 * gsplat API that calls these (collab-splats_b200/gsplat/) is the reference-side binding; see
 * INTEGRATION.md.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to a dense, contiguous array unless it says "host";
 *   - the caller owns every buffer, including scratch (sizes from the *_temp_bytes functions);
 *   - `stream` is a cudaStream_t passed as void*; nothing synchronises, nothing allocates;
 *   - functions return RS_OK (0) or a negative status; rs_error_string() explains it;
 *   - C cameras, N Gaussians, images W x H, tiles 16x16 (tile_w = ceil(W/16), tile_h = ceil(H/16)),
 *     M tile intersections, D blended colour channels.
 */
#ifndef RADE_B200_H
#define RADE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RS_OK 0
#define RS_ERR_BAD_ARG (-1)
#define RS_ERR_LAUNCH (-2)
#define RS_ERR_UNSUPPORTED (-3)

int rs_version(void);
const char* rs_error_string(int status);
int rs_last_cuda_error(void);
void rs_set_last_cuda_error(int code);
void rs_count_launches(int n);
unsigned long long rs_launch_count(void); /* kernels launched through this library so far */
/* Optional in-situ timing: when enabled every entry point brackets its launches with CUDA events on the launching
 * stream; rs_timing_collect() synchronises them and returns (name[48], total ms, calls) per entry point. */
void rs_timing_enable(int on);
int rs_timing_begin(const char* name, void* stream);
void rs_timing_end(int span, void* stream);
int rs_timing_collect(char* names /* cap x 48 */, float* ms, int* calls, int cap);
/* FP32 FMA throughput probe: executes blocks*256*iters*32 flops; time it with events on `stream`. */
int rs_fma_peak_probe(int blocks, int iters, float* out, void* stream);

/* ---- projection: replaces gsplat-rade fully_fused_projection fwd (packed=False, pinhole).
 * means[N,3] quats[N,4] (wxyz, un-normalised) scales[N,3] viewmats[C,4,4] Ks[C,3,3] ->
 * radii[C,N,2] i32, means2d[C,N,2], depths[C,N], conics[C,N,3], compensations[C,N] (if requested),
 * ray_ts[C,N], ray_planes[C,N,2], normals[C,N,3]; culled entries are zero-filled. */
int rs_project_fwd(const float* means, const float* quats, const float* scales, const float* viewmats,
                   const float* Ks, int C, int N, int width, int height, float eps2d, float near_plane,
                   float far_plane, float radius_clip, int calc_compensations, int32_t* radii, float* means2d,
                   float* depths, float* conics, float* compensations, float* ray_ts, float* ray_planes,
                   float* normals, void* stream);

/* VJP of the above.  Optional inputs (may be NULL = zero): v_depths, v_compensations, v_ray_ts, v_ray_planes,
 * v_normals.  Outputs v_means[N,3], v_quats[N,4], v_scales[N,3] are overwritten (summed over cameras);
 * v_viewmats[C,4,4] is optional (NULL = not wanted). */
int rs_project_bwd(const float* means, const float* quats, const float* scales, const float* viewmats,
                   const float* Ks, int C, int N, int width, int height, float eps2d, float near_plane,
                   float far_plane, float radius_clip, const float* v_means2d, const float* v_depths,
                   const float* v_conics, const float* v_compensations, const float* v_ray_ts,
                   const float* v_ray_planes, const float* v_normals, float* v_means, float* v_quats,
                   float* v_scales, float* v_viewmats, void* stream);

/* ---- spherical harmonics: replaces gsplat spherical_harmonics fwd/bwd (degree <= 3, K <= 16).
 * dirs[n_elems,3], coeffs[n_coeff_rows,K,3] with row = elem % n_coeff_rows (n_coeff_rows == n_elems, or N when the
 * coefficients are shared by all cameras), masks[n_elems] u8 or NULL -> colors[n_elems,3]. */
int rs_sh_fwd(int degree, int K, long long n_elems, long long n_coeff_rows, const float* dirs, const float* coeffs,
              const uint8_t* masks, float* colors, void* stream);
int rs_sh_bwd(int degree, int K, long long n_elems, long long n_coeff_rows, const float* dirs, const float* coeffs,
              const uint8_t* masks, const float* v_colors, float* v_coeffs, float* v_dirs /* NULL ok */,
              void* stream);

/* ---- tile intersection: replaces gsplat isect_tiles (count pass, scan, emit pass) and isect_offset_encode. */
int rs_tile_bits(int tile_w, int tile_h); /* bit_length(tile_w*tile_h) */
int rs_isect_count(const float* means2d, const int32_t* radii, long long n_elems, int tile_w, int tile_h,
                   int32_t* tiles_per_gauss, void* stream);
long long rs_cumsum_temp_bytes(long long n);
int rs_cumsum_i32_i64(const int32_t* in, long long* out_inclusive, long long n, void* temp, long long temp_bytes,
                      void* stream);
/* key = cam << (32+tile_bits) | tile << 32 | bits(depth); value = c*N+n; emission order (c, n, tile row, tile col) */
int rs_isect_emit(const float* means2d, const int32_t* radii, const float* depths, const long long* cum_tiles,
                  int C, int N, int tile_w, int tile_h, long long* isect_ids, int32_t* flatten_ids, void* stream);
int rs_offset_encode(const long long* sorted_isect_ids, long long M, int C, int tile_w, int tile_h,
                     int32_t* offsets /* [C,tile_h,tile_w] */, void* stream);
/* Depth-presorted emission (default inside isect_tiles(sort=True); same sorted output, bit for bit): the LSD sort's
 * four depth passes over the M intersections are replaced by one stable argsort of the C*N depth bits
 * (rs_argsort_u32 -> order), the tile counts are scanned in that order (rs_cumsum_gather_i32_i64) and the keys are
 * emitted in that order (rs_isect_emit_ordered); only key bits [32, end_bit) then remain for rs_sort_pairs. */
int rs_cumsum_gather_i32_i64(const int32_t* in, const int32_t* order, long long* out_inclusive, long long n,
                             void* temp, long long temp_bytes, void* stream);
int rs_isect_emit_ordered(const float* means2d, const int32_t* radii, const float* depths, const int32_t* order,
                          const long long* cum_tiles_in_order, int C, int N, int tile_w, int tile_h,
                          long long* isect_ids, int32_t* flatten_ids, void* stream);

/* ---- radix sort: replaces cub::DeviceRadixSort::SortPairs inside isect_tiles(sort=True).
 * Stable, ascending, on key bits [begin_bit,end_bit).  Clobbers both buffer pairs.
 * Returns 0: result in (keys_b, vals_b); 1: result in (keys_a, vals_a); <0: error. */
long long rs_sort_pairs_temp_bytes(long long M, int begin_bit, int end_bit);
int rs_sort_pairs(long long* keys_a, int32_t* vals_a, long long* keys_b, int32_t* vals_b, long long M,
                  int begin_bit, int end_bit, void* temp, long long temp_bytes, void* stream);
/* Stable argsort of 32-bit keys: same contract and temp size (rs_sort_pairs_temp_bytes), but the value of pair i
 * is i; vals_a is only scratch (the second ping-pong buffer). */
int rs_argsort_u32(uint32_t* keys_a, int32_t* vals_a, uint32_t* keys_b, int32_t* vals_b, long long M, int begin_bit,
                   int end_bit, void* temp, long long temp_bytes, void* stream);

/* ---- compositing: replaces gsplat-rade rasterize_to_pixels fwd/bwd.
 * Inputs are first packed: geom[C*N,16] (rs_pack_geom) and colours padded to DP = rs_raster_padded_channels(D)
 * channels (rs_pack_colors; colours may be [C*N,D] (color_per_cam=1) or [N,D] shared by all cameras). */
int rs_raster_padded_channels(int D); /* -1 if D > 72: split the channels on the host */
/* Per-call compositing options (`flags` of rs_pack_geom / rs_rasterize_fwd / rs_rasterize_bwd; 0 = the defaults).
 * There is no process-global state: a render passes the SAME flags to its pack, forward and backward calls (the
 * torch layer stores them with the saved tensors), so concurrent callers cannot disturb each other.  Every
 * combination gives the same results (all footprint tests are conservative, all reductions fp32-accurate); the
 * non-default ones exist for A/B measurements and tests. */
#define RS_RASTER_CULL_BBOX 0x1      /* footprint test: padded bbox of the alpha >= 1/255 ellipse instead of the exact
                                        ellipse-vs-rectangle test */
#define RS_RASTER_ONE_PIXEL 0x2      /* <= 4 channels: one pixel per lane (8x4 block per warp) instead of two (8x8) */
#define RS_RASTER_NO_COLOR_MMA 0x4   /* >= 20 channels: SIMT colour blend / colour-gradient reduction instead of mma.sync */
#define RS_RASTER_BWD_MMA 0x8        /* <= 4 channels backward: per-Gaussian reduction as an mma.sync contraction of parked
                                        (vis, v_sigma) against per-pixel constants instead of the shuffle tree: 27 % fewer
                                        instructions, but shared-memory-latency bound -- measured 0.96-0.99 ms vs 0.94 ms at
                                        config 2 (profiles/r02_bwd_mma_ab.txt), so not the default */
#define RS_RASTER_FWD_RING 0x10      /* <= 4 channels forward: three staged batches handed over with mbarriers
                                        (cp.async.mbarrier.arrive) instead of one CTA barrier per batch -- the backward's
                                        default (-1.5 %), measured 6 % slower for the forward */
#define RS_RASTER_BWD_BARRIER 0x20   /* <= 4 channels backward: the double-buffered CTA-barrier staging instead of the ring */
#define RS_RASTER_BWD_TUNE(x) (((x) & 0xf) << 8)   /* backward occupancy / batch variant (0 = default), see rasterize.cu */
int rs_pack_geom(const float* means2d, const float* conics,
                 const float* opacities /* [C*N] if opac_per_cam else [N] */, int opac_per_cam,
                 const float* compensations /* [C*N] or NULL: effective opacity = opacity * compensation */,
                 int C, int N, const float* ray_ts, const float* ray_planes, const float* normals,
                 const int32_t* radii /* NULL ok */, float* geom /* [C*N,16] */, int flags, void* stream);
int rs_pack_colors(const float* colors, long long rows, int D, int DP, float* out, void* stream);
/* ed_channel >= 0: that output channel is divided by max(alpha, 1e-10) (the "ED" render modes); -1: none.
 * stats (NULL ok; <= 4 channels only): 4 x u64 device counters, zeroed by the caller; the launch then runs an instrumented
 * kernel that adds {Q pairs a per-pixel loop visits, Qc blended pairs, warp evaluations, blending warp evaluations}. */
int rs_rasterize_fwd(const float* geom, const float* colors_padded, int color_per_cam, int D, int ed_channel,
                     const float* backgrounds /* [C,D] or NULL */, const float* Ks, int C, int N, int width,
                     int height, int tile_w, int tile_h, const int32_t* tile_offsets, const int32_t* flatten_ids,
                     long long M, float* out_colors /* [C,H,W,D] */, float* out_alphas /* [C,H,W] */,
                     float* out_expected_depths, float* out_median_depths, float* out_normals /* [C,H,W,3] */,
                     float* out_transmittance, int32_t* last_ids, int32_t* median_ids, int flags,
                     unsigned long long* stats, const int32_t* n_valid_dev /* NULL, or see below */, void* stream);
/* Gradient record geom_grad[C*N,16] = (S v_sigma*dx, S v_sigma*dy | ga gb gc | S v_sigma | g_ray_t g_rpx g_rpy |
 * gnx gny gnz | colour 0..3), S = sum over blended pixels; rs_unpack_geom_grad turns the three moments into the
 * means2d and opacity gradients (per-Gaussian linear maps with the conic / ray plane / opacity of `geom`);
 * geom_grad, color_grad[rows,DP] (only needed when DP > 4) and abs_grad[C*N,2] (NULL = absgrad off) must be
 * zero-filled by the caller; gradients are accumulated.  out_colors is only read when ed_channel >= 0. */
int rs_rasterize_bwd(const float* geom, const float* colors_padded, int color_per_cam, int D, int ed_channel,
                     const float* backgrounds, const float* Ks, int C, int N, int width, int height, int tile_w,
                     int tile_h, const int32_t* tile_offsets, const int32_t* flatten_ids, long long M,
                     const float* out_colors, const float* transmittance, const int32_t* last_ids,
                     const int32_t* median_ids, const float* v_colors, const float* v_alphas,
                     const float* v_expected_depths, const float* v_median_depths, const float* v_normals,
                     float* geom_grad, float* color_grad, float* abs_grad, int flags,
                     const int32_t* n_valid_dev /* NULL, or see below */, void* stream);

/* ---- sync-free intersections: the number of intersections M never travels to the host.  The caller sizes isect_ids /
 * flatten_ids (and the sort's temp) for a CAPACITY learned from an earlier step, passes the device address of the count
 * (the last element of the inclusive scan) and gets an overflow flag instead of a short buffer: entries past the
 * capacity are dropped and *overflow is raised (check it off the hot path, re-run with a larger capacity).  With these
 * entry points a whole training step contains no device->host read and can be captured in a CUDA graph.
 * rs_offset_encode_dev also writes *n_valid = min(count, capacity) (int32); rs_rasterize_fwd / rs_rasterize_bwd take
 * that word as n_valid_dev (their M is then the capacity). */
int rs_isect_emit_ordered_bounded(const float* means2d, const int32_t* radii, const float* depths, const int32_t* order,
                                  const long long* cum_tiles, int C, int N, int tile_w, int tile_h, long long* isect_ids,
                                  int32_t* flatten_ids, long long capacity, int32_t* overflow,
                                  long long* count_mirror /* NULL, or any device-visible word (e.g. mapped pinned
                                  host memory) that receives the count: lets the host follow it one step late */,
                                  void* stream);
int rs_sort_pairs_dev(long long* keys_a, int32_t* vals_a, long long* keys_b, int32_t* vals_b, long long capacity,
                      const long long* n_pairs_dev, int begin_bit, int end_bit, void* temp, long long temp_bytes,
                      void* stream);
int rs_offset_encode_dev(const long long* isect_ids, long long capacity, const long long* n_isects_dev, int C, int tile_w,
                         int tile_h, int32_t* offsets, int32_t* n_valid, void* stream);
/* Splits the gradient record into the per-input gradients; undoes the opacity * compensation fusion
 * (v_compensations[C,N] = go * opacity, v_opacities = go * compensation, summed over cameras when the
 * opacities are [N]); v_colors4 (NULL ok) receives the colour gradient when the colours have <= 4 channels. */
int rs_unpack_geom_grad(const float* geom_grad, const float* geom /* the rs_pack_geom records */,
                        const float* abs_grad /* NULL ok */, int C, int N,
                        const float* opacities, int opac_per_cam, const float* compensations /* NULL ok */,
                        float* v_means2d, float* v_means2d_abs /* NULL ok */, float* v_conics, float* v_opacities,
                        float* v_compensations /* NULL ok */, float* v_ray_ts, float* v_ray_planes,
                        float* v_normals, float* v_colors4 /* NULL ok */, int color_per_cam, int D, void* stream);
int rs_unpack_colors_grad(const float* color_grad, long long rows, int D, int DP, float* out, void* stream);

/* ---- fused view-dependent colours: replaces, inside rasterization(), inverse(viewmats) -> dirs ->
 * spherical_harmonics(masks = radii > 0) -> +0.5 -> clamp_min(0) -> cat(depth) and its autograd.
 * colors4[C,N,4] = (rgb, depth or 0); coeffs[N,K,3] are shared by all cameras. */
int rs_sh_colors_fwd(int degree, int K, int C, int N, const float* means, const float* coeffs,
                     const float* viewmats, const int32_t* radii, const float* depths /* NULL ok */,
                     float* colors4, void* stream);
int rs_sh_colors_bwd(int degree, int K, int C, int N, const float* means, const float* coeffs,
                     const float* viewmats, const int32_t* radii, const float* v_colors4, int has_depth,
                     float* v_coeffs, float* v_means, float* v_depths /* NULL unless has_depth */, void* stream);

/* ---- camera-sharded multi-GPU form of rs_sh_colors_bwd (Gaussians replicated, cameras sharded).
 * The SH-coefficient gradient of one camera is Y_k(dir) x v_rgb (48 floats carrying 3), so instead of all-reducing
 * 192 B per Gaussian each rank publishes its masked colour gradients (12 B per Gaussian and camera) in a
 * peer-visible `region` and every rank rebuilds the sum over ALL cameras of ALL ranks itself, reading the other
 * ranks' regions over NVLink (regions[] = peer-mapped pointers) or from an all-gathered copy.
 * region layout: [1024 B header: float4 camera position per camera][float vrgb[C][N][3]] rounded up to 16 B, rs_sh_region_bytes(C, N). */
long long rs_sh_region_bytes(int C, int N);
int rs_sh_colors_bwd_local(int degree, int K, int C, int N, const float* means, const float* coeffs,
                           const float* viewmats, const int32_t* radii, const float* v_colors4, int has_depth,
                           void* region, float* v_means, float* v_depths /* NULL unless has_depth */, void* stream);
int rs_sh_coeffs_gather(int degree, int K, int N, const float* means, const void* const* regions /* host array */,
                        const int* cams /* host array */, int n_sources /* <= 16 */, float* v_coeffs, void* stream);

/* ---- peer-visible memory + flag handshake for the above (one process per GPU, CUDA IPC).  rs_peer_alloc /
 * rs_peer_import are the only entry points of the library that allocate or map memory (set-up time only).
 * rs_peer_signal stores `value` into slot my_rank of every rank's flag array (u64[n_ranks]) after all earlier work
 * of `stream`; rs_peer_wait blocks `stream` on the device until every slot of the local array is >= value and
 * raises *timed_out_dev instead of hanging when a peer does not arrive within timeout_ms. */
int rs_peer_alloc(long long bytes, void** ptr);
int rs_peer_free(void* ptr);
int rs_peer_handle_bytes(void);
int rs_peer_export(void* ptr, void* handle_out);
int rs_peer_import(const void* handle, void** ptr);
int rs_peer_unimport(void* ptr);
int rs_peer_copy(void* dst, const void* src, long long bytes, void* stream);
/* the same transfer to n_dst <= 16 peers (dsts: HOST array of peer-mapped device pointers; bytes % 16 == 0) driven by
 * the SMs: one kernel, ctas_per_dst CTAs per destination, 16-byte posted stores over NVLink */
int rs_peer_push(void* const* dsts, int n_dst, const void* src, long long bytes, int ctas_per_dst, void* stream); /* copy engines, peer pointers ok */
/* Two-shot all-reduce (sum, in place) of n_floats fp32 values over the ranks' peer-mapped buffers in one kernel: rank r
 * pulls slice r of every buffer over NVLink, adds them in rank order (bit-identical results on every rank) and pushes
 * the sum into slice r of every buffer.  Replaces the NCCL all-reduce of the small Gaussian-parameter gradients
 * (SURVEY.md 8e).  bufs: HOST array of n_ranks (2, 4 or 8) pointers; n_floats % (4 * n_ranks) == 0.  The caller brackets
 * it with rs_peer_signal / rs_peer_wait on both sides. */
int rs_peer_allreduce(void* const* bufs, int n_ranks, int my_rank, long long n_floats, int ctas, void* stream);
int rs_peer_signal(void* const* flag_arrays_dev, int n_ranks, int my_rank, unsigned long long value, void* stream);
int rs_peer_wait(const void* local_flags, int n_ranks, unsigned long long value, int timeout_ms, int* timed_out_dev,
                 void* stream);

/* ---- per-Gaussian bookkeeping after the backward.
 * rs_densify_stats: what gsplat's DefaultStrategy._update_state accumulates (reached from
 * collab_splats/models/rade_gs_model.py:191-198): grad2d[n] += sum over visible cameras of |(gx*sx, gy*sy)|,
 * count[n] += number of visible cameras, radii_max[n] = max(radii_max[n], max(rx, ry) * inv_extent); visible =
 * both radii > 0.  rs_project_lookup: collab_splats/utils/utils.py:13-40 (project_gaussians) on the device. */
int rs_densify_stats(const float* grads /* [C,N,2] */, const int32_t* radii /* [C,N,2] */, int C, int N, float sx,
                     float sy, float inv_extent, float* grad2d, float* count, float* radii_max /* NULL ok */,
                     void* stream);
int rs_project_lookup(const float* means2d /* [N,2] */, const int32_t* radii /* [N,2] */, int N, int width, int height,
                      long long* proj_flattened /* [N] */, unsigned char* valid_mask /* [N] */, void* stream);

/* ---- small utilities the host layer uses in place of framework kernels.
 * rs_scale_unless_one: in-place x_b *= *scale (device scalar) over up to 8 fp32 buffers (16-byte aligned; bufs and
 * counts are HOST arrays); the kernel returns at once when *scale == 1 -- the upstream gradient of a loss -- without
 * the host ever reading it.  rs_zero_bytes: cudaMemsetAsync on the caller's stream. */
int rs_scale_unless_one(float* const* bufs, const long long* counts, int n_bufs, const float* scale, void* stream);
int rs_zero_bytes(void* ptr, long long bytes, void* stream);

/* ---- compact presorted intersection path (opt-in; measured on par with the 64-bit path): the pairs that go through the radix
 * passes are (camera|tile as u32, flatten id) -- the depth order is already established by the argsort -- and the
 * 64-bit keys + tile offsets are produced together after the sort.  RS_ERR_UNSUPPORTED when camera|tile needs more
 * than 32 bits. */
int rs_isect_emit_ordered32(const float* means2d, const int32_t* radii, const int32_t* order,
                            const long long* cum_tiles, int C, int N, int tile_w, int tile_h, uint32_t* keys32,
                            int32_t* flatten_ids, void* stream);
int rs_sort_pairs_u32(uint32_t* keys_a, int32_t* vals_a, uint32_t* keys_b, int32_t* vals_b, long long M, int begin_bit,
                      int end_bit, void* temp, long long temp_bytes, void* stream);
int rs_isect_finish32(const uint32_t* keys32, const int32_t* flatten_ids, const float* depths, long long M, int C,
                      int tile_w, int tile_h, long long* isect_ids, int32_t* offsets, void* stream);

/* ---- TSDF fusion of rendered frames on the device: replaces, in the meshing exporter
 * collab_splats/utils/mesh.py:1562-1632, the per-frame `.cpu().numpy()` + Open3D ScalableTSDFVolume.integrate (CPU).
 * Volume = hash map (64-bit packed unit coordinates, open addressing) from 16^3-voxel units to slots of a
 * caller-owned pool.  Caller-owned state: keys u64[capacity] (all 0xFF), vals i32[capacity] (all -1), capacity a power
 * of two > max_units; counters i32[4] = {units allocated, units touched by the last frame, overflow flag, -};
 * unit_xyz i32[max_units,3]; stamps i32[capacity] (zero); touched i32[max_units]; tsdf, weight f32[max_units,4096];
 * rgb f32[max_units,4096,3] or NULL -- all zero-initialised.  Voxel (x,y,z) of a unit is element x*256 + y*16 + z.
 * frame ids start at 1 and must increase.  No host synchronisation, no allocation. */
int rs_tsdf_integrate(const float* depth /* [H,W], 0 = none */, const unsigned char* color_u8 /* [H,W,3] or NULL */,
                      const float* color_f32 /* [H,W,3] or NULL */, int width, int height, float fx, float fy,
                      float cx, float cy, const float* extrinsic_3x4 /* host, world->camera */,
                      const float* pose_3x4 /* host, camera->world */, float voxel_length, float sdf_trunc,
                      float depth_trunc, int depth_sampling_stride, int frame_id, unsigned long long* keys,
                      int* vals, long long capacity, int* counters, int max_units, int* unit_xyz, int* stamps,
                      int* touched, float* tsdf, float* weight, float* rgb, void* stream);

/* ---- feature decode + cosine loss of the rade-features model: replaces
 * collab_splats/models/rade_features_model.py:149-189 (decode_features: bilinear resize + TwoLayerMLP + per-branch
 * resize), collab_splats/utils/features.py:408-449 (TwoLayerMLP) and rade_features_model.py:564-582 (weighted cosine
 * loss), forward and backward.  render [H,W,ld] is the rasterizer's colour output, features are its columns
 * ch0..ch0+F-1 (no permuted copy); x [Hm*Wm,F] and h [Hm*Wm,Hd] are caller-owned intermediates; F, Hd <= 128.
 * Outputs of a branch are channel-first [C, Hb*Wb] like the reference's.  All v_* buffers and *loss are ACCUMULATED
 * into (caller zero-fills).  rs_feature_branch with gt == NULL only decodes; with v_h == NULL it is forward-only. */
int rs_feature_hidden_fwd(const float* render, int H, int W, int ld, int ch0, int F, int Hm, int Wm,
                          const float* W1 /* [Hd,F] */, const float* b1 /* [Hd] */, int Hd, float* x, float* h,
                          void* stream);
int rs_feature_branch(const float* h, int Hm, int Wm, int Hd, const float* W2 /* [C,Hd] */, const float* b2 /* [C] */,
                      int C, const float* gt /* [C,Hb*Wb] or NULL */, int Hb, int Wb,
                      float scale /* branch weight * lambda / (Hb*Wb) */, float* loss /* [1] */,
                      float* psum /* [Hb*Wb,3] zero-filled scratch, needed with gt */,
                      float* v_h /* [Hm*Wm,Hd] or NULL */, float* v_W2 /* [C,Hd] */, float* v_b2 /* [C] */,
                      float* decoded /* [C,Hb*Wb] or NULL */, void* stream);
int rs_feature_hidden_bwd(const float* x, const float* h, const float* v_h, int Hm, int Wm, int Hd, int F,
                          const float* W1, float* v_W1 /* [Hd,F] */, float* v_b1 /* [Hd] */,
                          float* v_render /* [H,W,ld] or NULL */, int H, int W, int ld, int ch0,
                          const float* gscale /* device scalar multiplying v_h, or NULL */, void* stream);

/* ---- fused post-render loss: L1 on RGB + RaDe depth-normal consistency for one camera, forward
 * and gradients in one pass.  Replaces collab_splats/utils/camera_utils.py:176-279 (depth_double_to_normal) and
 * collab_splats/models/rade_gs_model.py:202-219,292-307 as run by the training step.
 * `sums`[4], v_exp_depth and v_med_depth must be zero-filled by the caller.
 * On return sums = (weighted L1, weighted expected-depth term, weighted median-depth term, total loss); the
 * written gradients are d(loss)/d(input). */
int rs_rade_loss_fwd_bwd(const float* render /* [H,W,D] */, const float* alphas, const float* exp_depth,
                         const float* med_depth, const float* normals /* [H,W,3] */,
                         const unsigned char* gt_rgb_u8 /* [H,W,3] */, const float* background /* [3] or NULL */,
                         float fx, float fy, int width, int height, int D, float w_l1, float w_exp, float w_med,
                         int use_depth_normal, float* sums, float* v_render, float* v_alphas, float* v_exp_depth,
                         float* v_med_depth, float* v_normals, void* stream);

/* ---- get_outputs epilogue of the rade-gs model, one camera (replaces the torch glue of
 * collab_splats/models/rade_gs_model.py:200-271 after the rasterization call): rgb = clamp(render[:3] + (1 - alpha) *
 * background, 0, 1); depth_im / depth / median_depth / (normals + 1) / 2 with where(alpha > 0, x, max(x)); the two
 * depth-normal error maps 1 - <normals, depth_double_to_normal(...)[k]> (zero when use_depth_normal == 0).
 * maxima: 4 x u32 device scratch, zero-filled by the caller.  depth_channel: the "ED" channel of render (3) or -1.
 * error_maps: [2,H,W].  The backward takes upstream gradients of the outputs (NULL = zero) and writes the gradients of
 * the five inputs; g_exp_depth / g_med_depth must be zero-filled by the caller. */
int rs_rade_outputs_fwd(const float* render, const float* alphas, const float* exp_depth, const float* med_depth,
                        const float* normals, const float* background, float fx, float fy, int width, int height, int D,
                        int depth_channel, int use_depth_normal, uint32_t* maxima, float* rgb, float* depth,
                        float* median_depth, float* depth_im, float* normals_out, float* error_maps, void* stream);
int rs_rade_outputs_bwd(const float* render, const float* alphas, const float* exp_depth, const float* med_depth,
                        const float* normals, const float* background, float fx, float fy, int width, int height, int D,
                        int depth_channel, int use_depth_normal, const float* v_rgb, const float* v_depth,
                        const float* v_median_depth, const float* v_depth_im, const float* v_normals_out,
                        const float* v_error_maps, const float* v_accumulation, float* g_render, float* g_alphas,
                        float* g_exp_depth, float* g_med_depth, float* g_normals, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RADE_B200_H */
