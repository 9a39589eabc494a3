"""Stage-wise operators with the names and argument meaning of ``gsplat.cuda._wrapper``.

collab-splats imports ``fully_fused_projection`` (collab_splats/models/rade_gs_model.py:20, called at
:373-389) and ``spherical_harmonics`` (collab_splats/models/rade_features_model.py:20, called at :430-434)
from this module; ``isect_tiles``, ``isect_offset_encode`` and ``rasterize_to_pixels`` are what
``gsplat.rendering.rasterization`` is made of.

Every operator is a thin ``torch.autograd.Function`` around the C ABI of librade_b200.so
(include/rade_b200.h) -- torch only owns device memory and the stream.  Options of upstream gsplat that
collab-splats never uses (packed=True, sparse_grad, covars, non-pinhole cameras, 2DGS ...) raise
``NotImplementedError``; nothing silently falls back.
"""

from __future__ import annotations

import ctypes
import math
import os
import warnings
from typing import Optional, Tuple

import torch
from torch import Tensor

from radegs_b200 import backend as _be

TILE_SIZE = 16
# How isect_tiles(sort=True) orders the intersections (identical output either way; tests compare them):
#   "presort": argsort the C*N depths once, emit the keys in depth order, radix-sort only the (camera|tile) bits
#   "radix":   emit in (camera, Gaussian, tile) order, radix-sort all 32 + tile_bits + cam_bits key bits
ISECT_SORT_METHOD = "presort"
# Which pipeline rasterization() uses for isect_tiles + isect_offset_encode (identical outputs; tests compare them):
#   "radix": isect_tiles (ISECT_SORT_METHOD) + isect_offset_encode -- the fastest measured on B200, the default
#   "compact": depth argsort, (camera|tile u32, flatten id) pairs through two radix passes, 64-bit keys and offsets
#            rebuilt together after the sort (8-byte instead of 12-byte pairs): the radix passes gain 7 % (they are
#            latency-bound), the rebuild costs more than offset_encode alone -> 0.389 vs 0.378 ms, not the default
#   (a chunked counting sort and a tile-partitioned bitonic sort were measured slower: experiments/)
ISECT_PIPELINE = "radix"
# Compositing options (backend.RS_RASTER_* bits, 0 = defaults) and optional work counters (a 4 x int64 device tensor,
# <= 4 channels) that rasterize_to_pixels passes with every call.  They are read when the forward runs and travel
# with its saved tensors, so the backward of a render always uses the flags its forward used.
RASTER_FLAGS = int(os.environ.get("RADE_RASTER_FLAGS", "0"), 0)   # (the environment variable only seeds the default)
RASTER_STATS = None
# Sync-free intersections (opt-in, for training loops and CUDA-graph capture): when True, `rasterization()` learns the
# number of intersections of a (device, C, N, tile grid) problem on its FIRST call (one device->host read, as always) and
# from then on sizes `isect_ids` / `flatten_ids` for ISECT_HEADROOM x that count and never reads the count back: the
# kernels take it from device memory.  `meta["isect_ids"]` / `meta["flatten_ids"]` then have the CAPACITY's length,
# `meta["n_isects"]` is a device scalar and `meta["isect_overflow"]` a device flag that is raised (check it off the hot
# path with `isect_overflowed()`) if a step ever produced more intersections than the capacity.
SYNC_FREE = False
ISECT_HEADROOM = 1.25
# One step late and without a synchronisation the host still follows the count: the emit kernel mirrors it into a word of
# pinned host memory, which the next render of the same problem reads as plain host memory.  With ISECT_AUTO_GROW the
# capacity then follows the largest count seen (x ISECT_HEADROOM) -- views with more intersections than the first one
# grow the buffers before they overflow, as long as the count does not jump by more than the headroom between two
# consecutive renders -- and a render that did overflow is reported with a warning on the next call.  (Not under CUDA
# graph capture, where Python does not run per replay: check `isect_overflowed()` there.)
ISECT_AUTO_GROW = True
_ISECT_CAPACITY = {}      # (device index, C, N, tile_w, tile_h) -> [capacity, overflow flag (device), count mirror (pinned host)]


def isect_overflowed(reset: bool = True) -> bool:
    """True if any sync-free call since the last check dropped intersections (device->host read: off the hot path).
    The capacities of the problems that overflowed are forgotten, so their next call re-learns them."""
    bad = False
    for key, (cap, flag, _mirror) in list(_ISECT_CAPACITY.items()):
        if int(flag.item()) != 0:
            bad = True
            if reset:
                del _ISECT_CAPACITY[key]
    return bad


# Set by radegs_b200.multiview.ShGradExchange while a camera-sharded multi-GPU step runs: the backward of the
# fused SH colours then publishes per-camera colour gradients instead of producing the coefficient gradient.
SH_GRAD_SINK = None


def _c(t: Optional[Tensor], dtype=torch.float32) -> Optional[Tensor]:
    if t is None:
        return None
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("gsplat (B200 build): tensors must live on a CUDA device; there is no CPU path")


# ------------------------------------------------------------------------------------------------ projection
class _FullyFusedProjection(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means, quats, scales, viewmats, Ks, width, height, eps2d, near_plane, far_plane,
                radius_clip, calc_compensations):
        lib = _be.load()
        C, N = viewmats.shape[0], means.shape[0]
        dev = means.device
        f32 = dict(device=dev, dtype=torch.float32)
        radii = torch.empty(C, N, 2, device=dev, dtype=torch.int32)
        means2d = torch.empty(C, N, 2, **f32)
        depths = torch.empty(C, N, **f32)
        conics = torch.empty(C, N, 3, **f32)
        comps = torch.empty(C, N, **f32) if calc_compensations else None
        ray_ts = torch.empty(C, N, **f32)
        ray_planes = torch.empty(C, N, 2, **f32)
        normals = torch.empty(C, N, 3, **f32)
        with torch.cuda.device(dev):
            _be.check(lib.rs_project_fwd(
                _be.ptr(means), _be.ptr(quats), _be.ptr(scales), _be.ptr(viewmats), _be.ptr(Ks), C, N, width, height,
                eps2d, near_plane, far_plane, radius_clip, int(calc_compensations), _be.ptr(radii), _be.ptr(means2d),
                _be.ptr(depths), _be.ptr(conics), _be.ptr(comps), _be.ptr(ray_ts), _be.ptr(ray_planes),
                _be.ptr(normals), _be.stream_ptr(dev)), "rs_project_fwd")
        ctx.save_for_backward(means, quats, scales, viewmats, Ks)
        ctx.cfg = (width, height, eps2d, near_plane, far_plane, radius_clip, calc_compensations)
        ctx.mark_non_differentiable(radii)
        if comps is None:
            return radii, means2d, depths, conics, ray_ts, ray_planes, normals
        return radii, means2d, depths, conics, comps, ray_ts, ray_planes, normals

    @staticmethod
    def backward(ctx, *grads):
        lib = _be.load()
        means, quats, scales, viewmats, Ks = ctx.saved_tensors
        width, height, eps2d, near_plane, far_plane, radius_clip, calc_comp = ctx.cfg
        if calc_comp:
            _, v_means2d, v_depths, v_conics, v_comps, v_ray_ts, v_ray_planes, v_normals = grads
        else:
            _, v_means2d, v_depths, v_conics, v_ray_ts, v_ray_planes, v_normals = grads
            v_comps = None
        C, N = viewmats.shape[0], means.shape[0]
        dev = means.device
        if v_means2d is None:
            v_means2d = torch.zeros(C, N, 2, device=dev)
        if v_conics is None:
            v_conics = torch.zeros(C, N, 3, device=dev)
        v_means = torch.empty_like(means)
        v_quats = torch.empty_like(quats)
        v_scales = torch.empty_like(scales)
        v_viewmats = torch.empty_like(viewmats) if ctx.needs_input_grad[3] else None
        with torch.cuda.device(dev):
            _be.check(lib.rs_project_bwd(
                _be.ptr(means), _be.ptr(quats), _be.ptr(scales), _be.ptr(viewmats), _be.ptr(Ks), C, N, width, height,
                eps2d, near_plane, far_plane, radius_clip, _be.ptr(_c(v_means2d)), _be.ptr(_c(v_depths)),
                _be.ptr(_c(v_conics)), _be.ptr(_c(v_comps)), _be.ptr(_c(v_ray_ts)), _be.ptr(_c(v_ray_planes)),
                _be.ptr(_c(v_normals)), _be.ptr(v_means), _be.ptr(v_quats), _be.ptr(v_scales), _be.ptr(v_viewmats),
                _be.stream_ptr(dev)), "rs_project_bwd")
        return (v_means, v_quats, v_scales, v_viewmats, None, None, None, None, None, None, None, None)


def fully_fused_projection(
    means: Tensor,                 # [N,3]
    covars: Optional[Tensor],      # must be None (quats/scales path only)
    quats: Optional[Tensor],       # [N,4] wxyz
    scales: Optional[Tensor],      # [N,3]
    viewmats: Tensor,              # [C,4,4]
    Ks: Tensor,                    # [C,3,3]
    width: int,
    height: int,
    eps2d: float = 0.3,
    near_plane: float = 0.01,
    far_plane: float = 1e10,
    radius_clip: float = 0.0,
    packed: bool = False,
    sparse_grad: bool = False,
    calc_compensations: bool = False,
    camera_model: str = "pinhole",
    opacities: Optional[Tensor] = None,
):
    """Same call as gsplat-rade ``fully_fused_projection`` (rade_gs_model.py:373-389).  Returns the 8-tuple
    ``radii [C,N,2] i32, means2d [C,N,2], depths [C,N], conics [C,N,3], compensations [C,N]|None,
    ray_ts [C,N], ray_planes [C,N,2], normals [C,N,3]`` the reference unpacks at rade_gs_model.py:392-394."""
    if covars is not None:
        raise NotImplementedError("covars input is not on the collab-splats path; pass quats and scales")
    if packed or sparse_grad:
        raise NotImplementedError("packed=True / sparse_grad=True are not on the collab-splats path")
    if camera_model != "pinhole":
        raise NotImplementedError("only pinhole cameras are on the collab-splats path")
    if opacities is not None:
        raise NotImplementedError("opacity-aware radii are not used by the reference (rade_gs_model.py:373-389)")
    if quats is None or scales is None:
        raise ValueError("quats and scales are required")
    N, C = means.shape[0], viewmats.shape[0]
    assert means.shape == (N, 3), means.shape
    assert quats.shape == (N, 4), quats.shape
    assert scales.shape == (N, 3), scales.shape
    assert viewmats.shape == (C, 4, 4), viewmats.shape
    assert Ks.shape == (C, 3, 3), Ks.shape
    _need_cuda(means, quats, scales, viewmats, Ks)
    out = _FullyFusedProjection.apply(_c(means), _c(quats), _c(scales), _c(viewmats), _c(Ks), int(width), int(height),
                                      float(eps2d), float(near_plane), float(far_plane), float(radius_clip),
                                      bool(calc_compensations))
    if calc_compensations:
        return out
    radii, means2d, depths, conics, ray_ts, ray_planes, normals = out
    return radii, means2d, depths, conics, None, ray_ts, ray_planes, normals


# ------------------------------------------------------------------------------------------------ SH
class _SphericalHarmonics(torch.autograd.Function):
    @staticmethod
    def forward(ctx, degree, dirs, coeffs, masks):
        lib = _be.load()
        K = coeffs.shape[-2]
        n_elems = dirs.numel() // 3
        n_rows = coeffs.numel() // (K * 3)
        colors = torch.empty(dirs.shape, device=dirs.device, dtype=torch.float32)
        m8 = None if masks is None else masks.to(torch.uint8).contiguous()
        with torch.cuda.device(dirs.device):
            _be.check(lib.rs_sh_fwd(degree, K, n_elems, n_rows, _be.ptr(dirs), _be.ptr(coeffs), _be.ptr(m8),
                                    _be.ptr(colors), _be.stream_ptr(dirs.device)), "rs_sh_fwd")
        ctx.save_for_backward(dirs, coeffs, m8)
        ctx.degree = degree
        return colors

    @staticmethod
    def backward(ctx, v_colors):
        lib = _be.load()
        dirs, coeffs, m8 = ctx.saved_tensors
        K = coeffs.shape[-2]
        n_elems = dirs.numel() // 3
        n_rows = coeffs.numel() // (K * 3)
        v_coeffs = torch.empty_like(coeffs)
        v_dirs = torch.empty_like(dirs) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(dirs.device):
            _be.check(lib.rs_sh_bwd(ctx.degree, K, n_elems, n_rows, _be.ptr(dirs), _be.ptr(coeffs), _be.ptr(m8),
                                    _be.ptr(_c(v_colors)), _be.ptr(v_coeffs), _be.ptr(v_dirs),
                                    _be.stream_ptr(dirs.device)), "rs_sh_bwd")
        return None, v_dirs, v_coeffs, None


def spherical_harmonics(degrees_to_use: int, dirs: Tensor, coeffs: Tensor, masks: Optional[Tensor] = None) -> Tensor:
    """Same call as gsplat ``spherical_harmonics`` (rade_features_model.py:430-434): dirs [...,3],
    coeffs [...,K,3] -> colours [...,3].  Extension: coeffs may drop the leading camera axis
    (dirs [C,N,3], coeffs [N,K,3]) so shared coefficients are not expanded per camera."""
    assert dirs.shape[-1] == 3 and coeffs.shape[-1] == 3, (dirs.shape, coeffs.shape)
    K = coeffs.shape[-2]
    if not 0 <= degrees_to_use <= 3:
        raise NotImplementedError("SH degree must be 0..3 (the reference uses <= 3)")
    assert (degrees_to_use + 1) ** 2 <= K <= 16, (degrees_to_use, K)
    if coeffs.shape[:-2] != dirs.shape[:-1]:
        assert coeffs.dim() == 3 and dirs.dim() == 3 and coeffs.shape[0] == dirs.shape[1], (dirs.shape, coeffs.shape)
    if masks is not None:
        assert masks.shape == dirs.shape[:-1], masks.shape
    _need_cuda(dirs, coeffs, masks)
    return _SphericalHarmonics.apply(int(degrees_to_use), _c(dirs), _c(coeffs), masks)


# ------------------------------------------------------------------------------------------------ isect
@torch.no_grad()
def isect_tiles(means2d: Tensor, radii: Tensor, depths: Tensor, tile_size: int, tile_width: int, tile_height: int,
                sort: bool = True, packed: bool = False, n_cameras: Optional[int] = None,
                camera_ids: Optional[Tensor] = None, gaussian_ids: Optional[Tensor] = None
                ) -> Tuple[Tensor, Tensor, Tensor]:
    """Same call as gsplat ``isect_tiles``: -> tiles_per_gauss [C,N] i32, isect_ids [M] i64 (sorted),
    flatten_ids [M] i32.  One device->host read of M (the output size is data dependent)."""
    if packed:
        raise NotImplementedError("packed=True is not on the collab-splats path")
    if tile_size != TILE_SIZE:
        raise NotImplementedError("tile_size must be 16")
    lib = _be.load()
    C, N = depths.shape
    assert means2d.shape == (C, N, 2) and radii.shape == (C, N, 2), (means2d.shape, radii.shape)
    _need_cuda(means2d, radii, depths)
    dev = means2d.device
    means2d, depths = _c(means2d), _c(depths)
    radii = _c(radii, torch.int32)
    n_elems = C * N
    tiles = torch.empty(C, N, device=dev, dtype=torch.int32)
    cum = torch.empty(max(n_elems, 1), device=dev, dtype=torch.int64)
    tile_bits = lib.rs_tile_bits(tile_width, tile_height)
    # the key layout reserves floor(log2 C) + 1 camera bits (gsplat), but camera ids only reach C - 1: the sort stops at
    # the highest bit that can be set (C = 8: 16 instead of 17 key bits above the depth -> two 8-bit passes, not three)
    cam_bits = max(int(C - 1).bit_length(), 0)
    end_bit = 32 + tile_bits + cam_bits
    presort = sort and ISECT_SORT_METHOD == "presort" and n_elems > 0
    with torch.cuda.device(dev):
        st = _be.stream_ptr(dev)
        _be.check(lib.rs_isect_count(_be.ptr(means2d), _be.ptr(radii), n_elems, tile_width, tile_height,
                                     _be.ptr(tiles), st), "rs_isect_count")
        tb = lib.rs_cumsum_temp_bytes(n_elems)
        temp = torch.empty(tb, device=dev, dtype=torch.uint8)
        order = None
        if presort:
            # stable argsort of the depth bits of the C*N entries: stands in for the LSD sort's four depth passes
            # over the (6.9x more numerous) intersections -- see rs_isect_emit_ordered
            dkeys = depths.clone().view(torch.int32)          # the sort clobbers its key buffers
            dkeys_b = torch.empty_like(dkeys)
            ord_a = torch.empty(n_elems, device=dev, dtype=torch.int32)
            ord_b = torch.empty_like(ord_a)
            sb = lib.rs_sort_pairs_temp_bytes(n_elems, 0, 32)
            stemp = torch.empty(sb, device=dev, dtype=torch.uint8)
            where = _be.check(lib.rs_argsort_u32(_be.ptr(dkeys), _be.ptr(ord_a), _be.ptr(dkeys_b), _be.ptr(ord_b),
                                                 n_elems, 0, 32, _be.ptr(stemp), sb, st), "rs_argsort_u32")
            order = ord_b if where == 0 else ord_a
            _be.check(lib.rs_cumsum_gather_i32_i64(_be.ptr(tiles), _be.ptr(order), _be.ptr(cum), n_elems,
                                                   _be.ptr(temp), tb, st), "rs_cumsum_gather_i32_i64")
        else:
            _be.check(lib.rs_cumsum_i32_i64(_be.ptr(tiles), _be.ptr(cum), n_elems, _be.ptr(temp), tb, st),
                      "rs_cumsum_i32_i64")
        M = int(cum[n_elems - 1].item()) if n_elems > 0 else 0
        ids_a = torch.empty(M, device=dev, dtype=torch.int64)
        flat_a = torch.empty(M, device=dev, dtype=torch.int32)
        if M == 0:
            return tiles, ids_a, flat_a
        if presort:
            _be.check(lib.rs_isect_emit_ordered(_be.ptr(means2d), _be.ptr(radii), _be.ptr(depths), _be.ptr(order),
                                                _be.ptr(cum), C, N, tile_width, tile_height, _be.ptr(ids_a),
                                                _be.ptr(flat_a), st), "rs_isect_emit_ordered")
        else:
            _be.check(lib.rs_isect_emit(_be.ptr(means2d), _be.ptr(radii), _be.ptr(depths), _be.ptr(cum), C, N,
                                        tile_width, tile_height, _be.ptr(ids_a), _be.ptr(flat_a), st), "rs_isect_emit")
        if not sort:
            return tiles, ids_a, flat_a
        begin_bit = 32 if presort else 0
        ids_b = torch.empty_like(ids_a)
        flat_b = torch.empty_like(flat_a)
        sb = lib.rs_sort_pairs_temp_bytes(M, begin_bit, end_bit)
        stemp = torch.empty(sb, device=dev, dtype=torch.uint8)
        where = _be.check(lib.rs_sort_pairs(_be.ptr(ids_a), _be.ptr(flat_a), _be.ptr(ids_b), _be.ptr(flat_b), M,
                                            begin_bit, end_bit, _be.ptr(stemp), sb, st), "rs_sort_pairs")
    return (tiles, ids_b, flat_b) if where == 0 else (tiles, ids_a, flat_a)


def _isect_compact(means2d: Tensor, radii: Tensor, depths: Tensor, tile_width: int, tile_height: int, key_bits: int):
    """Presorted path with compact pairs: count -> stable depth argsort -> scan in depth order -> emit (camera|tile u32,
    flatten id) -> radix sort on the key bits -> 64-bit keys + tile offsets in one pass."""
    lib = _be.load()
    C, N = depths.shape
    assert means2d.shape == (C, N, 2) and radii.shape == (C, N, 2), (means2d.shape, radii.shape)
    _need_cuda(means2d, radii, depths)
    dev = means2d.device
    means2d, depths = _c(means2d), _c(depths)
    radii = _c(radii, torch.int32)
    n_elems = C * N
    tiles = torch.empty(C, N, device=dev, dtype=torch.int32)
    cum = torch.empty(n_elems, device=dev, dtype=torch.int64)
    offsets = torch.empty(C, tile_height, tile_width, device=dev, dtype=torch.int32)
    with torch.cuda.device(dev):
        st = _be.stream_ptr(dev)
        _be.check(lib.rs_isect_count(_be.ptr(means2d), _be.ptr(radii), n_elems, tile_width, tile_height,
                                     _be.ptr(tiles), st), "rs_isect_count")
        dkeys = depths.clone().view(torch.int32)          # the sort clobbers its key buffers
        dkeys_b = torch.empty_like(dkeys)
        ord_a = torch.empty(n_elems, device=dev, dtype=torch.int32)
        ord_b = torch.empty_like(ord_a)
        sb = lib.rs_sort_pairs_temp_bytes(n_elems, 0, 32)
        stemp = torch.empty(sb, device=dev, dtype=torch.uint8)
        where = _be.check(lib.rs_argsort_u32(_be.ptr(dkeys), _be.ptr(ord_a), _be.ptr(dkeys_b), _be.ptr(ord_b),
                                             n_elems, 0, 32, _be.ptr(stemp), sb, st), "rs_argsort_u32")
        order = ord_b if where == 0 else ord_a
        tb = lib.rs_cumsum_temp_bytes(n_elems)
        temp = torch.empty(tb, device=dev, dtype=torch.uint8)
        _be.check(lib.rs_cumsum_gather_i32_i64(_be.ptr(tiles), _be.ptr(order), _be.ptr(cum), n_elems,
                                               _be.ptr(temp), tb, st), "rs_cumsum_gather_i32_i64")
        M = int(cum[n_elems - 1].item())
        ids = torch.empty(M, device=dev, dtype=torch.int64)
        if M == 0:
            flat = torch.empty(0, device=dev, dtype=torch.int32)
            _be.check(lib.rs_isect_finish32(None, None, None, 0, C, tile_width, tile_height, None, _be.ptr(offsets),
                                            st), "rs_isect_finish32")
            return tiles, ids, flat, offsets
        k_a = torch.empty(M, device=dev, dtype=torch.int32)
        f_a = torch.empty(M, device=dev, dtype=torch.int32)
        k_b, f_b = torch.empty_like(k_a), torch.empty_like(f_a)
        _be.check(lib.rs_isect_emit_ordered32(_be.ptr(means2d), _be.ptr(radii), _be.ptr(order), _be.ptr(cum), C, N,
                                              tile_width, tile_height, _be.ptr(k_a), _be.ptr(f_a), st),
                  "rs_isect_emit_ordered32")
        sb = lib.rs_sort_pairs_temp_bytes(M, 0, key_bits)
        stemp = torch.empty(sb, device=dev, dtype=torch.uint8)
        where = _be.check(lib.rs_sort_pairs_u32(_be.ptr(k_a), _be.ptr(f_a), _be.ptr(k_b), _be.ptr(f_b), M, 0, key_bits,
                                                _be.ptr(stemp), sb, st), "rs_sort_pairs_u32")
        keys, flat = (k_b, f_b) if where == 0 else (k_a, f_a)
        _be.check(lib.rs_isect_finish32(_be.ptr(keys), _be.ptr(flat), _be.ptr(depths), M, C, tile_width, tile_height,
                                        _be.ptr(ids), _be.ptr(offsets), st), "rs_isect_finish32")
    return tiles, ids, flat, offsets


@torch.no_grad()
def isect_tiles_and_offsets(means2d: Tensor, radii: Tensor, depths: Tensor, tile_size: int, tile_width: int,
                            tile_height: int, method: Optional[str] = None) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """``isect_tiles(sort=True)`` + ``isect_offset_encode`` in one go.
    -> tiles_per_gauss [C,N] i32, isect_ids [M] i64, flatten_ids [M] i32, isect_offsets [C,TH,TW] i32.

    ``method=None``: the module default ``ISECT_PIPELINE`` ("radix": emit + onesweep radix sort + offset encode).
    ``method="compact"``: the presorted path with 32-bit camera|tile keys through the radix passes (falls back to "radix"
    when camera|tile needs more than 32 bits).  Both are bit-identical (tests check it).  Two further pipelines (a
    tile-partitioned bitonic sort and a chunked counting sort) were measured slower and live under ``experiments/``."""
    if tile_size != TILE_SIZE:
        raise NotImplementedError("tile_size must be 16")
    if method is None:
        method = ISECT_PIPELINE
    if method == "compact":
        C_ = depths.shape[0]
        bits = _be.load().rs_tile_bits(tile_width, tile_height) + int(math.floor(math.log2(C_))) + 1
        if bits > 32 or depths.numel() == 0:
            method = "radix"
        else:
            return _isect_compact(means2d, radii, depths, tile_width, tile_height, bits)
    if method != "radix":
        raise ValueError(f"unknown intersection pipeline {method!r}")
    C = depths.shape[0]
    tiles, ids, flat = isect_tiles(means2d, radii, depths, tile_size, tile_width, tile_height)
    return tiles, ids, flat, isect_offset_encode(ids, C, tile_width, tile_height)


@torch.no_grad()
def isect_tiles_and_offsets_sync_free(means2d: Tensor, radii: Tensor, depths: Tensor, tile_width: int, tile_height: int):
    """`isect_tiles_and_offsets` without the device->host read of the intersection count (see SYNC_FREE).
    -> tiles_per_gauss, isect_ids [capacity], flatten_ids [capacity], isect_offsets, n_isects (device int32 [1]: the
    number of valid list entries = min(count, capacity)), overflow (device i32 flag, raised when the count exceeded the
    capacity), the count itself (device i64 scalar); None if the capacity of this problem has not been learned yet."""
    lib = _be.load()
    C, N = depths.shape
    dev = means2d.device
    key = (dev.index, C, N, tile_width, tile_height)
    if key not in _ISECT_CAPACITY or C * N == 0:
        return None
    entry = _ISECT_CAPACITY[key]
    cap, overflow, mirror = entry
    if ISECT_AUTO_GROW and not torch.cuda.is_current_stream_capturing():
        last = int(mirror[0])                  # the count of an earlier render of this problem: host memory, no sync
        if last > cap:
            warnings.warn(f"sync-free intersections: an earlier render of this problem produced {last} intersections "
                          f"for a capacity of {cap} and was truncated; the capacity has been raised "
                          "(raise gsplat.cuda._wrapper.ISECT_HEADROOM if views differ this much)")
        want = (int(last * ISECT_HEADROOM) + 4096 + 2047) // 2048 * 2048
        if last > 0 and want > cap:
            cap = entry[0] = want
    means2d, depths = _c(means2d), _c(depths)
    radii = _c(radii, torch.int32)
    n_elems = C * N
    tiles = torch.empty(C, N, device=dev, dtype=torch.int32)
    cum = torch.empty(n_elems, device=dev, dtype=torch.int64)
    tile_bits = lib.rs_tile_bits(tile_width, tile_height)
    end_bit = 32 + tile_bits + max(int(C - 1).bit_length(), 0)
    offsets = torch.empty(C, tile_height, tile_width, device=dev, dtype=torch.int32)
    with torch.cuda.device(dev):
        st = _be.stream_ptr(dev)
        _be.check(lib.rs_isect_count(_be.ptr(means2d), _be.ptr(radii), n_elems, tile_width, tile_height,
                                     _be.ptr(tiles), st), "rs_isect_count")
        dkeys = depths.clone().view(torch.int32)          # the sort clobbers its key buffers
        dkeys_b = torch.empty_like(dkeys)
        ord_a = torch.empty(n_elems, device=dev, dtype=torch.int32)
        ord_b = torch.empty_like(ord_a)
        sb = lib.rs_sort_pairs_temp_bytes(n_elems, 0, 32)
        stemp = torch.empty(sb, device=dev, dtype=torch.uint8)
        where = _be.check(lib.rs_argsort_u32(_be.ptr(dkeys), _be.ptr(ord_a), _be.ptr(dkeys_b), _be.ptr(ord_b),
                                             n_elems, 0, 32, _be.ptr(stemp), sb, st), "rs_argsort_u32")
        order = ord_b if where == 0 else ord_a
        tb = lib.rs_cumsum_temp_bytes(n_elems)
        temp = torch.empty(tb, device=dev, dtype=torch.uint8)
        _be.check(lib.rs_cumsum_gather_i32_i64(_be.ptr(tiles), _be.ptr(order), _be.ptr(cum), n_elems,
                                               _be.ptr(temp), tb, st), "rs_cumsum_gather_i32_i64")
        n_isects = cum[n_elems - 1]                        # stays on the device
        ids_a = torch.empty(cap, device=dev, dtype=torch.int64)
        flat_a = torch.empty(cap, device=dev, dtype=torch.int32)
        ids_b, flat_b = torch.empty_like(ids_a), torch.empty_like(flat_a)
        _be.check(lib.rs_isect_emit_ordered_bounded(_be.ptr(means2d), _be.ptr(radii), _be.ptr(depths), _be.ptr(order),
                                                    _be.ptr(cum), C, N, tile_width, tile_height, _be.ptr(ids_a),
                                                    _be.ptr(flat_a), cap, _be.ptr(overflow),
                                                    ctypes.c_void_p(mirror.data_ptr()), st),
                  "rs_isect_emit_ordered_bounded")
        sb = lib.rs_sort_pairs_temp_bytes(cap, 32, end_bit)
        stemp = torch.empty(sb, device=dev, dtype=torch.uint8)
        where = _be.check(lib.rs_sort_pairs_dev(_be.ptr(ids_a), _be.ptr(flat_a), _be.ptr(ids_b), _be.ptr(flat_b), cap,
                                                _be.ptr(n_isects), 32, end_bit, _be.ptr(stemp), sb, st),
                          "rs_sort_pairs_dev")
        ids, flat = (ids_b, flat_b) if where == 0 else (ids_a, flat_a)
        n_valid = torch.empty(1, device=dev, dtype=torch.int32)     # min(count, capacity): what the compositing reads
        _be.check(lib.rs_offset_encode_dev(_be.ptr(ids), cap, _be.ptr(n_isects), C, tile_width, tile_height,
                                           _be.ptr(offsets), _be.ptr(n_valid), st), "rs_offset_encode_dev")
    return tiles, ids, flat, offsets, n_valid, overflow, n_isects


def isect_learn_capacity(device, C: int, N: int, tile_width: int, tile_height: int, n_isects: int):
    """Records the capacity for later sync-free calls of this problem (called by rasterization() after a synchronising
    call while SYNC_FREE is on)."""
    key = (device.index, C, N, tile_width, tile_height)
    cap = max(int(n_isects * ISECT_HEADROOM) + 4096, 4096)
    cap = (cap + 2047) // 2048 * 2048
    _ISECT_CAPACITY[key] = [cap, torch.zeros(1, device=device, dtype=torch.int32),
                            torch.zeros(1, dtype=torch.int64).pin_memory()]


@torch.no_grad()
def isect_offset_encode(isect_ids: Tensor, n_cameras: int, tile_width: int, tile_height: int) -> Tensor:
    """Same call as gsplat ``isect_offset_encode``: -> offsets [C, tile_height, tile_width] i32."""
    lib = _be.load()
    _need_cuda(isect_ids)
    dev = isect_ids.device
    offsets = torch.empty(n_cameras, tile_height, tile_width, device=dev, dtype=torch.int32)
    ids = _c(isect_ids, torch.int64)
    with torch.cuda.device(dev):
        _be.check(lib.rs_offset_encode(_be.ptr(ids) if ids.numel() else None, ids.numel(), n_cameras, tile_width,
                                       tile_height, _be.ptr(offsets), _be.stream_ptr(dev)), "rs_offset_encode")
    return offsets


# ------------------------------------------------------------------------------------------------ SH colours (fused)
class _SHColors(torch.autograd.Function):
    """campos -> dirs -> SH -> +0.5 -> clamp_min(0) -> [rgb | depth] in one kernel each way (csrc/colors.cu)."""

    @staticmethod
    def forward(ctx, degree, means, coeffs, viewmats, radii, depths):
        lib = _be.load()
        C, N, K = viewmats.shape[0], means.shape[0], coeffs.shape[-2]
        colors4 = torch.empty(C, N, 4, device=means.device, dtype=torch.float32)
        with torch.cuda.device(means.device):
            _be.check(lib.rs_sh_colors_fwd(degree, K, C, N, _be.ptr(means), _be.ptr(coeffs), _be.ptr(viewmats),
                                           _be.ptr(radii), _be.ptr(depths), _be.ptr(colors4),
                                           _be.stream_ptr(means.device)), "rs_sh_colors_fwd")
        ctx.save_for_backward(means, coeffs, viewmats, radii)
        ctx.cfg = (degree, depths is not None)
        return colors4

    @staticmethod
    def backward(ctx, v_colors4):
        lib = _be.load()
        means, coeffs, viewmats, radii = ctx.saved_tensors
        degree, has_depth = ctx.cfg
        C, N, K = viewmats.shape[0], means.shape[0], coeffs.shape[-2]
        v_means = torch.empty_like(means)
        v_depths = torch.empty(C, N, device=means.device, dtype=torch.float32) if has_depth else None
        sink = SH_GRAD_SINK
        if sink is not None:
            # camera-sharded multi-GPU step: publish the masked colour gradients; the coefficient gradient (summed
            # over the cameras of ALL ranks) is rebuilt by sink.finish() -- see radegs_b200.multiview.ShGradExchange
            with torch.cuda.device(means.device):
                st = _be.stream_ptr(means.device)
                _be.check(lib.rs_sh_colors_bwd_local(degree, K, C, N, _be.ptr(means), _be.ptr(coeffs),
                                                     _be.ptr(viewmats), _be.ptr(radii), _be.ptr(_c(v_colors4)),
                                                     int(has_depth), sink.local_region_ptr(C, N), _be.ptr(v_means),
                                                     _be.ptr(v_depths), st), "rs_sh_colors_bwd_local")
                sink.published(means, degree, K, st)
            return None, v_means, None, None, None, v_depths
        v_coeffs = torch.empty_like(coeffs)
        with torch.cuda.device(means.device):
            _be.check(lib.rs_sh_colors_bwd(degree, K, C, N, _be.ptr(means), _be.ptr(coeffs), _be.ptr(viewmats),
                                           _be.ptr(radii), _be.ptr(_c(v_colors4)), int(has_depth), _be.ptr(v_coeffs),
                                           _be.ptr(v_means), _be.ptr(v_depths), _be.stream_ptr(means.device)),
                      "rs_sh_colors_bwd")
        return None, v_means, v_coeffs, None, None, v_depths


def sh_colors(degree: int, means: Tensor, coeffs: Tensor, viewmats: Tensor, radii: Tensor,
              depths: Optional[Tensor] = None) -> Tensor:
    """View-dependent colours as `rasterization()` computes them, fused:
    ``clamp_min(spherical_harmonics(degree, means - campos, coeffs, masks=radii>0) + 0.5, 0)`` with the depth
    channel of the RGB+D / RGB+ED modes appended -> [C,N,4] (4th channel 0 when `depths` is None).
    means [N,3], coeffs [N,K,3] (shared by all cameras), viewmats [C,4,4], radii [C,N,2], depths [C,N]."""
    N, K = means.shape[0], coeffs.shape[-2]
    assert coeffs.shape == (N, K, 3) and 0 <= degree <= 3 and (degree + 1) ** 2 <= K <= 16, (coeffs.shape, degree)
    _need_cuda(means, coeffs, viewmats, radii, depths)
    return _SHColors.apply(int(degree), _c(means), _c(coeffs), _c(viewmats), _c(radii, torch.int32), _c(depths))


# ------------------------------------------------------------------------------------------------ compositing
class _RasterizeToPixels(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means2d, conics, colors, opacities, compensations, ray_ts, ray_planes, normals, backgrounds,
                Ks, width, height, isect_offsets, flatten_ids, absgrad, ed_channel, n_isects=None):
        lib = _be.load()
        C, N = means2d.shape[:2]
        dev = means2d.device
        D = colors.shape[-1]
        DP = lib.rs_raster_padded_channels(D)
        if DP < 0:
            raise NotImplementedError(f"{D} colour channels in one pass (max 72): use channel chunks")
        color_per_cam = colors.dim() == 3
        opac_per_cam = opacities.dim() == 2
        rows = C * N if color_per_cam else N
        tile_h, tile_w = isect_offsets.shape[1:]
        M = flatten_ids.numel()
        flags = int(RASTER_FLAGS)
        stats = RASTER_STATS
        f32 = dict(device=dev, dtype=torch.float32)
        geom = torch.empty(C * N, 16, **f32)
        out_colors = torch.empty(C, height, width, D, **f32)
        out_alphas = torch.empty(C, height, width, 1, **f32)
        out_dexp = torch.empty(C, height, width, 1, **f32)
        out_dmed = torch.empty(C, height, width, 1, **f32)
        out_normals = torch.empty(C, height, width, 3, **f32)
        out_T = torch.empty(C, height, width, **f32)
        last_ids = torch.empty(C, height, width, device=dev, dtype=torch.int32)
        median_ids = torch.empty(C, height, width, device=dev, dtype=torch.int32)
        with torch.cuda.device(dev):
            st = _be.stream_ptr(dev)
            _be.check(lib.rs_pack_geom(_be.ptr(means2d), _be.ptr(conics), _be.ptr(opacities), int(opac_per_cam),
                                       _be.ptr(compensations), C, N, _be.ptr(ray_ts), _be.ptr(ray_planes),
                                       _be.ptr(normals), None, _be.ptr(geom), flags, st), "rs_pack_geom")
            if DP == D:
                colors_p = colors
            else:
                colors_p = torch.empty(rows, DP, **f32)
                _be.check(lib.rs_pack_colors(_be.ptr(colors), rows, D, DP, _be.ptr(colors_p), st), "rs_pack_colors")
            _be.check(lib.rs_rasterize_fwd(
                _be.ptr(geom), _be.ptr(colors_p), int(color_per_cam), D, ed_channel, _be.ptr(backgrounds), _be.ptr(Ks),
                C, N, width, height, tile_w, tile_h, _be.ptr(isect_offsets), _be.ptr(flatten_ids) if M else None, M,
                _be.ptr(out_colors), _be.ptr(out_alphas), _be.ptr(out_dexp), _be.ptr(out_dmed), _be.ptr(out_normals),
                _be.ptr(out_T), _be.ptr(last_ids), _be.ptr(median_ids), flags, _be.ptr(stats), _be.ptr(n_isects), st),
                "rs_rasterize_fwd")
        ctx.save_for_backward(geom, colors_p, backgrounds, Ks, isect_offsets, flatten_ids, out_T, last_ids,
                              median_ids, opacities, compensations, out_colors if ed_channel >= 0 else None, n_isects)
        ctx.cfg = (C, N, D, DP, color_per_cam, opac_per_cam, rows, width, height, tile_w, tile_h, M, absgrad,
                   ed_channel, colors.shape, flags)
        ctx.means2d_ref = means2d if absgrad else None
        ctx.mark_non_differentiable(last_ids, median_ids)
        ctx.set_materialize_grads(False)   # missing output gradients arrive as None (handled by z() below)
        return out_colors, out_alphas, out_dexp, out_dmed, out_normals, last_ids, median_ids

    @staticmethod
    def backward(ctx, v_colors, v_alphas, v_dexp, v_dmed, v_normals, _v_last, _v_med):
        lib = _be.load()
        (geom, colors_p, backgrounds, Ks, isect_offsets, flatten_ids, out_T, last_ids, median_ids, opacities,
         compensations, out_colors, n_isects) = ctx.saved_tensors
        (C, N, D, DP, color_per_cam, opac_per_cam, rows, width, height, tile_w, tile_h, M, absgrad, ed_channel,
         colors_shape, flags) = ctx.cfg
        dev = geom.device
        f32 = dict(device=dev, dtype=torch.float32)

        def z(g, shape):
            return torch.zeros(shape, **f32) if g is None else _c(g)

        v_colors = z(v_colors, (C, height, width, D))
        v_alphas = z(v_alphas, (C, height, width, 1))
        v_dexp = z(v_dexp, (C, height, width, 1))
        v_dmed = z(v_dmed, (C, height, width, 1))
        v_normals = z(v_normals, (C, height, width, 3))
        geom_grad = _be.zeros((C * N, 16), dev)               # 64 MB at 1 M Gaussians: driver memset, not a fill kernel
        color_grad = _be.zeros((rows, DP), dev) if DP > 4 else None
        abs_grad = torch.zeros(C * N, 2, **f32) if absgrad else None
        v_means2d = torch.empty(C, N, 2, **f32)
        v_abs = torch.empty(C, N, 2, **f32) if absgrad else None
        v_conics = torch.empty(C, N, 3, **f32)
        v_opac = torch.empty(opacities.shape, **f32)
        v_comps = torch.empty(C, N, **f32) if compensations is not None else None
        v_ray_ts = torch.empty(C, N, **f32)
        v_ray_planes = torch.empty(C, N, 2, **f32)
        v_nrm = torch.empty(C, N, 3, **f32)
        v_col = torch.empty(colors_shape, **f32)
        with torch.cuda.device(dev):
            st = _be.stream_ptr(dev)
            _be.check(lib.rs_rasterize_bwd(
                _be.ptr(geom), _be.ptr(colors_p), int(color_per_cam), D, ed_channel, _be.ptr(backgrounds), _be.ptr(Ks),
                C, N, width, height, tile_w, tile_h, _be.ptr(isect_offsets), _be.ptr(flatten_ids) if M else None, M,
                _be.ptr(out_colors), _be.ptr(out_T), _be.ptr(last_ids), _be.ptr(median_ids), _be.ptr(v_colors),
                _be.ptr(v_alphas), _be.ptr(v_dexp), _be.ptr(v_dmed), _be.ptr(v_normals), _be.ptr(geom_grad),
                _be.ptr(color_grad), _be.ptr(abs_grad), flags, _be.ptr(n_isects), st), "rs_rasterize_bwd")
            _be.check(lib.rs_unpack_geom_grad(
                _be.ptr(geom_grad), _be.ptr(geom), _be.ptr(abs_grad), C, N, _be.ptr(opacities), int(opac_per_cam),
                _be.ptr(compensations), _be.ptr(v_means2d), _be.ptr(v_abs), _be.ptr(v_conics), _be.ptr(v_opac),
                _be.ptr(v_comps), _be.ptr(v_ray_ts), _be.ptr(v_ray_planes), _be.ptr(v_nrm),
                _be.ptr(v_col) if DP == 4 else None, int(color_per_cam), D, st), "rs_unpack_geom_grad")
            if DP > 4:
                if DP == D:
                    v_col = color_grad.view(colors_shape)
                else:
                    _be.check(lib.rs_unpack_colors_grad(_be.ptr(color_grad), rows, D, DP, _be.ptr(v_col), st),
                              "rs_unpack_colors_grad")
        if absgrad and ctx.means2d_ref is not None:
            # a render with more than 72 channels runs one compositing pass per channel chunk over the SAME means2d:
            # the chunks' |gradient| sums add up (rasterization() clears the attribute before the passes)
            prev = getattr(ctx.means2d_ref, "absgrad", None)
            ctx.means2d_ref.absgrad = v_abs if prev is None else prev + v_abs
        v_bg = None
        if backgrounds is not None and ctx.needs_input_grad[8]:
            v_bg = (v_colors * out_T[..., None]).sum(dim=(1, 2))
        return (v_means2d, v_conics, v_col, v_opac, v_comps, v_ray_ts, v_ray_planes, v_nrm, v_bg, None, None, None,
                None, None, None, None, None)


def rasterize_to_pixels(
    means2d: Tensor,            # [C,N,2]
    conics: Tensor,             # [C,N,3]
    colors: Tensor,             # [C,N,D] or [N,D] (shared by all cameras)
    opacities: Tensor,          # [C,N], or [N] (shared by all cameras)
    image_width: int,
    image_height: int,
    tile_size: int,
    isect_offsets: Tensor,      # [C,tile_h,tile_w] i32
    flatten_ids: Tensor,        # [M] i32
    backgrounds: Optional[Tensor] = None,   # [C,D]
    masks: Optional[Tensor] = None,
    packed: bool = False,
    absgrad: bool = False,
    ray_ts: Optional[Tensor] = None,        # [C,N]     RaDe: ray distance of the centre
    ray_planes: Optional[Tensor] = None,    # [C,N,2]   RaDe: -d(ray distance)/d(pixel)
    normals: Optional[Tensor] = None,       # [C,N,3]   RaDe: camera-space normals
    Ks: Optional[Tensor] = None,            # [C,3,3]   needed to turn ray distance into z depth
    return_ids: bool = False,
    compensations: Optional[Tensor] = None,  # [C,N]    fused: effective opacity = opacities * compensations
    ed_channel: int = -1,                    # fused "ED": that output channel is divided by max(alpha, 1e-10)
    n_isects: Optional[Tensor] = None,       # device int32 [1]: the number of valid entries of flatten_ids (sync-free
                                             # callers, whose buffers are capacity-sized); None = all of them
):
    """gsplat ``rasterize_to_pixels`` + the RaDe outputs.  Returns ``(colors [C,H,W,D], alphas [C,H,W,1])``
    or, when the RaDe inputs are given, ``(colors, alphas, expected_depths [C,H,W,1], median_depths
    [C,H,W,1], normals [C,H,W,3])``."""
    if packed or masks is not None:
        raise NotImplementedError("packed=True / tile masks are not on the collab-splats path")
    if tile_size != TILE_SIZE:
        raise NotImplementedError("tile_size must be 16")
    C, N = means2d.shape[:2]
    assert means2d.shape == (C, N, 2) and conics.shape == (C, N, 3), (means2d.shape, conics.shape)
    assert colors.shape[:-1] in ((C, N), (N,)), colors.shape
    assert opacities.shape in ((C, N), (N,)), opacities.shape
    assert compensations is None or compensations.shape == (C, N)
    rade = ray_ts is not None
    dev = means2d.device
    if rade:
        assert ray_planes is not None and normals is not None and Ks is not None
    else:
        ray_ts = torch.zeros(C, N, device=dev)
        ray_planes = torch.zeros(C, N, 2, device=dev)
        normals = torch.zeros(C, N, 3, device=dev)
        Ks = torch.eye(3, device=dev)[None].repeat(C, 1, 1)
    if backgrounds is not None:
        assert backgrounds.shape == (C, colors.shape[-1]), backgrounds.shape
    _need_cuda(means2d, conics, colors, opacities, isect_offsets, flatten_ids)
    out = _RasterizeToPixels.apply(_c(means2d) if not absgrad else means2d, _c(conics), _c(colors), _c(opacities),
                                   _c(compensations), _c(ray_ts), _c(ray_planes), _c(normals), _c(backgrounds),
                                   _c(Ks), int(image_width), int(image_height), _c(isect_offsets, torch.int32),
                                   _c(flatten_ids, torch.int32), bool(absgrad), int(ed_channel), n_isects)
    cols, alphas, dexp, dmed, nrm, last_ids, median_ids = out
    res = (cols, alphas, dexp, dmed, nrm) if rade else (cols, alphas)
    if return_ids:
        res = res + (last_ids, median_ids)
    return res
