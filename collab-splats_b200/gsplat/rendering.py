"""``gsplat.rendering.rasterization`` with the gsplat-rade call signature and outputs.

This is the function collab-splats calls every training step
(collab_splats/models/rade_gs_model.py:439-465, collab_splats/models/rade_features_model.py:450-476) and for
every view of the meshing sweep (collab_splats/utils/mesh.py:1582).  Stages:
projection -> [opacity compensation] -> colours (SH or N-D) -> [depth channel] -> tile intersection ->
radix sort -> offsets -> per-tile compositing -> [expected-depth normalisation of the ED channel].
All of them run as sm_100a kernels from librade_b200.so.
"""

from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from .cuda import _wrapper as _W
from .cuda._wrapper import (fully_fused_projection, isect_offset_encode, isect_tiles, isect_tiles_and_offsets,
                            rasterize_to_pixels, sh_colors, spherical_harmonics)

MAX_CHANNELS_PER_PASS = 72


def rasterization(
    means: Tensor,       # [N,3]
    quats: Tensor,       # [N,4]
    scales: Tensor,      # [N,3]
    opacities: Tensor,   # [N]
    colors: Tensor,      # [N,D] | [C,N,D] | SH: [N,K,3] | [C,N,K,3]
    viewmats: Tensor,    # [C,4,4]
    Ks: Tensor,          # [C,3,3]
    width: int,
    height: int,
    near_plane: float = 0.01,
    far_plane: float = 1e10,
    radius_clip: float = 0.0,
    eps2d: float = 0.3,
    sh_degree: Optional[int] = None,
    packed: bool = True,
    tile_size: int = 16,
    backgrounds: Optional[Tensor] = None,
    render_mode: str = "RGB",
    sparse_grad: bool = False,
    absgrad: bool = False,
    rasterize_mode: str = "classic",
    channel_chunk: int = 32,
    distributed: bool = False,
    camera_model: str = "pinhole",
    covars: Optional[Tensor] = None,
    return_depth_normal: bool = False,
    sync_free: Optional[bool] = None,
    **unsupported,
):
    """Returns ``(render_colors [C,H,W,D'], render_alphas [C,H,W,1], meta)`` or, with
    ``return_depth_normal=True`` (how the reference calls it), ``(render_colors, render_alphas,
    expected_depths [C,H,W,1], median_depths [C,H,W,1], expected_normals [C,H,W,3], meta)``.

    ``sync_free`` (not an upstream argument; None = ``gsplat.cuda._wrapper.SYNC_FREE``, default False): never read the
    number of intersections back to the host after the first render of a problem -- see ``_wrapper.SYNC_FREE``."""
    if unsupported:
        raise NotImplementedError(f"options not on the collab-splats path: {sorted(unsupported)}")
    if packed:
        raise NotImplementedError("packed=True is not on the collab-splats path (the reference passes packed=False, "
                                  "rade_gs_model.py:450)")
    if sparse_grad or distributed or covars is not None or camera_model != "pinhole":
        raise NotImplementedError("sparse_grad / distributed / covars / non-pinhole cameras are not on the "
                                  "collab-splats path")
    if tile_size != 16:
        raise NotImplementedError("tile_size must be 16")
    assert render_mode in ("RGB", "D", "ED", "RGB+D", "RGB+ED"), render_mode
    assert rasterize_mode in ("classic", "antialiased"), rasterize_mode
    N, C = means.shape[0], viewmats.shape[0]
    assert means.shape == (N, 3), means.shape
    assert quats.shape == (N, 4), quats.shape
    assert scales.shape == (N, 3), scales.shape
    assert opacities.shape == (N,), opacities.shape
    assert viewmats.shape == (C, 4, 4), viewmats.shape
    assert Ks.shape == (C, 3, 3), Ks.shape
    if sh_degree is None:
        assert (colors.dim() == 2 and colors.shape[0] == N) or (colors.dim() == 3 and colors.shape[:2] == (C, N)), \
            colors.shape
    else:
        assert (colors.dim() == 3 and colors.shape[0] == N and colors.shape[2] == 3) or \
               (colors.dim() == 4 and colors.shape[:2] == (C, N) and colors.shape[3] == 3), colors.shape
        assert (sh_degree + 1) ** 2 <= colors.shape[-2], colors.shape

    # ---- projection (+ RaDe ray-space terms)
    radii, means2d, depths, conics, compensations, ray_ts, ray_planes, normals = fully_fused_projection(
        means, None, quats, scales, viewmats, Ks, width, height, eps2d=eps2d, near_plane=near_plane,
        far_plane=far_plane, radius_clip=radius_clip, packed=False, sparse_grad=False,
        calc_compensations=(rasterize_mode == "antialiased"))
    # effective opacity = opacity * compensation is formed inside the pack kernel (and its VJP in the unpack
    # kernel); the [C,N] product below only feeds `meta` (detached, never differentiated)
    with torch.no_grad():
        opac_meta = opacities[None, :].expand(C, N) if compensations is None else opacities[None, :] * compensations

    # ---- colours
    with_depth = render_mode in ("RGB+D", "RGB+ED")
    slice_rgb = False
    if render_mode in ("D", "ED"):
        cols = depths[..., None]
        if backgrounds is not None:
            backgrounds = torch.zeros(C, 1, device=backgrounds.device)
    elif sh_degree is not None and colors.dim() == 3:
        # fused: campos -> dirs -> SH -> +0.5 -> clamp -> [rgb | depth], 4 channels (4th = 0 in plain RGB mode)
        cols = sh_colors(sh_degree, means, colors, viewmats, radii, depths if with_depth else None)
        slice_rgb = not with_depth
        if backgrounds is not None:
            backgrounds = torch.cat([backgrounds, torch.zeros(C, 1, device=backgrounds.device)], dim=-1)
    else:
        if sh_degree is None:
            cols = colors                          # [N,D] is shared by all cameras without expansion
        else:                                      # per-camera SH coefficients [C,N,K,3]
            campos = torch.linalg.inv_ex(viewmats, check_errors=False).inverse[:, :3, 3]
            dirs = means[None, :, :] - campos[:, None, :]
            cols = spherical_harmonics(sh_degree, dirs, colors, masks=(radii > 0).all(dim=-1))
            cols = torch.clamp_min(cols + 0.5, 0.0)
        if with_depth:
            if cols.dim() == 2:
                cols = cols[None].expand(C, N, cols.shape[-1])
            cols = torch.cat([cols, depths[..., None]], dim=-1)
            if backgrounds is not None:
                backgrounds = torch.cat([backgrounds, torch.zeros(C, 1, device=backgrounds.device)], dim=-1)

    # ---- tile intersection, sort, offsets
    tile_width = math.ceil(width / float(tile_size))
    tile_height = math.ceil(height / float(tile_size))
    n_isects = overflow = n_total = None
    sync_free = _W.SYNC_FREE if sync_free is None else bool(sync_free)
    sf = _W.isect_tiles_and_offsets_sync_free(means2d, radii, depths, tile_width, tile_height) if sync_free else None
    if sf is not None:      # no device->host read: capacity-sized lists, the count stays on the device
        tiles_per_gauss, isect_ids, flatten_ids, isect_offsets, n_isects, overflow, n_total = sf
    else:
        tiles_per_gauss, isect_ids, flatten_ids, isect_offsets = isect_tiles_and_offsets(
            means2d, radii, depths, tile_size, tile_width, tile_height)
        if sync_free:       # first call of this problem: remember how many intersections it has
            _W.isect_learn_capacity(means2d.device, C, N, tile_width, tile_height, flatten_ids.numel())

    # ---- compositing (one pass up to 72 channels; wider colours are split, geometry comes from the first pass).
    # The "ED" normalisation (depth channel / alpha) runs in the kernel epilogue.
    D = cols.shape[-1]
    ed = render_mode in ("ED", "RGB+ED")
    if D <= MAX_CHANNELS_PER_PASS:
        render_colors, render_alphas, exp_d, med_d, nrm = rasterize_to_pixels(
            means2d, conics, cols, opacities, width, height, tile_size, isect_offsets, flatten_ids,
            backgrounds=backgrounds, absgrad=absgrad, ray_ts=ray_ts, ray_planes=ray_planes, normals=normals, Ks=Ks,
            compensations=compensations, ed_channel=(D - 1) if ed else -1, n_isects=n_isects)
    else:
        chunk = min(max(int(channel_chunk), 1), 64) if channel_chunk > 32 else 64   # (upstream's default 32 -> one 64-wide pass)
        if absgrad and hasattr(means2d, "absgrad"):
            del means2d.absgrad        # the chunks' backward passes accumulate into it
        parts = []
        render_alphas = exp_d = med_d = nrm = None
        for k0 in range(0, D, chunk):
            k1 = min(k0 + chunk, D)
            bg = backgrounds[:, k0:k1] if backgrounds is not None else None
            out = rasterize_to_pixels(means2d, conics, cols[..., k0:k1], opacities, width, height, tile_size,
                                      isect_offsets, flatten_ids, backgrounds=bg, absgrad=absgrad, ray_ts=ray_ts,
                                      ray_planes=ray_planes, normals=normals, Ks=Ks, compensations=compensations,
                                      ed_channel=(k1 - k0 - 1) if (ed and k1 == D) else -1, n_isects=n_isects)
            parts.append(out[0])
            if k0 == 0:
                render_alphas, exp_d, med_d, nrm = out[1:]
        render_colors = torch.cat(parts, dim=-1)
    if slice_rgb:
        render_colors = render_colors[..., :3]

    meta: Dict = {
        "camera_ids": None, "gaussian_ids": None, "radii": radii, "means2d": means2d, "depths": depths,
        "conics": conics, "opacities": opac_meta, "ray_ts": ray_ts, "ray_planes": ray_planes, "normals": normals,
        "tile_width": tile_width, "tile_height": tile_height, "tiles_per_gauss": tiles_per_gauss,
        "isect_ids": isect_ids, "flatten_ids": flatten_ids, "isect_offsets": isect_offsets, "width": width,
        "height": height, "tile_size": tile_size, "n_cameras": C,
        "n_isects": n_total if n_total is not None else flatten_ids.numel(), "isect_overflow": overflow,
    }
    if return_depth_normal:
        return render_colors, render_alphas, exp_d, med_d, nrm, meta
    return render_colors, render_alphas, meta
