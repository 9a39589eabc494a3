"""CPU oracle for the RaDe-GS rasterizer hot path.  TEST INFRASTRUCTURE ONLY.

This file is a *restatement* (PyTorch on CPU, fp32 or fp64, autograd for the
backward pass) of the algorithm that collab-splats reaches through
``gsplat.rendering.rasterization`` (reference call sites:
``collab_splats/models/rade_gs_model.py:439-465`` and
``collab_splats/models/rade_features_model.py:450-476``; direct projection call at
``rade_gs_model.py:373-389``).  The arithmetic itself lives in the third-party
dependency ``gsplat-rade`` (``pyproject.toml:38``: ``gsplat @
git+https://github.com/brian-xu/gsplat-rade.git``, branch head, *no pinned
version*), whose source is not under ``/root/reference`` and which cannot be
installed in this container.  The reference's own tests never execute a render
(``tests/test_models.py:45-63`` only construct the models) and hold no golden
vectors for this path, therefore

    **PARITY UNPINNED**: this oracle restates the published gsplat 1.5.x
    ``_torch_impl`` semantics plus the RaDe-GS ray-space depth/normal terms
    (SURVEY.md Appendix A1-A10).  It is pinned only by analytic known-answer
    tests (tests/test_oracle_known_answers.py) and by finite differences.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module, and only as the checker or the
CPU baseline; the product path (``collab-splats_b200/``) never does.

Every convention that could not be verified against gsplat-rade source is a named
constant below (open questions Q1-Q7); the CUDA kernels use the
same values (``collab-splats_b200/csrc/rade_config.h``).

Arithmetic-order contract: the projection up to ``means2d``, ``depths``, the 2-D
covariance and ``radii`` is written as a fixed sequence of individually rounded
IEEE fp32 operations (no fused multiply-add, no library reductions).  The CUDA
kernel performs the same sequence with ``__fmul_rn/__fadd_rn/...`` so the integer
artefacts derived from those floats (radii, tile lists, sort keys, offsets) can be
compared bit-exactly end to end.
"""

from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch
from torch import Tensor

# --------------------------------------------------------------------------- constants
TILE_SIZE = 16
ALPHA_MIN = 1.0 / 255.0          # gsplat: skip pair if alpha < 1/255          
ALPHA_MAX = 0.999                # Q3: gsplat 0.999 (Inria/RaDe-GS 0.99)      
T_STOP = 1e-4                    # stop when T*(1-alpha) <= 1e-4              
RADIUS_SIGMA = 3.33              # radius = ceil(3.33*sqrt(cov_ii))           
DET_MIN = 1e-10                  # det = max(det, 1e-10)                      
FOV_PAD = 0.3                    # lim = ... + 0.3*tan_fov                    
VBN_EPS = 1e-7                   # Q5: vbn = max(n_hat . r, 1e-7)             
MEDIAN_INCLUSIVE = True          # Q2: median hit when T>0.5 and T' <= 0.5    
NORMALIZE_EXPECTED_DEPTH = False  # Q1: expected depth is raw sum(vis*t)/ln    
SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005,
         -1.0925484305920792, 0.5462742152960396)
SH_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658,
         0.3731763325901154, -0.4570457994644658, 1.445305721320277,
         -0.5900435899266435)


# --------------------------------------------------------------------------- projection
def _sqrt_rn(x: Tensor) -> Tensor:
    """Correctly rounded sqrt.  torch.sqrt on CPU goes through MKL VML (< 1 ulp, not 0.5 ulp:
    measured 0.6 % of fp32 results off by one ulp), which would break the arithmetic-order
    contract; sqrt in fp64 followed by one rounding to fp32 is correctly rounded."""
    if x.dtype == torch.float32:
        return torch.sqrt(x.double()).float()
    return torch.sqrt(x)


def quat_to_rotmat(quats: Tensor) -> Tuple[Tensor, ...]:
    """Un-normalised wxyz quaternion -> 9 rotation entries (same matrix as the
    in-tree ``build_rotation``, ``collab_splats/utils/camera_utils.py:138-168``)."""
    w, x, y, z = quats[..., 0], quats[..., 1], quats[..., 2], quats[..., 3]
    n2 = ((w * w + x * x) + y * y) + z * z
    inv = 1.0 / _sqrt_rn(n2)
    w, x, y, z = w * inv, x * inv, y * inv, z * inv
    x2, y2, z2 = x * x, y * y, z * z
    xy, xz, yz = x * y, x * z, y * z
    wx, wy, wz = w * x, w * y, w * z
    R00 = 1.0 - 2.0 * (y2 + z2)
    R01 = 2.0 * (xy - wz)
    R02 = 2.0 * (xz + wy)
    R10 = 2.0 * (xy + wz)
    R11 = 1.0 - 2.0 * (x2 + z2)
    R12 = 2.0 * (yz - wx)
    R20 = 2.0 * (xz - wy)
    R21 = 2.0 * (yz + wx)
    R22 = 1.0 - 2.0 * (x2 + y2)
    return R00, R01, R02, R10, R11, R12, R20, R21, R22


def fully_fused_projection(
    means: Tensor,      # [N,3]
    quats: Tensor,      # [N,4] wxyz, un-normalised
    scales: Tensor,     # [N,3] already exp()-activated (rade_gs_model.py:443)
    viewmats: Tensor,   # [C,4,4] world->camera
    Ks: Tensor,         # [C,3,3]
    width: int,
    height: int,
    eps2d: float = 0.3,
    near_plane: float = 0.01,
    far_plane: float = 1e10,
    radius_clip: float = 0.0,
    calc_compensations: bool = False,
    radii_override: Optional[Tensor] = None,
):
    """Restates gsplat ``fully_fused_projection`` (packed=False, pinhole) + RaDe terms.

    ``radii_override`` (test aid): take the integer radii -- and with them the culling decision -- from another run
    instead of deriving them here, so that an fp64 evaluation follows the discrete decisions of the fp32 one.

    Follows SURVEY.md rows a5 / Appendix A1-A5; call site rade_gs_model.py:373-389.
    Returns the 8-tuple the reference unpacks at rade_gs_model.py:392-394:
    radii [C,N,2] int32, means2d [C,N,2], depths [C,N], conics [C,N,3],
    compensations [C,N] | None, ray_ts [C,N], ray_planes [C,N,2], normals [C,N,3].
    Entries with radii == 0 are zero-filled.
    """
    dt = means.dtype
    C = viewmats.shape[0]
    N = means.shape[0]
    one = torch.ones((), dtype=dt)

    # ---- A1 covariance (per Gaussian, [N])
    R00, R01, R02, R10, R11, R12, R20, R21, R22 = quat_to_rotmat(quats)
    s0, s1, s2 = scales[:, 0], scales[:, 1], scales[:, 2]
    M00, M01, M02 = R00 * s0, R01 * s1, R02 * s2
    M10, M11, M12 = R10 * s0, R11 * s1, R12 * s2
    M20, M21, M22 = R20 * s0, R21 * s1, R22 * s2
    S00 = (M00 * M00 + M01 * M01) + M02 * M02
    S01 = (M00 * M10 + M01 * M11) + M02 * M12
    S02 = (M00 * M20 + M01 * M21) + M02 * M22
    S11 = (M10 * M10 + M11 * M11) + M12 * M12
    S12 = (M10 * M20 + M11 * M21) + M12 * M22
    S22 = (M20 * M20 + M21 * M21) + M22 * M22

    # ---- A2 world -> camera, broadcast to [C,N]
    def vm(i, j):
        return viewmats[:, i, j][:, None]

    W00, W01, W02, W10, W11, W12, W20, W21, W22 = (vm(0, 0), vm(0, 1), vm(0, 2), vm(1, 0),
                                                   vm(1, 1), vm(1, 2), vm(2, 0), vm(2, 1), vm(2, 2))
    t0, t1, t2 = vm(0, 3), vm(1, 3), vm(2, 3)
    mx, my, mz = means[:, 0][None], means[:, 1][None], means[:, 2][None]
    x = ((W00 * mx + W01 * my) + W02 * mz) + t0
    y = ((W10 * mx + W11 * my) + W12 * mz) + t1
    z = ((W20 * mx + W21 * my) + W22 * mz) + t2

    S00, S01, S02, S11, S12, S22 = (S00[None], S01[None], S02[None], S11[None], S12[None], S22[None])
    # A = W * Sigma
    A00 = (W00 * S00 + W01 * S01) + W02 * S02
    A01 = (W00 * S01 + W01 * S11) + W02 * S12
    A02 = (W00 * S02 + W01 * S12) + W02 * S22
    A10 = (W10 * S00 + W11 * S01) + W12 * S02
    A11 = (W10 * S01 + W11 * S11) + W12 * S12
    A12 = (W10 * S02 + W11 * S12) + W12 * S22
    A20 = (W20 * S00 + W21 * S01) + W22 * S02
    A21 = (W20 * S01 + W21 * S11) + W22 * S12
    A22 = (W20 * S02 + W21 * S12) + W22 * S22
    # Sigma_c = A * W^T (6 unique)
    V00 = (A00 * W00 + A01 * W01) + A02 * W02
    V01 = (A00 * W10 + A01 * W11) + A02 * W12
    V02 = (A00 * W20 + A01 * W21) + A02 * W22
    V11 = (A10 * W10 + A11 * W11) + A12 * W12
    V12 = (A10 * W20 + A11 * W21) + A12 * W22
    V22 = (A20 * W20 + A21 * W21) + A22 * W22

    # ---- A3 perspective EWA
    fx, fy = Ks[:, 0, 0][:, None], Ks[:, 1, 1][:, None]
    cx, cy = Ks[:, 0, 2][:, None], Ks[:, 1, 2][:, None]
    Wf = torch.full((), float(width), dtype=dt)
    Hf = torch.full((), float(height), dtype=dt)
    tanx = (0.5 * Wf) / fx
    tany = (0.5 * Hf) / fy
    lim_xp = (Wf - cx) / fx + FOV_PAD * tanx
    lim_xn = cx / fx + FOV_PAD * tanx
    lim_yp = (Hf - cy) / fy + FOV_PAD * tany
    lim_yn = cy / fy + FOV_PAD * tany
    rz = one / z
    u = torch.minimum(lim_xp, torch.maximum(-lim_xn, x * rz))
    v = torch.minimum(lim_yp, torch.maximum(-lim_yn, y * rz))
    tx = z * u
    ty = z * v
    rz2 = rz * rz
    J00 = fx * rz
    J02 = -((fx * tx) * rz2)
    J11 = fy * rz
    J12 = -((fy * ty) * rz2)
    B00 = J00 * V00 + J02 * V02
    B01 = J00 * V01 + J02 * V12
    B02 = J00 * V02 + J02 * V22
    B11 = J11 * V11 + J12 * V12
    B12 = J11 * V12 + J12 * V22
    c00 = B00 * J00 + B02 * J02
    c01 = B01 * J11 + B02 * J12
    c11 = B11 * J11 + B12 * J12
    m2x = (fx * x) * rz + cx
    m2y = (fy * y) * rz + cy

    # ---- A4 blur / conic / radius / cull
    det0 = c00 * c11 - c01 * c01
    c00b = c00 + eps2d
    c11b = c11 + eps2d
    det = torch.clamp(c00b * c11b - c01 * c01, min=DET_MIN)
    comp = _sqrt_rn(torch.clamp(det0 / det, min=0.0))
    conic_a = c11b / det
    conic_b = -c01 / det
    conic_c = c00b / det
    rx = torch.ceil(RADIUS_SIGMA * _sqrt_rn(c00b))
    ry = torch.ceil(RADIUS_SIGMA * _sqrt_rn(c11b))
    valid = (z > near_plane) & (z < far_plane)
    valid = valid & ~((rx <= radius_clip) & (ry <= radius_clip))
    valid = valid & ~((m2x + rx <= 0) | (m2x - rx >= width) | (m2y + ry <= 0) | (m2y - ry >= height))
    valid = valid & torch.isfinite(rx) & torch.isfinite(ry)
    if radii_override is not None:
        valid = (radii_override > 0).all(dim=-1)
        rx, ry = radii_override[..., 0].to(dt), radii_override[..., 1].to(dt)

    # ---- A5 RaDe ray-space plane + normal (clamped u,v)
    l2 = (u * u + v * v) + 1.0
    l = torch.sqrt((tx * tx + ty * ty) + z * z)
    # n = Sigma_c^{-1} r = W R diag(s^-2) R^T W^T r
    aw0 = W00 * u + W10 * v + W20
    aw1 = W01 * u + W11 * v + W21
    aw2 = W02 * u + W12 * v + W22
    R00, R01, R02, R10, R11, R12, R20, R21, R22 = (r[None] for r in (R00, R01, R02, R10, R11, R12, R20, R21, R22))
    bl0 = (R00 * aw0 + R10 * aw1 + R20 * aw2) / (s0 * s0)[None]
    bl1 = (R01 * aw0 + R11 * aw1 + R21 * aw2) / (s1 * s1)[None]
    bl2 = (R02 * aw0 + R12 * aw1 + R22 * aw2) / (s2 * s2)[None]
    cw0 = R00 * bl0 + R01 * bl1 + R02 * bl2
    cw1 = R10 * bl0 + R11 * bl1 + R12 * bl2
    cw2 = R20 * bl0 + R21 * bl1 + R22 * bl2
    n0 = W00 * cw0 + W01 * cw1 + W02 * cw2
    n1 = W10 * cw0 + W11 * cw1 + W12 * cw2
    n2 = W20 * cw0 + W21 * cw1 + W22 * cw2
    nn = torch.sqrt(n0 * n0 + n1 * n1 + n2 * n2)
    ok = torch.isfinite(nn) & (nn > 0)
    nn_safe = torch.where(ok, nn, torch.ones_like(nn))
    h0, h1, h2 = n0 / nn_safe, n1 / nn_safe, n2 / nn_safe
    vbn = torch.clamp(h0 * u + h1 * v + h2, min=VBN_EPS)
    w0, w1, w2 = h0 / vbn, h1 / vbn, h2 / vbn
    # plane = nJ_inv * w   (RaDe-GS computeCov2D): ((v^2+1) w0 - uv w1 - u w2, -uv w0 + (u^2+1) w1 - v w2)
    uv = u * v
    pl0 = (v * v + 1.0) * w0 - uv * w1 - u * w2
    pl1 = -uv * w0 + (u * u + 1.0) * w1 - v * w2
    fac = l / l2
    rp0 = pl0 * fac / fx
    rp1 = pl1 * fac / fy
    # camera-space normal = nJ * (-pl0*fac, -pl1*fac, -1)
    rn0, rn1 = -(pl0 * fac), -(pl1 * fac)
    cn0 = rn0 * rz - tx / l
    cn1 = rn1 * rz - ty / l
    cn2 = -(rn0 * tx + rn1 * ty) * rz2 - z / l
    cnn = torch.sqrt(cn0 * cn0 + cn1 * cn1 + cn2 * cn2)
    ok = ok & torch.isfinite(cnn) & (cnn > 0)
    cnn_safe = torch.where(ok, cnn, torch.ones_like(cnn))
    zero = torch.zeros_like(l)
    ray_t = torch.where(ok, l, zero)
    rp0 = torch.where(ok, rp0, zero)
    rp1 = torch.where(ok, rp1, zero)
    nx = torch.where(ok, cn0 / cnn_safe, zero)
    ny = torch.where(ok, cn1 / cnn_safe, zero)
    nz = torch.where(ok, cn2 / cnn_safe, zero)

    def vz(t):
        return torch.where(valid, t, torch.zeros_like(t))

    radii = torch.stack([vz(rx), vz(ry)], dim=-1).to(torch.int32)
    means2d = torch.stack([vz(m2x), vz(m2y)], dim=-1)
    depths = vz(z)
    conics = torch.stack([vz(conic_a), vz(conic_b), vz(conic_c)], dim=-1)
    comps = vz(comp) if calc_compensations else None
    ray_ts = vz(ray_t)
    ray_planes = torch.stack([vz(rp0), vz(rp1)], dim=-1)
    normals = torch.stack([vz(nx), vz(ny), vz(nz)], dim=-1)
    return radii, means2d, depths, conics, comps, ray_ts, ray_planes, normals


# --------------------------------------------------------------------------- spherical harmonics
def spherical_harmonics(degrees_to_use: int, dirs: Tensor, coeffs: Tensor, masks: Optional[Tensor] = None) -> Tensor:
    """Real SH of normalised ``dirs`` up to ``degrees_to_use`` (<=3), gsplat ordering/signs
    (reference call rade_features_model.py:430-434).
    dirs [...,3], coeffs [...,K,3] -> [...,3].  Masked-out entries return 0."""
    assert 0 <= degrees_to_use <= 3
    assert (degrees_to_use + 1) ** 2 <= coeffs.shape[-2]
    nrm = torch.sqrt(dirs[..., 0] ** 2 + dirs[..., 1] ** 2 + dirs[..., 2] ** 2)
    nrm = torch.clamp(nrm, min=1e-12)  # F.normalize eps
    x, y, z = dirs[..., 0] / nrm, dirs[..., 1] / nrm, dirs[..., 2] / nrm
    out = SH_C0 * coeffs[..., 0, :]
    if degrees_to_use >= 1:
        out = out + SH_C1 * (-y[..., None] * coeffs[..., 1, :] + z[..., None] * coeffs[..., 2, :]
                             - x[..., None] * coeffs[..., 3, :])
    if degrees_to_use >= 2:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        out = (out + (SH_C2[0] * xy)[..., None] * coeffs[..., 4, :]
               + (SH_C2[1] * yz)[..., None] * coeffs[..., 5, :]
               + (SH_C2[2] * (2.0 * zz - xx - yy))[..., None] * coeffs[..., 6, :]
               + (SH_C2[3] * xz)[..., None] * coeffs[..., 7, :]
               + (SH_C2[4] * (xx - yy))[..., None] * coeffs[..., 8, :])
    if degrees_to_use >= 3:
        out = (out + (SH_C3[0] * y * (3.0 * xx - yy))[..., None] * coeffs[..., 9, :]
               + (SH_C3[1] * xy * z)[..., None] * coeffs[..., 10, :]
               + (SH_C3[2] * y * (4.0 * zz - xx - yy))[..., None] * coeffs[..., 11, :]
               + (SH_C3[3] * z * (2.0 * zz - 3.0 * xx - 3.0 * yy))[..., None] * coeffs[..., 12, :]
               + (SH_C3[4] * x * (4.0 * zz - xx - yy))[..., None] * coeffs[..., 13, :]
               + (SH_C3[5] * z * (xx - yy))[..., None] * coeffs[..., 14, :]
               + (SH_C3[6] * x * (xx - 3.0 * yy))[..., None] * coeffs[..., 15, :])
    if masks is not None:
        out = torch.where(masks[..., None], out, torch.zeros_like(out))
    return out


# --------------------------------------------------------------------------- tile intersection
def tile_bits_for(n_tiles: int) -> int:
    return int(n_tiles).bit_length()


def isect_tiles(means2d: Tensor, radii: Tensor, depths: Tensor, tile_size: int, tile_width: int,
                tile_height: int, sort: bool = True):
    """Restates gsplat ``isect_tiles``.  Integer outputs, exact.

    means2d [C,N,2] fp32, radii [C,N,2] int32, depths [C,N] fp32 ->
    tiles_per_gauss [C,N] int32, isect_ids [M] int64, flatten_ids [M] int32."""
    C, N = depths.shape
    m = means2d.detach().to(torch.float32).numpy()
    r = radii.detach().numpy().astype(np.float32)
    d = depths.detach().to(torch.float32).numpy()
    ts = np.float32(tile_size)
    tile_xy = m / ts
    tile_r = r / ts
    tmin = np.floor(tile_xy - tile_r).astype(np.int64)
    tmax = np.ceil(tile_xy + tile_r).astype(np.int64)
    tmin[..., 0] = np.clip(tmin[..., 0], 0, tile_width)
    tmin[..., 1] = np.clip(tmin[..., 1], 0, tile_height)
    tmax[..., 0] = np.clip(tmax[..., 0], 0, tile_width)
    tmax[..., 1] = np.clip(tmax[..., 1], 0, tile_height)
    visible = (radii.detach().numpy() > 0).all(-1)
    tiles = (tmax[..., 0] - tmin[..., 0]) * (tmax[..., 1] - tmin[..., 1])
    tiles = np.where(visible, tiles, 0).astype(np.int32)
    n_tiles = tile_width * tile_height
    tile_bits = tile_bits_for(n_tiles)
    cum = np.cumsum(tiles.reshape(-1).astype(np.int64))
    M = int(cum[-1]) if cum.size else 0
    isect_ids = np.zeros(M, dtype=np.int64)
    flatten_ids = np.zeros(M, dtype=np.int32)
    depth_bits = d.view(np.int32).astype(np.int64) & 0xFFFFFFFF
    flat_tiles = tiles.reshape(-1)
    nz = np.nonzero(flat_tiles)[0]
    # vectorised emission in (c, n, y, x) order
    if M > 0:
        starts = cum[nz] - flat_tiles[nz]
        reps = flat_tiles[nz].astype(np.int64)
        owner = np.repeat(nz, reps)                        # flatten id per isect
        local = np.arange(M, dtype=np.int64) - np.repeat(starts, reps)
        cc, nn = owner // N, owner % N
        wx = (tmax[cc, nn, 0] - tmin[cc, nn, 0])
        ty = tmin[cc, nn, 1] + local // wx
        tx = tmin[cc, nn, 0] + local % wx
        tile_id = ty * tile_width + tx
        isect_ids = (cc << (32 + tile_bits)) | (tile_id << 32) | depth_bits[cc, nn]
        flatten_ids = owner.astype(np.int32)
    if sort and M > 0:
        order = np.argsort(isect_ids, kind="stable")
        isect_ids = isect_ids[order]
        flatten_ids = flatten_ids[order]
    return (torch.from_numpy(tiles.reshape(C, N)), torch.from_numpy(isect_ids),
            torch.from_numpy(flatten_ids))


def isect_offset_encode(isect_ids: Tensor, n_cameras: int, tile_width: int, tile_height: int) -> Tensor:
    """Restates gsplat ``isect_offset_encode``: offsets[c,ty,tx] = lower bound of
    (c,tile) in the sorted key list.  -> [C,TH,TW] int32."""
    n_tiles = tile_width * tile_height
    tile_bits = tile_bits_for(n_tiles)
    ids = isect_ids.numpy()
    hi = ids >> 32
    cam = np.repeat(np.arange(n_cameras, dtype=np.int64), n_tiles)
    tile = np.tile(np.arange(n_tiles, dtype=np.int64), n_cameras)
    q = (cam << tile_bits) | tile
    off = np.searchsorted(hi, q, side="left").astype(np.int32)
    return torch.from_numpy(off.reshape(n_cameras, tile_height, tile_width))


# --------------------------------------------------------------------------- compositing
def rasterize_to_pixels(
    means2d: Tensor,      # [C,N,2]
    conics: Tensor,       # [C,N,3]
    colors: Tensor,       # [C,N,D]
    opacities: Tensor,    # [C,N]
    ray_ts: Tensor,       # [C,N]
    ray_planes: Tensor,   # [C,N,2]
    normals: Tensor,      # [C,N,3]
    Ks: Tensor,           # [C,3,3]
    width: int,
    height: int,
    tile_size: int,
    isect_offsets: Tensor,  # [C,TH,TW] int32
    flatten_ids: Tensor,    # [M] int32
    backgrounds: Optional[Tensor] = None,  # [C,D]
    return_aux: bool = False,
    tile_window: Optional[Tuple[int, int, int, int]] = None,  # (ty0, ty1, tx0, tx1): composite only these tiles
):
    """Front-to-back alpha compositing per tile, autograd-differentiable.

    ``tile_window`` (test aid for full-size scenes): tiles outside it are left at zero (no background), so a
    1080p view of a million Gaussians can be checked on a window the CPU finishes in seconds.

    Returns colors [C,H,W,D], alphas [C,H,W,1], expected_depths [C,H,W,1] (z-depth),
    median_depths [C,H,W,1] (z-depth, 0 where never crossed), normals [C,H,W,3]
    (+ aux dict: last_ids, median_ids [C,H,W] int32, fragile [C,H,W] bool, n_tested, n_contrib).

    ``fragile`` marks pixels where a discrete decision (alpha threshold, T stop, median
    crossing) sits within a relative 2e-5 of its threshold, i.e. where a 1-ulp difference in
    exp() legitimately changes the result; parity tests exclude exactly those pixels.
    """
    assert tile_size == TILE_SIZE
    C, N = opacities.shape
    D = colors.shape[-1]
    dt = means2d.dtype
    TH, TW = isect_offsets.shape[1:]
    M = flatten_ids.shape[0]
    offs = isect_offsets.reshape(-1).tolist() + [M]
    out_c = torch.zeros(C, height, width, D, dtype=dt)
    out_a = torch.zeros(C, height, width, 1, dtype=dt)
    out_de = torch.zeros(C, height, width, 1, dtype=dt)
    out_dm = torch.zeros(C, height, width, 1, dtype=dt)
    out_n = torch.zeros(C, height, width, 3, dtype=dt)
    last_ids = torch.zeros(C, height, width, dtype=torch.int32)
    med_ids = torch.full((C, height, width), -1, dtype=torch.int32)
    fragile = torch.zeros(C, height, width, dtype=torch.bool)
    n_tested = 0
    n_contrib = 0
    pieces = []  # (c, y0, y1, x0, x1, tensors...) assembled with differentiable cat at the end

    flat_m2 = means2d.reshape(C * N, 2)
    flat_con = conics.reshape(C * N, 3)
    flat_col = colors.reshape(C * N, D)
    flat_op = opacities.reshape(C * N)
    flat_rt = ray_ts.reshape(C * N)
    flat_rp = ray_planes.reshape(C * N, 2)
    flat_nr = normals.reshape(C * N, 3)

    rows_c, rows_a, rows_de, rows_dm, rows_n = [], [], [], [], []
    for c in range(C):
        fx, fy, cx, cy = Ks[c, 0, 0], Ks[c, 1, 1], Ks[c, 0, 2], Ks[c, 1, 2]
        tile_rows = []
        for ty in range(TH):
            y0, y1 = ty * tile_size, min((ty + 1) * tile_size, height)
            row_tiles = []
            for tx_ in range(TW):
                x0, x1 = tx_ * tile_size, min((tx_ + 1) * tile_size, width)
                tid = (c * TH + ty) * TW + tx_
                s, e = offs[tid], offs[tid + 1]
                ph, pw = y1 - y0, x1 - x0
                P = ph * pw
                py = (torch.arange(y0, y1, dtype=dt) + 0.5)[:, None].expand(ph, pw).reshape(P)
                px = (torch.arange(x0, x1, dtype=dt) + 0.5)[None, :].expand(ph, pw).reshape(P)
                ln = torch.sqrt(((px - cx) / fx) ** 2 + ((py - cy) / fy) ** 2 + 1.0)
                bg = backgrounds[c] if backgrounds is not None else None
                outside = tile_window is not None and not (tile_window[0] <= ty < tile_window[1] and
                                                           tile_window[2] <= tx_ < tile_window[3])
                if e <= s or outside:
                    col = torch.zeros(P, D, dtype=dt)
                    if bg is not None and not outside:
                        col = col + bg[None, :]
                    row_tiles.append((col.reshape(ph, pw, D), torch.zeros(ph, pw, 1, dtype=dt),
                                      torch.zeros(ph, pw, 1, dtype=dt), torch.zeros(ph, pw, 1, dtype=dt),
                                      torch.zeros(ph, pw, 3, dtype=dt)))
                    last_ids[c, y0:y1, x0:x1] = s - 1
                    continue
                ids = flatten_ids[s:e].long()
                G = ids.shape[0]
                xy = flat_m2[ids]          # [G,2]
                con = flat_con[ids]
                op = flat_op[ids]
                dx = xy[None, :, 0] - px[:, None]   # [P,G]
                dy = xy[None, :, 1] - py[:, None]
                sigma = 0.5 * (con[None, :, 0] * dx * dx + con[None, :, 2] * dy * dy) + con[None, :, 1] * dx * dy
                a_raw = op[None, :] * torch.exp(-sigma)
                alpha = torch.clamp(a_raw, max=ALPHA_MAX)
                ok = (sigma >= 0) & (alpha >= ALPHA_MIN)
                alpha = torch.where(ok, alpha, torch.zeros_like(alpha))
                one_m = 1.0 - alpha
                T_incl = torch.cumprod(one_m, dim=1)                 # T after Gaussian g
                T_excl = torch.cat([torch.ones(P, 1, dtype=dt), T_incl[:, :-1]], dim=1)
                live = T_incl > T_STOP                               # prefix mask (monotone)
                # a pixel stops at the first g with T_incl <= T_STOP, which must be a contributing one
                live = torch.cummin(live.to(torch.int8), dim=1).values.bool()
                contrib = ok & live
                vis = torch.where(contrib, alpha * T_excl, torch.zeros_like(alpha))
                n_live = live.sum(dim=1)                             # Gaussians visited w/o stopping
                T_fin = torch.where(n_live > 0,
                                    T_incl.gather(1, (n_live - 1).clamp(min=0)[:, None])[:, 0],
                                    torch.ones(P, dtype=dt))
                t = flat_rt[ids][None, :] + flat_rp[ids][None, :, 0] * dx + flat_rp[ids][None, :, 1] * dy
                col = vis @ flat_col[ids]                            # [P,D]
                if bg is not None:
                    col = col + T_fin[:, None] * bg[None, :]
                dsum = (vis * t).sum(dim=1)
                nrm = vis @ flat_nr[ids]
                if MEDIAN_INCLUSIVE:
                    cross = contrib & (T_excl > 0.5) & (T_incl <= 0.5)
                else:
                    cross = contrib & (T_excl > 0.5) & (T_incl < 0.5)
                has_med = cross.any(dim=1)
                mg = cross.to(torch.int8).argmax(dim=1)
                t_med = torch.where(has_med, t.gather(1, mg[:, None])[:, 0], torch.zeros(P, dtype=dt))
                a_out = 1.0 - T_fin
                d_exp = dsum / ln
                if NORMALIZE_EXPECTED_DEPTH:
                    d_exp = d_exp / torch.clamp(a_out, min=1e-10)
                d_med = t_med / ln
                row_tiles.append((col.reshape(ph, pw, D), a_out.reshape(ph, pw, 1), d_exp.reshape(ph, pw, 1),
                                  d_med.reshape(ph, pw, 1), nrm.reshape(ph, pw, 3)))
                with torch.no_grad():
                    idxs = torch.arange(G)[None, :].expand(P, G)
                    last = torch.where(contrib, idxs, torch.full_like(idxs, -1)).max(dim=1).values
                    last_ids[c, y0:y1, x0:x1] = (last + s).to(torch.int32).reshape(ph, pw)  # s-1 if none
                    med_ids[c, y0:y1, x0:x1] = torch.where(has_med, mg + s, torch.full_like(mg, -1)
                                                           ).to(torch.int32).reshape(ph, pw)
                    # fragility: decisions within rel 2e-5 of a threshold, among Gaussians actually visited
                    visited = torch.cat([torch.ones(P, 1, dtype=torch.bool), live[:, :-1]], dim=1)
                    a_chk = torch.clamp(a_raw, max=ALPHA_MAX)
                    near_a = ((a_chk - ALPHA_MIN).abs() < 2e-5 * ALPHA_MIN * 10) & (sigma >= 0)
                    near_t = (T_incl - T_STOP).abs() < 2e-5 * T_STOP * 10
                    near_m = (T_incl - 0.5).abs() < 1e-5
                    fr = (visited & (near_a | near_t | near_m)).any(dim=1)
                    fragile[c, y0:y1, x0:x1] = fr.reshape(ph, pw)
                    n_tested += int(visited.sum())
                    n_contrib += int(contrib.sum())
            tile_rows.append(tuple(torch.cat([rt[k] for rt in row_tiles], dim=1) for k in range(5)))
        rows_c.append(torch.cat([tr[0] for tr in tile_rows], dim=0))
        rows_a.append(torch.cat([tr[1] for tr in tile_rows], dim=0))
        rows_de.append(torch.cat([tr[2] for tr in tile_rows], dim=0))
        rows_dm.append(torch.cat([tr[3] for tr in tile_rows], dim=0))
        rows_n.append(torch.cat([tr[4] for tr in tile_rows], dim=0))
    out = (torch.stack(rows_c), torch.stack(rows_a), torch.stack(rows_de), torch.stack(rows_dm),
           torch.stack(rows_n))
    if return_aux:
        aux = dict(last_ids=last_ids, median_ids=med_ids, fragile=fragile, n_tested=n_tested,
                   n_contrib=n_contrib)
        return out + (aux,)
    return out


# --------------------------------------------------------------------------- orchestration
def rasterization(
    means: Tensor, quats: Tensor, scales: Tensor, opacities: Tensor, colors: Tensor,
    viewmats: Tensor, Ks: Tensor, width: int, height: int,
    near_plane: float = 0.01, far_plane: float = 1e10, radius_clip: float = 0.0, eps2d: float = 0.3,
    sh_degree: Optional[int] = None, tile_size: int = 16, backgrounds: Optional[Tensor] = None,
    render_mode: str = "RGB", rasterize_mode: str = "classic", return_depth_normal: bool = False,
    return_aux: bool = False, tile_window: Optional[Tuple[int, int, int, int]] = None,
    compositor: str = "torch", threads: Optional[int] = None, discrete_from: Optional[Dict] = None,
    fragile_scale: float = 1.0,
):
    """Restates ``gsplat.rendering.rasterization`` for the options the reference uses
    (call site rade_gs_model.py:439-465): packed=False, pinhole, 3DGS.

    ``discrete_from`` (test aid): the ``meta`` of another run of the same scene; its integer artefacts (radii, tile
    lists, sort order, offsets) are used instead of recomputing them.  The render is a discontinuous function of the
    inputs through those integers (a radius that rounds the other way adds or removes tiles), so an fp64 evaluation
    that is to serve as the ground truth of an fp32 one has to follow the fp32 run's discrete decisions.

    ``compositor="c"`` runs the per-tile compositing stage (forward and backward) through the C restatement
    (oracle/raster_oracle.c, ``threads`` host threads) instead of the PyTorch one; everything else is unchanged."""
    assert render_mode in ("RGB", "D", "ED", "RGB+D", "RGB+ED")
    assert rasterize_mode in ("classic", "antialiased")
    C, N = viewmats.shape[0], means.shape[0]
    radii, means2d, depths, conics, comps, ray_ts, ray_planes, normals = fully_fused_projection(
        means, quats, scales, viewmats, Ks, width, height, eps2d=eps2d, near_plane=near_plane,
        far_plane=far_plane, radius_clip=radius_clip, calc_compensations=(rasterize_mode == "antialiased"),
        radii_override=None if discrete_from is None else discrete_from["radii"])
    opac = opacities[None, :].expand(C, N)
    if comps is not None:
        opac = opac * comps
    if sh_degree is None:
        cols = colors[None].expand(C, N, colors.shape[-1]) if colors.dim() == 2 else colors
    else:
        campos = torch.linalg.inv(viewmats)[:, :3, 3]                 # [C,3]
        dirs = means[None, :, :] - campos[:, None, :]
        sh = colors[None].expand(C, *colors.shape) if colors.dim() == 3 else colors
        cols = spherical_harmonics(sh_degree, dirs, sh, masks=(radii > 0).all(-1))
        cols = torch.clamp_min(cols + 0.5, 0.0)
    if render_mode in ("RGB+D", "RGB+ED"):
        cols = torch.cat([cols, depths[..., None]], dim=-1)
        if backgrounds is not None:
            backgrounds = torch.cat([backgrounds, torch.zeros(C, 1, dtype=backgrounds.dtype)], dim=-1)
    elif render_mode in ("D", "ED"):
        cols = depths[..., None]
        if backgrounds is not None:
            backgrounds = torch.zeros(C, 1, dtype=backgrounds.dtype)
    TW = math.ceil(width / tile_size)
    TH = math.ceil(height / tile_size)
    if discrete_from is None:
        tiles_per_gauss, isect_ids, flatten_ids = isect_tiles(means2d, radii, depths, tile_size, TW, TH)
        isect_offsets = isect_offset_encode(isect_ids, C, TW, TH)
    else:
        tiles_per_gauss, isect_ids, flatten_ids, isect_offsets = (
            discrete_from[k] for k in ("tiles_per_gauss", "isect_ids", "flatten_ids", "isect_offsets"))
    if compositor == "c":
        from . import raster_oracle as _RO
        res = _RO.rasterize_to_pixels(means2d, conics, cols, opac, ray_ts, ray_planes, normals, Ks, width, height,
                                      tile_size, isect_offsets, flatten_ids, backgrounds=backgrounds,
                                      return_aux=True, threads=threads, fragile_scale=fragile_scale)
    else:
        assert compositor == "torch", compositor
        res = rasterize_to_pixels(means2d, conics, cols, opac, ray_ts, ray_planes, normals, Ks, width, height,
                                  tile_size, isect_offsets, flatten_ids, backgrounds=backgrounds, return_aux=True,
                                  tile_window=tile_window)
    render_colors, render_alphas, exp_d, med_d, nrm, aux = res
    if render_mode in ("ED", "RGB+ED"):
        render_colors = torch.cat([render_colors[..., :-1],
                                   render_colors[..., -1:] / render_alphas.clamp(min=1e-10)], dim=-1)
    meta = dict(camera_ids=None, gaussian_ids=None, radii=radii, means2d=means2d, depths=depths, conics=conics,
                opacities=opac, ray_ts=ray_ts, ray_planes=ray_planes, normals=normals, colors=cols,
                tile_width=TW, tile_height=TH, tiles_per_gauss=tiles_per_gauss, isect_ids=isect_ids,
                flatten_ids=flatten_ids, isect_offsets=isect_offsets, width=width, height=height,
                tile_size=tile_size, n_cameras=C)
    if return_aux:
        meta.update(aux)
    if return_depth_normal:
        return render_colors, render_alphas, exp_d, med_d, nrm, meta
    return render_colors, render_alphas, meta


# --------------------------------------------------------------------------- model-side loss (in-tree, exact)
def depth_double_to_normal(Ks_c: Tensor, width: int, height: int, depth1: Tensor, depth2: Tensor) -> Tensor:
    """Restates ``collab_splats/utils/camera_utils.py:176-279`` for one pinhole camera with a
    centred principal point: two z-depth maps [H,W] -> normals [2,H,W,3] (border = 0)."""
    dt = depth1.dtype
    fx, fy = Ks_c[0, 0], Ks_c[1, 1]
    dev = depth1.device  # device-agnostic so the tests can apply the same loss to the CUDA outputs
    gx = (torch.arange(width, dtype=dt, device=dev) + 0.5)[None, :].expand(height, width)
    gy = (torch.arange(height, dtype=dt, device=dev) + 0.5)[:, None].expand(height, width)
    rx = gx / fx - width / (2 * fx)
    ry = gy / fy - height / (2 * fy)
    rays = torch.stack([rx, ry, torch.ones_like(rx)], dim=0)                # [3,H,W]
    pts = torch.stack([depth1[None] * rays, depth2[None] * rays], dim=0)    # [2,3,H,W]
    out = torch.zeros_like(pts)
    d_row = pts[..., 2:, 1:-1] - pts[..., :-2, 1:-1]
    d_col = pts[..., 1:-1, 2:] - pts[..., 1:-1, :-2]
    nm = torch.nn.functional.normalize(torch.cross(d_row, d_col, dim=1), dim=1)
    out[..., 1:-1, 1:-1] = nm
    return out.permute(0, 2, 3, 1)


def depth_normal_loss(Ks_c: Tensor, width: int, height: int, exp_depth: Tensor, med_depth: Tensor,
                      rendered_normals: Tensor, lam: float = 0.05, depth_ratio: float = 0.6):
    """Restates rade_gs_model.py:202-219 + :292-307 (lambda 0.05, ratio 0.6: rade_gs_method.py:38-40).
    exp_depth/med_depth [H,W], rendered_normals [H,W,3] -> (loss, error maps [2,H,W])."""
    n_d = depth_double_to_normal(Ks_c, width, height, exp_depth, med_depth)      # [2,H,W,3]
    err = 1.0 - (rendered_normals[None] * n_d).sum(dim=-1)                       # [2,H,W]
    loss = lam * ((1.0 - depth_ratio) * err[0].mean() + depth_ratio * err[1].mean())
    return loss, err


def get_outputs_glue(Ks_c: Tensor, width: int, height: int, render: Tensor, alpha: Tensor, expected_depths: Tensor,
                     median_depths: Tensor, expected_normals: Tensor, background: Tensor, render_mode: str = "RGB+ED",
                     use_depth_normal: bool = True) -> Dict[str, Optional[Tensor]]:
    """Restates the part of ``RadegsModel.get_outputs`` that follows the rasterization call
    (collab_splats/models/rade_gs_model.py:200-271) for one camera: render [H,W,D], alpha /
    depths [H,W,1], expected_normals [H,W,3], background [3] -> the reference's output dict.  Pinned to the
    reference's own lines by tests/golden/rade_outputs.npz (tests/golden/make_outputs_golden.py)."""
    if use_depth_normal:
        n_d = depth_double_to_normal(Ks_c, width, height, expected_depths[..., 0], median_depths[..., 0])  # [2,H,W,3]
        err = 1.0 - (expected_normals[None] * n_d).sum(dim=-1)                                            # [2,H,W]
    else:
        err = torch.zeros(2, height, width, dtype=render.dtype, device=render.device)
    normals = (expected_normals + 1) / 2
    rgb = torch.clamp(render[..., :3] + (1 - alpha) * background, 0.0, 1.0)
    hit = alpha > 0
    depth_im = None
    if render_mode == "RGB+ED":
        d = render[..., 3:4]
        depth_im = torch.where(hit, d, d.detach().max())
    return {"rgb": rgb, "depth": torch.where(hit, expected_depths, expected_depths.detach().max()),
            "median_depth": torch.where(hit, median_depths, median_depths.detach().max()), "depth_im": depth_im,
            "accumulation": alpha, "normals": torch.where(hit, normals, normals.detach().max()),
            "depth_normal_error_map": err[0][..., None], "middepth_normal_error_map": err[1][..., None],
            "background": background}
