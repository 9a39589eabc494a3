"""CPU oracle for the TSDF fusion of rendered frames.  TEST INFRASTRUCTURE ONLY.

Restates, in numpy, what the reference's meshing exporter does per frame at
``collab_splats/utils/mesh.py:1562-1632``: ``o3d.pipelines.integration.ScalableTSDFVolume(voxel_length,
sdf_trunc, color_type=RGB8)`` followed by ``volume.integrate(rgbd, intrinsic, extrinsic)`` for every training
camera, with ``rgbd = RGBDImage.create_from_color_and_depth(color_u8, depth_f32, depth_scale=1.0,
depth_trunc=..., convert_rgb_to_intensity=False)``.

The arithmetic lives in the third-party dependency Open3D (``pyproject.toml``: ``open3d``, unpinned), which is not
under ``/root/reference`` and is not installed here, and the reference has no test or fixture for this path,
therefore **PARITY UNPINNED**: this file restates Open3D's published algorithm (0.17-0.19 line,
``ScalableTSDFVolume::Integrate`` + ``UniformTSDFVolume::IntegrateWithDepthToCameraDistanceMultiplier``):

  1. depth values >= depth_trunc are dropped (``RGBDImage::CreateFromColorAndDepth`` ->
     ``ConvertDepthToFloatImage``); 0 means "no measurement";
  2. every ``depth_sampling_stride``-th pixel (default 4) with depth > 0 is back-projected with
     ``x = (j - cx) * z / fx``, ``y = (i - cy) * z / fy`` and moved to the world with the inverse extrinsic;
     every volume unit (16^3 voxels, ``unit_length = 16 * voxel_length``) whose index lies in
     ``floor((p -+ sdf_trunc) / unit_length)`` is opened and marked as touched by this frame;
  3. every voxel of a touched unit: centre ``(index + 0.5) * voxel_length`` -> camera; if ``z > 0`` and the pixel
     ``(int(x*fx/z + cx + 0.5), int(y*fy/z + cy + 0.5))`` lies inside ``[0.0001, W - 0.0001)`` and holds a depth:
     ``sdf = (d - z) * sqrt(xn^2 + yn^2 + 1)``; if ``sdf > -sdf_trunc``: ``tsdf' = (tsdf*w + min(1, sdf/sdf_trunc))
     / (w + 1)``, ``color' = (color*w + rgb) / (w + 1)``, ``w += 1``.

Open3D back-projects in fp64 and steps the camera-space point along z incrementally; here (and in
``collab-splats_b200/csrc/tsdf.cu``) every quantity is a fixed sequence of individually rounded fp32 operations, so
the oracle and the kernel agree bit for bit; against Open3D itself the voxel values would agree to fp32 round-off.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this module.
"""

from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np

UNIT_RES = 16                      # Open3D ScalableTSDFVolume default volume_unit_resolution
DEFAULT_STRIDE = 4                 # Open3D ScalableTSDFVolume default depth_sampling_stride
F = np.float32


def _sqrt32(x: np.ndarray) -> np.ndarray:
    return np.sqrt(x.astype(np.float64)).astype(F)      # correctly rounded fp32 sqrt


def frame_matrices(extrinsic_4x4) -> Tuple[np.ndarray, np.ndarray]:
    """world->camera 4x4 (any float type) -> (E, P): fp32 [3,4] extrinsic and its fp64-inverted, fp32-rounded pose.
    The host mirror (radegs_b200/tsdf.py) builds the same two matrices."""
    e64 = np.asarray(extrinsic_4x4, dtype=np.float64).reshape(4, 4)
    return e64[:3].astype(F), np.linalg.inv(e64)[:3].astype(F)


class TsdfOracleVolume:
    """dict: unit (ux,uy,uz) -> {"tsdf": f32[16,16,16], "weight": f32[16,16,16], "rgb": f32[16,16,16,3]}."""

    def __init__(self, voxel_length: float, sdf_trunc: float, with_color: bool = True,
                 depth_sampling_stride: int = DEFAULT_STRIDE):
        self.voxel_length = F(voxel_length)
        self.sdf_trunc = F(sdf_trunc)
        self.with_color = with_color
        self.stride = int(depth_sampling_stride)
        self.units: Dict[Tuple[int, int, int], Dict[str, np.ndarray]] = {}
        self.last_touched = []

    # -- step 2
    def touched_units(self, depth: np.ndarray, fx, fy, cx, cy, P: np.ndarray, depth_trunc: float):
        H, W = depth.shape
        fx, fy, cx, cy = F(fx), F(fy), F(cx), F(cy)
        ii, jj = np.meshgrid(np.arange(0, H, self.stride), np.arange(0, W, self.stride), indexing="ij")
        d = depth[ii, jj].astype(F)
        ok = (d > 0) & ~(d >= F(depth_trunc))
        d, ii, jj = d[ok], ii[ok].astype(F), jj[ok].astype(F)
        x = ((jj - cx) * d) / fx
        y = ((ii - cy) * d) / fy
        w = [((P[r, 0] * x + P[r, 1] * y) + P[r, 2] * d) + P[r, 3] for r in range(3)]
        ul = F(self.voxel_length * F(UNIT_RES))
        lo = [np.floor((w[r] - self.sdf_trunc) / ul).astype(np.int64) for r in range(3)]
        hi = [np.floor((w[r] + self.sdf_trunc) / ul).astype(np.int64) for r in range(3)]
        out = set()
        span = max(int((hi[r] - lo[r]).max()) if len(d) else 0 for r in range(3))
        for ox in range(span + 1):
            for oy in range(span + 1):
                for oz in range(span + 1):
                    m = (lo[0] + ox <= hi[0]) & (lo[1] + oy <= hi[1]) & (lo[2] + oz <= hi[2])
                    if m.any():
                        u = np.stack([lo[0][m] + ox, lo[1][m] + oy, lo[2][m] + oz], axis=1)
                        out.update(map(tuple, np.unique(u, axis=0).tolist()))
        return sorted(out)

    # -- step 3
    def integrate(self, depth: np.ndarray, color: Optional[np.ndarray], fx, fy, cx, cy, extrinsic_4x4,
                  depth_trunc: float):
        """depth f32[H,W]; color u8[H,W,3] or f32[H,W,3] (values used as they are) or None."""
        depth = np.ascontiguousarray(depth, dtype=F)
        H, W = depth.shape
        E, P = frame_matrices(extrinsic_4x4)
        fx, fy, cx, cy = F(fx), F(fy), F(cx), F(cy)
        touched = self.touched_units(depth, fx, fy, cx, cy, P, depth_trunc)
        self.last_touched = touched
        vl, trunc = self.voxel_length, self.sdf_trunc
        inv_fx, inv_fy, inv_trunc = F(1) / fx, F(1) / fy, F(1) / trunc
        safe_w, safe_h = F(W) - F(0.0001), F(H) - F(0.0001)
        idx = np.arange(UNIT_RES)
        vx, vy, vz = np.meshgrid(idx, idx, idx, indexing="ij")
        for key in touched:
            unit = self.units.get(key)
            if unit is None:
                unit = {"tsdf": np.zeros((UNIT_RES,) * 3, F), "weight": np.zeros((UNIT_RES,) * 3, F),
                        "rgb": np.zeros((UNIT_RES,) * 3 + (3,), F)}
                self.units[key] = unit
            wx = ((key[0] * UNIT_RES + vx).astype(F) + F(0.5)) * vl
            wy = ((key[1] * UNIT_RES + vy).astype(F) + F(0.5)) * vl
            wz = ((key[2] * UNIT_RES + vz).astype(F) + F(0.5)) * vl
            pc = [((E[r, 0] * wx + E[r, 1] * wy) + E[r, 2] * wz) + E[r, 3] for r in range(3)]
            m = pc[2] > 0
            z = np.where(m, pc[2], F(1))
            u_f = ((pc[0] * fx) / z + cx) + F(0.5)
            v_f = ((pc[1] * fy) / z + cy) + F(0.5)
            m &= (u_f >= F(0.0001)) & (u_f < safe_w) & (v_f >= F(0.0001)) & (v_f < safe_h)
            u = np.where(m, u_f, 0).astype(np.int64)
            v = np.where(m, v_f, 0).astype(np.int64)
            d = depth[v, u]
            m &= (d > 0) & ~(d >= F(depth_trunc))
            xn = (u.astype(F) - cx) * inv_fx
            yn = (v.astype(F) - cy) * inv_fy
            mult = _sqrt32((xn * xn + yn * yn) + F(1))
            sdf = (d - z) * mult
            m &= sdf > -trunc
            t = np.minimum(F(1), sdf * inv_trunc)
            w0 = unit["weight"]
            w1 = w0 + F(1)
            unit["tsdf"] = np.where(m, (unit["tsdf"] * w0 + t) / w1, unit["tsdf"]).astype(F)
            if self.with_color and color is not None:
                c = color[v, u].astype(F)
                unit["rgb"] = np.where(m[..., None], (unit["rgb"] * w0[..., None] + c) / w1[..., None],
                                       unit["rgb"]).astype(F)
            unit["weight"] = np.where(m, w1, w0).astype(F)
        return touched
