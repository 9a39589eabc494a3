"""TSDF fusion of rendered frames without leaving the device.

Device-side mirror of the part of Open3D the reference's meshing exporter uses per frame
(collab_splats/utils/mesh.py:1562-1632):

    volume = o3d.pipelines.integration.ScalableTSDFVolume(voxel_length=..., sdf_trunc=..., color_type=RGB8)
    rgbd = o3d.geometry.RGBDImage.create_from_color_and_depth(color_u8, depth, depth_trunc=..., depth_scale=1.0,
                                                              convert_rgb_to_intensity=False)
    volume.integrate(rgbd, intrinsic=PinholeCameraIntrinsic(W, H, fx, fy, cx, cy), extrinsic=inv(c2w))

Here the rendered depth / colour tensors stay in HBM (the reference copies both to the host every frame,
mesh.py:1612-1620) and ``rs_tsdf_integrate`` (csrc/tsdf.cu) fuses them into a hashed pool of 16^3-voxel units.
There is no CPU path: CPU tensors raise ``RuntimeError``.
"""

from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch
from torch import Tensor

UNIT_RES = 16
UNIT_VOXELS = UNIT_RES ** 3


class TSDFVolumeColorType:
    """Names of o3d.pipelines.integration.TSDFVolumeColorType that the reference can pass (mesh.py:1566)."""
    NoColor = 0
    RGB8 = 1


class PinholeCameraIntrinsic:
    """o3d.camera.PinholeCameraIntrinsic as used at mesh.py:1598-1605."""

    def __init__(self, width: int, height: int, fx: float, fy: float, cx: float, cy: float):
        self.width, self.height = int(width), int(height)
        self.fx, self.fy, self.cx, self.cy = float(fx), float(fy), float(cx), float(cy)


class ScalableTSDFVolume:
    """Same constructor arguments and ``integrate`` / ``reset`` meaning as Open3D's class; the volume is device
    memory owned by this object (``max_units`` volume units of 16^3 voxels, 80 KB each with colour)."""

    def __init__(self, voxel_length: float, sdf_trunc: float, color_type: int = TSDFVolumeColorType.RGB8,
                 volume_unit_resolution: int = UNIT_RES, depth_sampling_stride: int = 4,
                 max_units: int = 32768, device="cuda"):
        if volume_unit_resolution != UNIT_RES:
            raise NotImplementedError("volume_unit_resolution is fixed at 16 (Open3D's default; the reference "
                                      "never changes it, mesh.py:1563-1567)")
        if color_type not in (TSDFVolumeColorType.NoColor, TSDFVolumeColorType.RGB8):
            raise NotImplementedError("color_type must be NoColor or RGB8 (mesh.py:1566 uses RGB8)")
        self.voxel_length, self.sdf_trunc = float(voxel_length), float(sdf_trunc)
        self.color_type = color_type
        self.depth_sampling_stride = int(depth_sampling_stride)
        self.max_units = int(max_units)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("ScalableTSDFVolume lives in device memory: there is no CPU path")
        cap = 1
        while cap < 2 * self.max_units:
            cap *= 2
        self.capacity = cap
        self.reset()

    def reset(self):
        dev, mu = self.device, self.max_units
        self.keys = torch.full((self.capacity,), -1, dtype=torch.int64, device=dev)      # all 0xFF
        self.vals = torch.full((self.capacity,), -1, dtype=torch.int32, device=dev)
        self.stamps = torch.zeros(self.capacity, dtype=torch.int32, device=dev)
        self.counters = torch.zeros(4, dtype=torch.int32, device=dev)
        self.unit_xyz = torch.zeros(mu, 3, dtype=torch.int32, device=dev)
        self.touched = torch.zeros(mu, dtype=torch.int32, device=dev)
        self.tsdf = torch.zeros(mu, UNIT_VOXELS, dtype=torch.float32, device=dev)
        self.weight = torch.zeros(mu, UNIT_VOXELS, dtype=torch.float32, device=dev)
        self.rgb = (torch.zeros(mu, UNIT_VOXELS, 3, dtype=torch.float32, device=dev)
                    if self.color_type == TSDFVolumeColorType.RGB8 else None)
        self.frames = 0

    def integrate(self, depth: Tensor, color: Optional[Tensor], intrinsic: PinholeCameraIntrinsic, extrinsic,
                  depth_trunc: float = 3.0):
        """One frame.  ``depth`` f32 [H,W] or [H,W,1] (0 = no measurement, values >= depth_trunc dropped, which is
        what ``create_from_color_and_depth(depth_trunc=...)`` does before Open3D integrates); ``color`` uint8
        [H,W,3] (what mesh.py:1612-1616 builds with ``* 255`` and a uint8 cast) or float32 [H,W,3] already scaled
        to 0..255; ``extrinsic`` world->camera 4x4 (host array or tensor).  Nothing is copied to the host and the
        call does not synchronise."""
        from radegs_b200 import backend as be
        lib = be.load()
        if not depth.is_cuda:
            raise RuntimeError("ScalableTSDFVolume.integrate needs CUDA tensors: there is no CPU path")
        if depth.dim() == 3:
            depth = depth.squeeze(-1)
        depth = depth.detach().to(torch.float32).contiguous()
        H, W = depth.shape
        if (W, H) != (intrinsic.width, intrinsic.height):
            raise ValueError(f"depth is {W}x{H} but the intrinsic says {intrinsic.width}x{intrinsic.height}")
        c8 = c32 = None
        if self.rgb is not None:
            if color is None:
                raise ValueError("this volume was created with color_type=RGB8: integrate() needs a colour image")
            if tuple(color.shape) != (H, W, 3):
                raise ValueError(f"colour must be [H,W,3], got {tuple(color.shape)}")
            color = color.detach().contiguous()
            if color.dtype == torch.uint8:
                c8 = color
            elif color.dtype == torch.float32:
                c32 = color
            else:
                raise ValueError("colour must be uint8 or float32")
        if isinstance(extrinsic, torch.Tensor):
            extrinsic = extrinsic.detach().cpu().numpy()
        e64 = np.asarray(extrinsic, dtype=np.float64).reshape(4, 4)
        E = np.ascontiguousarray(e64[:3], dtype=np.float32)
        P = np.ascontiguousarray(np.linalg.inv(e64)[:3], dtype=np.float32)
        self.frames += 1
        with torch.cuda.device(self.device):
            be.check(lib.rs_tsdf_integrate(
                be.ptr(depth), be.ptr(c8), be.ptr(c32), W, H, intrinsic.fx, intrinsic.fy, intrinsic.cx, intrinsic.cy,
                E.ctypes.data, P.ctypes.data, self.voxel_length, self.sdf_trunc, float(depth_trunc),
                self.depth_sampling_stride, self.frames, be.ptr(self.keys), be.ptr(self.vals), self.capacity,
                be.ptr(self.counters), self.max_units, be.ptr(self.unit_xyz), be.ptr(self.stamps),
                be.ptr(self.touched), be.ptr(self.tsdf), be.ptr(self.weight), be.ptr(self.rgb),
                be.stream_ptr(self.device)), "rs_tsdf_integrate")

    # ------------------------------------------------------------------ read-back (synchronises)
    def n_units(self) -> int:
        c = self.counters.cpu()
        if int(c[2]) != 0:
            raise RuntimeError(f"TSDF volume overflow: more than max_units={self.max_units} volume units (or unit "
                               "coordinates beyond +-2^20) were touched; create the volume with a larger max_units")
        return int(c[0])

    def units(self) -> Dict[str, Tensor]:
        """Allocated units: ``xyz`` i32 [U,3] (unit coordinates; voxel (x,y,z) of a unit has world centre
        ``((16*unit + (x,y,z)) + 0.5) * voxel_length``), ``tsdf`` / ``weight`` [U,16,16,16], ``rgb`` [U,16,16,16,3]."""
        U = self.n_units()
        out = {"xyz": self.unit_xyz[:U], "tsdf": self.tsdf[:U].view(U, UNIT_RES, UNIT_RES, UNIT_RES),
               "weight": self.weight[:U].view(U, UNIT_RES, UNIT_RES, UNIT_RES)}
        if self.rgb is not None:
            out["rgb"] = self.rgb[:U].view(U, UNIT_RES, UNIT_RES, UNIT_RES, 3)
        return out

    def extract_voxel_point_cloud(self, weight_min: float = 0.0, tsdf_abs_max: float = 0.98):
        """Open3D's ``extract_voxel_point_cloud``: centres (and colours / 255) of the voxels with weight > 0 and
        |tsdf| < 0.98, on the device."""
        u = self.units()
        m = (u["weight"] > weight_min) & (u["tsdf"].abs() < tsdf_abs_max)
        idx = m.nonzero(as_tuple=False)                                   # [K,4] = unit, x, y, z
        pts = ((u["xyz"][idx[:, 0]].to(torch.float32) * UNIT_RES + idx[:, 1:].to(torch.float32)) + 0.5) \
            * self.voxel_length
        out = {"points": pts, "tsdf": u["tsdf"][m]}
        if self.rgb is not None:
            out["colors"] = u["rgb"][m] / 255.0
        return out


def fuse_render_sweep(volume: ScalableTSDFVolume, gaussians, viewmats: Tensor, Ks: Tensor, width: int, height: int,
                      sh_degree: Optional[int] = 3, depth_name: str = "depth", depth_trunc: float = 20.0,
                      background: Optional[Tensor] = None, rasterize_mode: str = "antialiased",
                      views_per_launch: int = 1) -> int:
    """The frame loop of ``Open3DTSDFFusion.main`` (collab_splats/utils/mesh.py:1571-1632) with every frame kept on
    the device: for each camera, render forward-only (``RadegsModel.get_outputs`` in eval mode: ``RGB+ED``,
    rade_gs_model.py:153-154,439-465), build the same ``rgb`` / ``depth`` outputs (rade_gs_model.py:227-262:
    background blend + clamp; depth where alpha > 0 else the frame's max) and integrate.

    ``gaussians`` = (means, quats, scales, opacities, colors) already activated, as ``rasterization`` takes them;
    ``viewmats`` [V,4,4] world->camera (the extrinsic the reference hands to Open3D is exactly this matrix:
    ``inv(c2w @ diag(1,-1,-1,1))``, mesh.py:1592-1596,1630); ``depth_name`` "depth" (expected) or "median_depth".
    ``views_per_launch`` > 1 renders that many cameras per ``rasterization`` call (the camera axis of the
    rasterizer; identical frames, fewer launches); frames are integrated in camera order either way.
    Returns the number of frames integrated.  No per-frame host synchronisation besides the rasterizer's own
    intersection-count read; the volume's overflow counter is checked once after the last frame (a full unit pool
    would otherwise drop units silently for the whole sweep).  Triangle extraction (the reference's next call,
    ``volume.extract_triangle_mesh()``, mesh.py:1632) is NOT provided: the marching-cubes tables are not available
    offline; ``ScalableTSDFVolume.units()`` / ``extract_voxel_point_cloud()`` export (xyz, tsdf, weight, rgb) for an
    Open3D or skimage marching-cubes pass on the host."""
    from gsplat.rendering import rasterization
    if depth_name not in ("depth", "median_depth"):
        raise ValueError("depth_name must be 'depth' or 'median_depth'")
    if views_per_launch < 1:
        raise ValueError("views_per_launch must be >= 1")
    means, quats, scales, opacities, colors = gaussians
    V = viewmats.shape[0]
    ext = viewmats.detach().double().cpu().numpy()
    Kh = Ks.detach().double().cpu().numpy()
    bg = background if background is not None else torch.zeros(3, device=means.device)
    with torch.no_grad():
        for v0 in range(0, V, views_per_launch):
            v1 = min(V, v0 + views_per_launch)
            render, alpha, exp_d, med_d, _, _ = rasterization(
                means, quats, scales, opacities, colors, viewmats[v0:v1], Ks[v0:v1], width, height, packed=False,
                sh_degree=sh_degree, render_mode="RGB+ED", rasterize_mode=rasterize_mode, return_depth_normal=True)
            rgb8 = (torch.clamp(render[..., :3] + (1 - alpha) * bg, 0.0, 1.0) * 255).to(torch.uint8)
            d = exp_d if depth_name == "depth" else med_d
            d = torch.where(alpha > 0, d, d.amax(dim=(1, 2, 3), keepdim=True))
            for v in range(v0, v1):
                intr = PinholeCameraIntrinsic(width, height, Kh[v, 0, 0], Kh[v, 1, 1], Kh[v, 0, 2], Kh[v, 1, 2])
                volume.integrate(d[v - v0], rgb8[v - v0], intr, ext[v], depth_trunc=depth_trunc)
    volume.n_units()          # one read-back at the END of the sweep: raises if the unit pool overflowed on any frame
    return V
