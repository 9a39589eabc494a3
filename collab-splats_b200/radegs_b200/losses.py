"""Model-side glue that runs right after ``rasterization`` every training step.

Mirrors, on device tensors and without the per-step host work of the reference (camera matrices rebuilt on
the CPU and pixel grids uploaded each step, collab_splats/utils/camera_utils.py:54-71,228-238):

* ``depth_double_to_normal``  -- collab_splats/utils/camera_utils.py:176-279
* ``depth_normal_loss``       -- collab_splats/models/rade_gs_model.py:202-219 and :292-307
  (lambda 0.05, depth_ratio 0.6: collab_splats/configs/rade_gs_method.py:38-40)

SURVEY.md section 8(f) row f1 lists this stencil as the first "next" component (a fused kernel); this module is
its plain device-side form and the reference for that fusion.
"""

from __future__ import annotations

import torch
from torch import Tensor


def depth_double_to_normal(Ks_c: Tensor, width: int, height: int, depth1: Tensor, depth2: Tensor) -> Tensor:
    """Two z-depth maps [H,W] of one pinhole camera with centred principal point -> normals [2,H,W,3];
    central differences, border rows/columns are zero (camera_utils.py:253-279)."""
    dt, dev = depth1.dtype, depth1.device
    fx, fy = Ks_c[0, 0], Ks_c[1, 1]
    gx = (torch.arange(width, dtype=dt, device=dev) + 0.5)[None, :].expand(height, width)
    gy = (torch.arange(height, dtype=dt, device=dev) + 0.5)[:, None].expand(height, width)
    rays = torch.stack([gx / fx - width / (2 * fx), gy / fy - height / (2 * fy), torch.ones_like(gx)], dim=0)
    pts = torch.stack([depth1[None] * rays, depth2[None] * rays], dim=0)      # [2,3,H,W]
    out = torch.zeros_like(pts)
    d_row = pts[..., 2:, 1:-1] - pts[..., :-2, 1:-1]
    d_col = pts[..., 1:-1, 2:] - pts[..., 1:-1, :-2]
    out[..., 1:-1, 1:-1] = torch.nn.functional.normalize(torch.cross(d_row, d_col, dim=1), dim=1)
    return out.permute(0, 2, 3, 1)


def depth_normal_loss(Ks_c: Tensor, width: int, height: int, expected_depth: Tensor, median_depth: Tensor,
                      rendered_normals: Tensor, lam: float = 0.05, depth_ratio: float = 0.6):
    """expected/median depth [H,W], rendered normals [H,W,3] -> (loss, error maps [2,H,W])."""
    n_d = depth_double_to_normal(Ks_c, width, height, expected_depth, median_depth)
    err = 1.0 - (rendered_normals[None] * n_d).sum(dim=-1)
    loss = lam * ((1.0 - depth_ratio) * err[0].mean() + depth_ratio * err[1].mean())
    return loss, err
