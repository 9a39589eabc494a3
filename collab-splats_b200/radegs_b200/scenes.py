"""Seeded synthetic scenes and cameras.

The same generator feeds the CUDA path, the CPU oracle and the CPU baseline so that every
comparison is on identical inputs.  Nothing here touches a GPU: tensors are created on the
CPU with a seeded ``torch.Generator`` and moved by the caller.

Parameters are returned *pre-activation* exactly as the reference models hold them
(``collab_splats/models/rade_gs_model.py:110-122``): log-scales, opacity logits, SH
coefficients; ``activate()`` applies what ``RadegsModel._render`` applies before calling
``rasterization`` (``rade_gs_model.py:443-444``).
"""

from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional

import torch


@dataclass
class SceneConfig:
    name: str
    n_gaussians: int
    width: int
    height: int
    n_views: int = 1
    sh_degree: Optional[int] = 3     # None -> colours are sigmoid(features_dc) (sh0 path)
    n_features: int = 0              # extra distilled feature channels (rade-features)
    seed: int = 1234


# BASELINE.json configs 1..5 (SURVEY.md section 8 size table)
BASELINE_CONFIGS = {
    1: SceneConfig("cfg1_10k_256", 10_000, 256, 256, 1, 3, 0, 1235),
    2: SceneConfig("cfg2_1M_1080p", 1_000_000, 1920, 1080, 1, 3, 0, 1236),
    3: SceneConfig("cfg3_500k_feat64_540p", 500_000, 960, 540, 1, None, 64, 1237),
    4: SceneConfig("cfg4_3M_8view_1080p", 3_000_000, 1920, 1080, 8, 3, 0, 1238),
    5: SceneConfig("cfg5_2M_sweep_1080p", 2_000_000, 1920, 1080, 300, 3, 0, 1239),
}


def make_gaussians(n: int, sh_degree: Optional[int], n_features: int, seed: int,
                   dtype=torch.float32) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    means = torch.rand(n, 3, generator=g, dtype=dtype) * 2.0 - 1.0
    quats = torch.randn(n, 4, generator=g, dtype=dtype)              # un-normalised on purpose
    lo, hi = math.log(0.002), math.log(0.02)
    log_scales = torch.rand(n, 3, generator=g, dtype=dtype) * (hi - lo) + lo
    flat_axis = torch.randint(0, 3, (n,), generator=g)
    log_scales[torch.arange(n), flat_axis] += math.log(0.2)          # flat, surface-like splats
    opacity_logits = torch.randn(n, generator=g, dtype=dtype) * 1.5
    features_dc = torch.randn(n, 3, generator=g, dtype=dtype) * 0.5
    out = dict(means=means, quats=quats, log_scales=log_scales, opacity_logits=opacity_logits,
               features_dc=features_dc)
    if sh_degree is not None:
        k = (sh_degree + 1) ** 2
        out["features_rest"] = torch.randn(n, k - 1, 3, generator=g, dtype=dtype) * 0.05
    if n_features > 0:
        out["distill_features"] = torch.randn(n, n_features, generator=g, dtype=dtype) * 0.3
    return out


def make_cameras(n_views: int, width: int, height: int, seed: int, radius: float = 3.0,
                 dtype=torch.float32):
    """Look-at-origin cameras on a seeded ring, OpenCV axes (x right, y down, z forward);
    fx = fy = 0.9*W, centred principal point (the reference forces it, rade_gs_model.py:327-334).
    Returns viewmats [C,4,4] (world->camera) and Ks [C,3,3]."""
    g = torch.Generator().manual_seed(seed + 7919)
    phase = torch.rand(1, generator=g, dtype=torch.float64).item() * 2 * math.pi
    viewmats = torch.zeros(n_views, 4, 4, dtype=torch.float64)
    for i in range(n_views):
        ang = phase + 2 * math.pi * i / max(n_views, 1)
        hgt = (torch.rand(1, generator=g, dtype=torch.float64).item() * 2 - 1) * 0.3
        eye = torch.tensor([radius * math.cos(ang), hgt, radius * math.sin(ang)], dtype=torch.float64)
        fwd = -eye / eye.norm()
        up = torch.tensor([0.0, -1.0, 0.0], dtype=torch.float64)     # world "up" is -y (OpenCV y down)
        right = torch.linalg.cross(fwd, up)
        right = right / right.norm()
        down = torch.linalg.cross(fwd, right)
        Rcw = torch.stack([right, down, fwd], dim=0)                  # rows = camera axes in world
        viewmats[i, :3, :3] = Rcw
        viewmats[i, :3, 3] = -Rcw @ eye
        viewmats[i, 3, 3] = 1.0
    K = torch.tensor([[0.9 * width, 0.0, width / 2.0], [0.0, 0.9 * width, height / 2.0], [0.0, 0.0, 1.0]],
                     dtype=torch.float64)
    Ks = K[None].repeat(n_views, 1, 1)
    return viewmats.to(dtype), Ks.to(dtype)


def make_scene(cfg: SceneConfig, dtype=torch.float32, n_views: Optional[int] = None):
    gs = make_gaussians(cfg.n_gaussians, cfg.sh_degree, cfg.n_features, cfg.seed, dtype)
    viewmats, Ks = make_cameras(n_views if n_views is not None else cfg.n_views, cfg.width, cfg.height,
                                cfg.seed, dtype=dtype)
    return gs, viewmats, Ks


def activate(gs: Dict[str, torch.Tensor], sh_degree: Optional[int]):
    """What the reference model does right before ``rasterization`` (rade_gs_model.py:125-127,158-164,
    443-444; rade_features_model.py:441): returns (means, quats, scales, opacities, colors)."""
    scales = torch.exp(gs["log_scales"])
    opacities = torch.sigmoid(gs["opacity_logits"])
    if sh_degree is None:
        colors = torch.sigmoid(gs["features_dc"])
    else:
        colors = torch.cat([gs["features_dc"][:, None, :], gs["features_rest"]], dim=1)
    if "distill_features" in gs:
        assert sh_degree is None
        colors = torch.cat([colors, gs["distill_features"]], dim=-1)
    return gs["means"], gs["quats"], scales, opacities, colors
