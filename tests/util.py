"""Shared helpers for the parity tests."""

from __future__ import annotations

import torch

from radegs_b200 import scenes

ATOL = 1e-4   # north_star tolerance: max-abs 1e-4 ...
RTOL = 1e-3   # ... rel 1e-3


def small_scene(n=3000, w=160, h=96, views=1, sh_degree=3, n_features=0, seed=11, scale_boost=1.2, spread=1.0):
    """A scene sized so the torch oracle finishes in seconds and the image is well covered."""
    cfg = scenes.SceneConfig("test", n, w, h, views, sh_degree, n_features, seed)
    gs, vm, Ks = scenes.make_scene(cfg)
    gs["log_scales"] = gs["log_scales"] + scale_boost
    gs["means"] = gs["means"] * spread
    return cfg, gs, vm, Ks


def close_report(name, got, ref, atol=ATOL, rtol=RTOL, mask=None):
    """Returns (ok, message) for |got-ref| <= atol + rtol*|ref| over `mask`."""
    got = got.detach().double().cpu()
    ref = ref.detach().double().cpu()
    assert got.shape == ref.shape, (name, got.shape, ref.shape)
    err = (got - ref).abs()
    tol = atol + rtol * ref.abs()
    bad = err > tol
    if mask is not None:
        m = mask
        while m.dim() < bad.dim():
            m = m[..., None]
        bad = bad & m.expand_as(bad)
    n_bad = int(bad.sum())
    msg = (f"{name}: max_abs={err.max().item():.3e} ref_max={ref.abs().max().item():.3e} "
           f"violations={n_bad}/{bad.numel()}")
    if n_bad:
        idx = torch.nonzero(bad)[:5]
        for i in idx:
            t = tuple(i.tolist())
            msg += f"\n    at {t}: got {got[t].item():.6e} ref {ref[t].item():.6e}"
    return n_bad == 0, msg


def grad_close_report(name, got, ref, rel=2e-3, floor=1e-6):
    """Gradient check: |got-ref| <= rel*max|ref| + floor (atomics reorder sums, so a global scale).
    A `None` gradient (input not used by the graph) counts as zeros."""
    if ref is None and got is None:
        return True, f"{name}: both None"
    if ref is None:
        ref = torch.zeros_like(got)
    if got is None:
        got = torch.zeros_like(ref)
    got = got.detach().double().cpu()
    ref = ref.detach().double().cpu()
    assert got.shape == ref.shape, (name, got.shape, ref.shape)
    err = (got - ref).abs()
    scale = ref.abs().max().item()
    tol = rel * scale + floor
    n_bad = int((err > tol).sum())
    msg = f"{name}: max_abs_err={err.max().item():.3e} scale={scale:.3e} tol={tol:.3e} violations={n_bad}/{err.numel()}"
    if n_bad:
        idx = torch.nonzero(err > tol)[:5]
        for i in idx:
            t = tuple(i.tolist())
            msg += f"\n    at {t}: got {got[t].item():.6e} ref {ref[t].item():.6e}"
    return n_bad == 0, msg
