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


def grad_parity_report(name, got, ref64, ref32=None, atol_rel=1e-4, rtol=RTOL, factor=4.0, floor_rel=2e-6,
                       outlier_frac=0.0, outlier_atol_rel=2e-3):
    """Gradient parity in the north star's form, per element, against the fp64 oracle:

        |got - ref64| <= atol_rel * s + rtol * |ref64|        with s = max |ref64| (the tensor's own unit)

    i.e. "max-abs 1e-4, rel 1e-3" with the absolute part expressed in units of the tensor's largest gradient (a
    gradient has no natural unit of its own: it scales with the loss).  When the fp32 oracle's gradient is given as
    well, the GPU's error must also stay within what fp32 arithmetic itself does to this sum:

        rms(got - ref64) <= factor * rms(ref32 - ref64) + floor_rel * s,   max likewise with 2 * factor

    (atomics reorder the sums and the kernels use ex2.approx / rcp.approx, so a small multiple, not equality).
    ``outlier_frac`` > 0 tolerates that fraction of elements outside the per-element bound as long as they stay
    within ``outlier_atol_rel * s``: at a million Gaussians a handful of pixels take a discrete decision differently in
    fp32 and fp64 despite the margins, which moves single Gaussians' gradients by ~1e-3 of their size.
    The message also carries the worst relative error over the elements with |ref64| > 1e-3 * s."""
    if got is None:
        got = torch.zeros_like(ref64)
    got = got.detach().double().cpu()
    ref64 = ref64.detach().double().cpu()
    assert got.shape == ref64.shape, (name, got.shape, ref64.shape)
    s = ref64.abs().max().item()
    err = (got - ref64).abs()
    bad = err > atol_rel * s + rtol * ref64.abs()
    n_bad = int(bad.sum())
    big = ref64.abs() > 1e-3 * s
    worst_rel = (err[big] / ref64.abs()[big]).max().item() if bool(big.any()) else 0.0
    msg = (f"{name}: scale={s:.3e} max_abs_err={err.max().item():.3e} ({err.max().item() / (s + 1e-300):.2e} of scale) "
           f"worst_rel(|ref|>1e-3 s)={worst_rel:.2e} violations={n_bad}/{err.numel()}")
    ok = n_bad <= int(outlier_frac * err.numel()) and (n_bad == 0 or err.max().item() <= outlier_atol_rel * s)
    if ref32 is not None:
        e32 = (ref32.detach().double().cpu() - ref64).abs()
        rms_g, rms_o = err.pow(2).mean().sqrt().item(), e32.pow(2).mean().sqrt().item()
        max_g, max_o = err.max().item(), e32.max().item()
        msg += (f" | rms err gpu {rms_g:.3e} vs fp32 oracle {rms_o:.3e}; max err gpu {max_g:.3e} vs fp32 oracle "
                f"{max_o:.3e}")
        if rms_g > factor * rms_o + floor_rel * s or max_g > 2 * factor * max_o + 10 * floor_rel * s:
            ok = False
            msg += "  <-- GPU error exceeds the fp32 oracle's own error budget"
    if n_bad:
        for i in torch.nonzero(bad)[:5]:
            t = tuple(i.tolist())
            msg += f"\n    at {t}: got {got[t].item():.6e} ref64 {ref64[t].item():.6e}"
    return ok, msg
