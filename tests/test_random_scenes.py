"""Randomised tiny scenes, kernel <-> oracle: hypothesis draws the shape parameters the fixed
parity tests do not sweep -- ragged image sizes (not multiples of the 16-pixel tile), 1-3 cameras, every SH degree /
colour layout, render and rasterize modes, backgrounds, camera distance (near-plane culls, huge and sub-pixel splats),
degenerate scales -- and the whole ``rasterization()`` call is compared with the oracle: integer artefacts bit for
bit, images within max-abs 1e-4 + rel 1e-3 outside the oracle's fragile pixels, gradients within 3e-3 of their scale.
Derandomised (fixed example sequence) so the GPU run is reproducible."""

import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import rade_oracle as O
from radegs_b200 import scenes
from tests.util import close_report, grad_close_report

pytestmark = pytest.mark.gpu


@st.composite
def scene_params(draw):
    return dict(
        n=draw(st.integers(1, 300)), w=draw(st.integers(1, 90)), h=draw(st.integers(1, 70)),
        views=draw(st.integers(1, 3)), sh_degree=draw(st.sampled_from([None, 0, 1, 2, 3])),
        n_features=draw(st.sampled_from([0, 0, 5, 29])), mode=draw(st.sampled_from(["RGB", "RGB+ED", "RGB+D", "ED", "D"])),
        raster=draw(st.sampled_from(["classic", "antialiased"])), bg=draw(st.booleans()),
        boost=draw(st.floats(1.0, 4.5)), radius=draw(st.sampled_from([0.6, 1.5, 3.0, 8.0])),
        flat_axis=draw(st.booleans()), seed=draw(st.integers(0, 10_000)))


@settings(max_examples=60, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(p=scene_params())
def test_random_tiny_scenes_match_oracle(cuda_dev, p):
    from gsplat.rendering import rasterization
    sh = p["sh_degree"] if p["n_features"] == 0 else None          # N-D features go with pre-evaluated colours
    gs = scenes.make_gaussians(p["n"], sh, p["n_features"], p["seed"])
    vm, Ks = scenes.make_cameras(p["views"], p["w"], p["h"], p["seed"], radius=p["radius"])
    gs["log_scales"] = gs["log_scales"] + p["boost"]
    if p["flat_axis"]:
        gs["log_scales"][:, 2] = -9.0                              # degenerate (disc-like) Gaussians, 1.2e-4 thick
    params = scenes.activate(gs, sh)
    W, H, C = p["w"], p["h"], p["views"]
    g = torch.Generator().manual_seed(p["seed"])
    D = 3 if sh is not None else params[4].shape[-1]
    bg = torch.rand(C, D, generator=g) if (p["bg"] and p["mode"] not in ("D", "ED")) else None
    cpu = [t.detach().clone().requires_grad_(True) for t in params]
    ref = O.rasterization(*cpu, vm, Ks, W, H, sh_degree=sh, render_mode=p["mode"], rasterize_mode=p["raster"],
                          backgrounds=bg, return_depth_normal=True, return_aux=True)
    gpu = [t.detach().clone().to(cuda_dev).requires_grad_(True) for t in params]
    got = rasterization(*gpu, vm.to(cuda_dev), Ks.to(cuda_dev), W, H, packed=False, sh_degree=sh,
                        render_mode=p["mode"], rasterize_mode=p["raster"],
                        backgrounds=None if bg is None else bg.to(cuda_dev), return_depth_normal=True)
    meta, rmeta = got[5], ref[5]
    for key in ("radii", "tiles_per_gauss", "isect_ids", "flatten_ids", "isect_offsets"):
        assert torch.equal(meta[key].cpu(), rmeta[key]), (key, p)
    keep = ~rmeta["fragile"]
    assert int((~keep).sum()) <= max(4, keep.numel() // 50), ("too many fragile pixels to be a meaningful test", p)
    for i, nm in enumerate(["render", "alpha", "expected_depths", "median_depths", "expected_normals"]):
        assert got[i].shape == ref[i].shape, (nm, got[i].shape, ref[i].shape)
        ok, msg = close_report(nm, got[i], ref[i], mask=keep)
        assert ok, (msg, p)
    ws = [torch.randn(t.shape, generator=g) * keep[..., None] for t in ref[:5]]
    loss_ref = sum((t * w).sum() for t, w in zip(ref[:5], ws))
    sum((t * w.to(cuda_dev)).sum() for t, w in zip(got[:5], ws)).backward()
    if loss_ref.requires_grad:
        loss_ref.backward()
    else:                                  # nothing visible: the oracle returns constants, the gradients must be zero
        assert rmeta["isect_ids"].numel() == 0
    # thin discs are ill-conditioned in fp32 (Sigma^-1 holds 1/thickness^2 ~ 7e7; at exp(-12) the fp64 oracle sits
    # between the fp32 oracle and the kernel, each ~0.5-1 % off for the thin axis): looser geometry bound there
    rel_geo = 2e-2 if p["flat_axis"] else 3e-3
    for nm, a, b in zip(("means", "quats", "scales", "opacities", "colors"), gpu, cpu):
        ok, msg = grad_close_report("v_" + nm, a.grad, b.grad, rel=rel_geo if nm in ("means", "quats", "scales") else 3e-3,
                                    floor=1e-5)
        assert ok, (msg, p)
