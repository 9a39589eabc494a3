"""The C compositing oracle (oracle/raster_oracle.c: per-pixel loops, hand-written chain rule) against the PyTorch
oracle (oracle/rade_oracle.py: vectorised per tile, autograd) on scenes small enough for the latter -- two independent
restatements of SURVEY.md rows a10/a11 that must agree before the C one is trusted at the BASELINE sizes
(tests/test_gpu_parity.py::test_full_view_matches_c_oracle, bench.py's CPU legs).  CPU only."""

import pytest
import torch

from oracle import rade_oracle as O
from oracle import raster_oracle as RO
from radegs_b200 import scenes
from tests.util import small_scene


def _stage_inputs(cfg, gs, vm, Ks, dtype, with_bg):
    params = [t.to(dtype) for t in scenes.activate(gs, cfg.sh_degree)]
    means, quats, scales, opac, colors = params
    vm, Ks = vm.to(dtype), Ks.to(dtype)
    with torch.no_grad():
        radii, m2, depths, conics, comps, ray_ts, ray_planes, normals = O.fully_fused_projection(
            means, quats, scales, vm, Ks, cfg.width, cfg.height, calc_compensations=True)
        C, N = depths.shape
        g = torch.Generator().manual_seed(5)
        cols = torch.rand(C, N, 4, generator=g, dtype=dtype)
        o = opac[None].expand(C, N) * comps
        tw, th = -(-cfg.width // 16), -(-cfg.height // 16)
        _, ids, flat = O.isect_tiles(m2, radii, depths, 16, tw, th)
        offs = O.isect_offset_encode(ids, C, tw, th)
        bg = torch.rand(C, 4, generator=g, dtype=dtype) if with_bg else None
    leaves = [t.detach().clone().requires_grad_(True) for t in (m2, conics, cols, o, ray_ts, ray_planes, normals)]
    return leaves, Ks, offs, flat, bg


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-11), (torch.float32, 2e-5)])
@pytest.mark.parametrize("n,w,h,views,with_bg", [(1500, 100, 70, 1, False), (2500, 64, 48, 2, True)])
def test_c_compositor_matches_torch_oracle(dtype, tol, n, w, h, views, with_bg):
    cfg, gs, vm, Ks = small_scene(n=n, w=w, h=h, views=views)
    leaves, Ks, offs, flat, bg = _stage_inputs(cfg, gs, vm, Ks, dtype, with_bg)
    bg_t = None if bg is None else bg.clone().requires_grad_(True)
    bg_c = None if bg is None else bg.clone().requires_grad_(True)
    ref = O.rasterize_to_pixels(*leaves[:4], *leaves[4:], Ks, w, h, 16, offs, flat, backgrounds=bg_t, return_aux=True)
    leaves_c = [t.detach().clone().requires_grad_(True) for t in leaves]
    got = RO.rasterize_to_pixels(*leaves_c[:4], *leaves_c[4:], Ks, w, h, 16, offs, flat, backgrounds=bg_c,
                                 return_aux=True, threads=3)
    aux_r, aux_g = ref[5], got[5]
    keep = ~(aux_r["fragile"] | aux_g["fragile"])
    assert float(keep.float().mean()) > 0.97
    assert float(ref[1].detach().mean()) > 0.2, "scene does not cover the image enough to be a meaningful test"
    # discrete artefacts: identical outside the pixels either oracle flags as sitting on a threshold
    assert bool(((aux_r["last_ids"] == aux_g["last_ids"]) | ~keep).all())
    assert bool(((aux_r["median_ids"] == aux_g["median_ids"]) | ~keep).all())
    if dtype == torch.float64:
        assert aux_r["n_tested"] == aux_g["n_tested"] and aux_r["n_contrib"] == aux_g["n_contrib"]
    gen = torch.Generator().manual_seed(9)
    loss_r, loss_g = 0.0, 0.0
    for a, b in zip(ref[:5], got[:5]):
        m = keep[..., None].expand_as(a)
        scale = float(a.abs().max()) + 1e-30
        assert float(((a - b).abs() * m).max()) <= tol * max(1.0, scale), (float(((a - b).abs() * m).max()), scale)
        wgt = torch.randn(a.shape, generator=gen, dtype=dtype) * m       # fragile pixels get no upstream gradient
        loss_r = loss_r + (a * wgt).sum()
        loss_g = loss_g + (b * wgt).sum()
    loss_r.backward()
    loss_g.backward()
    names = ["means2d", "conics", "colors", "opacities", "ray_ts", "ray_planes", "normals"]
    for nm, a, b in zip(names, leaves, leaves_c):
        scale = float(a.grad.abs().max()) + 1e-30
        err = float((a.grad - b.grad).abs().max())
        assert err <= 50 * tol * scale, (nm, err, scale)
    if bg is not None:
        assert float((bg_t.grad - bg_c.grad).abs().max()) <= 50 * tol * float(bg_t.grad.abs().max())


def test_c_compositor_through_rasterization_fp64():
    """The whole oracle call with compositor="c" against compositor="torch" (RGB+ED, antialiased, SH degree 3):
    checks the routing, the expected-depth normalisation on top of the C outputs and the end-to-end gradients."""
    cfg, gs, vm, Ks = small_scene(n=800, w=80, h=48, views=1)
    params = [t.double() for t in scenes.activate(gs, 3)]
    outs, grads = [], []
    for comp in ("torch", "c"):
        leaves = [t.detach().clone().requires_grad_(True) for t in params]
        r = O.rasterization(*leaves, vm.double(), Ks.double(), cfg.width, cfg.height, sh_degree=3,
                            render_mode="RGB+ED", rasterize_mode="antialiased", return_depth_normal=True,
                            return_aux=True, compositor=comp, threads=2)
        keep = ~r[5]["fragile"]
        gen = torch.Generator().manual_seed(2)
        loss = 0.0
        for o in r[:5]:
            loss = loss + (o * torch.randn(o.shape, generator=gen, dtype=torch.float64) * keep[..., None]).sum()
        loss.backward()
        outs.append([o.detach() for o in r[:5]])
        grads.append([t.grad for t in leaves])
    for a, b in zip(outs[0], outs[1]):
        assert float((a - b).abs().max()) <= 1e-10 * max(1.0, float(a.abs().max()))
    for a, b in zip(grads[0], grads[1]):
        assert float((a - b).abs().max()) <= 1e-9 * (float(a.abs().max()) + 1e-30)
