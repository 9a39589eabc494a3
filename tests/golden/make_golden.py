"""Generates tests/golden/rade_small.npz: a tiny seeded scene, its oracle outputs and oracle gradients.

The reference holds no golden vectors for this path and gsplat-rade cannot be imported here, so
these fixtures come from the CPU oracle (oracle/rade_oracle.py) at the commit that pinned it against the analytic
known-answer tests.  They guard the oracle against drift (CPU test) and give the GPU tests a fixture that does
not depend on running the oracle.  Re-generate with:  python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT / "collab-splats_b200"))
sys.path.insert(0, str(ROOT))

from oracle import rade_oracle as O          # noqa: E402
from radegs_b200 import scenes               # noqa: E402

W, H, N, C, SH = 64, 48, 400, 2, 2


def build():
    cfg = scenes.SceneConfig("golden", N, W, H, C, SH, 0, 4242)
    gs, vm, Ks = scenes.make_scene(cfg)
    gs["log_scales"] = gs["log_scales"] + 2.3
    params = [t.detach().clone().requires_grad_(True) for t in scenes.activate(gs, SH)]
    out = O.rasterization(*params, vm, Ks, W, H, sh_degree=SH, render_mode="RGB+ED", rasterize_mode="antialiased",
                          return_depth_normal=True, return_aux=True)
    rc, ra, de, dm, nr, meta = out
    g = torch.Generator().manual_seed(99)
    keep = ~meta["fragile"]
    ws = [torch.randn(t.shape, generator=g) * keep[..., None] for t in (rc, ra, de, dm, nr)]
    sum((t * w).sum() for t, w in zip((rc, ra, de, dm, nr), ws)).backward()
    d = dict(viewmats=vm, Ks=Ks, means=params[0], quats=params[1], scales=params[2], opacities=params[3],
             sh_coeffs=params[4], render=rc, alphas=ra, expected_depths=de, median_depths=dm, normals=nr,
             radii=meta["radii"], means2d=meta["means2d"], depths=meta["depths"], conics=meta["conics"],
             tiles_per_gauss=meta["tiles_per_gauss"], isect_ids=meta["isect_ids"], flatten_ids=meta["flatten_ids"],
             isect_offsets=meta["isect_offsets"], fragile=meta["fragile"], last_ids=meta["last_ids"],
             w_render=ws[0], w_alphas=ws[1], w_expected_depths=ws[2], w_median_depths=ws[3], w_normals=ws[4],
             g_means=params[0].grad, g_quats=params[1].grad, g_scales=params[2].grad, g_opacities=params[3].grad,
             g_sh_coeffs=params[4].grad)
    return {k: v.detach().numpy() for k, v in d.items()}


if __name__ == "__main__":
    data = build()
    out = Path(__file__).resolve().parent / "rade_small.npz"
    np.savez_compressed(out, **data)
    print(out, out.stat().st_size, "bytes;", "M =", data["isect_ids"].shape[0], "alpha mean", data["alphas"].mean())
