"""A/B of the forward compositing variants on BASELINE config 2 (run on a B200)."""
import math, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "collab-splats_b200")); sys.path.insert(0, str(ROOT))
import torch
from radegs_b200 import backend as be, scenes
from gsplat.rendering import rasterization

lib = be.load()
dev = torch.device("cuda:0")
cfg = scenes.BASELINE_CONFIGS[2]
gs, vm, Ks = scenes.make_scene(cfg, n_views=1)
p = [t.to(dev) for t in scenes.activate(gs, 3)]
vmd, Kd = vm.to(dev), Ks.to(dev)
def fwd():
    with torch.no_grad():
        return rasterization(*p, vmd, Kd, cfg.width, cfg.height, packed=False, sh_degree=3, render_mode="RGB+ED",
                             rasterize_mode="antialiased", return_depth_normal=True)
ref = None
for variant in (0, 1, 0, 1):
    lib.rs_raster_set_variant(variant)
    for _ in range(3): o = fwd()
    lib.rs_timing_enable(1)
    for _ in range(10): o = fwd()
    torch.cuda.synchronize()
    s = be.timing_collect(); lib.rs_timing_enable(0)
    if ref is None: ref = [t.clone() for t in o[:5]]
    err = max(float((a - b).abs().max()) for a, b in zip(o[:5], ref))
    print(f"variant {variant}: rs_rasterize_fwd {s['rs_rasterize_fwd'][0] / 10:.4f} ms   max|diff vs variant 0| {err:.2e}")
lib.rs_raster_set_variant(0)
