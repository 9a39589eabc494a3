"""A/B of the compositing variants (forward + backward) on BASELINE config 2 (run on a B200)."""
import sys
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
p = [t.to(dev).requires_grad_(True) for t in scenes.activate(gs, 3)]
vmd, Kd = vm.to(dev), Ks.to(dev)
cot = None
def step():
    global cot
    for t in p: t.grad = None
    o = rasterization(*p, vmd, Kd, cfg.width, cfg.height, packed=False, sh_degree=3, render_mode="RGB+ED",
                      rasterize_mode="antialiased", return_depth_normal=True)
    if cot is None:
        g = torch.Generator(device=dev).manual_seed(1)
        cot = [torch.randn(t.shape, device=dev, generator=g) for t in o[:5]]
    sum((a * b).sum() for a, b in zip(o[:5], cot)).backward()
    return [t.detach() for t in o[:5]], [t.grad.clone() for t in p]
ref = None
from gsplat.cuda import _wrapper as W
MMA = be.RS_RASTER_BWD_MMA
FR, BB = be.RS_RASTER_FWD_RING, be.RS_RASTER_BWD_BARRIER
VARIANTS = (("fwd barrier / bwd ring (default)", 0), ("fwd ring / bwd barrier", FR | BB), ("bwd ring, 4 stages", be.RS_RASTER_BWD_TUNE(2)),
            ("fwd barrier / bwd ring (default)", 0), ("fwd ring / bwd barrier", FR | BB), ("bwd ring, 4 stages", be.RS_RASTER_BWD_TUNE(2)),
            ("1px", be.RS_RASTER_ONE_PIXEL), ("2px mma bwd", MMA))
for name, flags in VARIANTS:
    W.RASTER_FLAGS = flags
    for _ in range(3): o, g = step()
    lib.rs_timing_enable(1)
    for _ in range(10): o, g = step()
    torch.cuda.synchronize()
    s = be.timing_collect(); lib.rs_timing_enable(0)
    if ref is None: ref = (o, g)
    err = max(float((a - b).abs().max()) for a, b in zip(o, ref[0]))
    gerr = max(float((a - b).abs().max() / (b.abs().max() + 1e-30)) for a, b in zip(g, ref[1]))
    print(f"{name:34s}: sh_fwd {s['rs_sh_colors_fwd'][0] / 10:.4f} sh_bwd {s['rs_sh_colors_bwd'][0] / 10:.4f} emit {s['rs_isect_emit_ordered'][0] / 10:.4f} fwd {s['rs_rasterize_fwd'][0] / 10:.4f} ms  bwd {s['rs_rasterize_bwd'][0] / 10:.4f} ms  "
          f"unpack {s['rs_unpack_geom_grad'][0] / 10:.4f} ms   max|out diff vs first| {err:.2e}  max rel grad diff {gerr:.2e}")
W.RASTER_FLAGS = 0
