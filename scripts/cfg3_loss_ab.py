import json, sys
sys.path.insert(0, "collab-splats_b200"); sys.path.insert(0, ".")
import torch
from gsplat.cuda import _wrapper as W
from radegs_b200 import backend as be, scenes
from gsplat.rendering import rasterization
dev = torch.device("cuda:0"); lib = be.load()
cfg = scenes.BASELINE_CONFIGS[3]
gs, vm, Ks = scenes.make_scene(cfg, n_views=1)
p = [t.to(dev).requires_grad_(True) for t in scenes.activate(gs, None)]
vmd, Kd = vm.to(dev), Ks.to(dev)
w = torch.randn(1, cfg.height, cfg.width, 68, device=dev)
def step(kind):
    for t in p: t.grad = None
    o = rasterization(*p, vmd, Kd, cfg.width, cfg.height, packed=False, render_mode="RGB+ED", rasterize_mode="antialiased", return_depth_normal=True)
    if kind == "sq": (o[0].square().mean() + o[2].mean() + o[3].mean() + o[4].mean()).backward()
    else: ((o[0] * w).sum() + o[2].mean() + o[3].mean() + o[4].mean()).backward()
for kind in ("sq", "rand"):
    for mode in (0, 1):
        W.RASTER_FLAGS = 0 if mode else be.RS_RASTER_NO_COLOR_MMA
        for _ in range(3): step(kind)
        lib.rs_timing_enable(1)
        for _ in range(5): step(kind)
        torch.cuda.synchronize()
        s = be.timing_collect(); lib.rs_timing_enable(0)
        print(kind, mode, {k: round(v[0] / 5, 4) for k, v in s.items() if "rasterize" in k})
