"""A/B of the wide-row (rade-features, 3 + 64 channels) backward at BASELINE config 3: tensor-core colour-gradient
reduction (default; RS_RASTER_NO_COLOR_MMA switches it off) vs the SIMT row walk (0); gradients compared between the two."""
import json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "collab-splats_b200")); sys.path.insert(0, str(ROOT))
import torch
from gsplat.cuda import _wrapper as W
from radegs_b200 import backend as be, scenes
from gsplat.rendering import rasterization

dev = torch.device("cuda:0")
lib = be.load()
cfg = scenes.BASELINE_CONFIGS[3]
NF = int(sys.argv[1]) if len(sys.argv) > 1 else 64          # feature channels (config 3: 64 -> 3 + 64 + 1 = 68 rows)
gs = scenes.make_gaussians(cfg.n_gaussians, None, NF, cfg.seed)
vm, Ks = scenes.make_cameras(1, cfg.width, cfg.height, cfg.seed)
p = [t.to(dev).requires_grad_(True) for t in scenes.activate(gs, None)]
vmd, Kd = vm.to(dev), Ks.to(dev)
torch.manual_seed(0)
w = torch.randn(1, cfg.height, cfg.width, 3 + NF + 1, device=dev)


def step():
    for t in p: t.grad = None
    o = rasterization(*p, vmd, Kd, cfg.width, cfg.height, packed=False, render_mode="RGB+ED",
                      rasterize_mode="antialiased", return_depth_normal=True)
    ((o[0] * w).sum() + o[2].mean() + o[3].mean() + o[4].mean()).backward()


out = {"channels": 3 + NF + 1}
grads = {}
for mode in (0, 1, 0, 1):
    W.RASTER_FLAGS = 0 if mode else be.RS_RASTER_NO_COLOR_MMA
    for _ in range(3): step()
    lib.rs_timing_enable(1)
    for _ in range(5): step()
    torch.cuda.synchronize()
    s = be.timing_collect(); lib.rs_timing_enable(0)
    out.setdefault(f"color_mma_{mode}", []).append({k: round(v[0] / 5, 4) for k, v in s.items() if "rasterize" in k})
    grads[mode] = p[4].grad.clone()
d = (grads[1] - grads[0]).abs().max().item()
out["max_abs_diff_color_grad"] = d
out["color_grad_scale"] = grads[0].abs().max().item()
W.RASTER_FLAGS = 0
print(json.dumps(out, indent=1))
