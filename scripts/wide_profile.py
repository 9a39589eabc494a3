"""One config-3 (rade-features, 68 channels) fwd+bwd step between cudaProfilerStart/Stop, for ncu."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "collab-splats_b200")); sys.path.insert(0, str(ROOT))
import torch
from gsplat.cuda import _wrapper as W
from radegs_b200 import backend as be, scenes
from gsplat.rendering import rasterization
dev = torch.device("cuda:0")
lib = be.load()
if len(sys.argv) > 1:
    W.RASTER_FLAGS = 0 if int(sys.argv[1]) else be.RS_RASTER_NO_COLOR_MMA
cfg = scenes.BASELINE_CONFIGS[3]
gs, vm, Ks = scenes.make_scene(cfg, n_views=1)
p = [t.to(dev).requires_grad_(True) for t in scenes.activate(gs, None)]
vmd, Kd = vm.to(dev), Ks.to(dev)
torch.manual_seed(0)
w = torch.randn(1, cfg.height, cfg.width, 68, device=dev)
def step():
    for t in p: t.grad = None
    o = rasterization(*p, vmd, Kd, cfg.width, cfg.height, packed=False, render_mode="RGB+ED",
                      rasterize_mode="antialiased", return_depth_normal=True)
    ((o[0] * w).sum() + o[2].mean() + o[3].mean() + o[4].mean()).backward()
for _ in range(3): step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
