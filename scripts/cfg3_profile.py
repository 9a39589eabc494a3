"""Every GPU kernel of one BASELINE config-3 step (rade-features: 500 k Gaussians, 3 + 64 channels + depth, 960x540),
torch ops included, by total device time (torch.profiler)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "collab-splats_b200")); sys.path.insert(0, str(ROOT))
import torch
from torch.profiler import profile, ProfilerActivity
from radegs_b200 import scenes
from gsplat.rendering import rasterization
dev = torch.device("cuda:0")
cfg = scenes.BASELINE_CONFIGS[3]
gs, vm, Ks = scenes.make_scene(cfg, n_views=1)
p = [t.to(dev).requires_grad_(True) for t in scenes.activate(gs, None)]
vmd, Kd = vm.to(dev), Ks.to(dev)
def step3():
    for t in p: t.grad = None
    o = rasterization(*p, vmd, Kd, cfg.width, cfg.height, packed=False, render_mode="RGB+ED",
                      rasterize_mode="antialiased", return_depth_normal=True)
    (o[0].square().mean() + o[2].mean() + o[3].mean() + o[4].mean()).backward()
for _ in range(3): step3()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(5): step3()
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / 5, e.count / 5) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f"total device time per step {tot / 1e3:.3f} ms")
for k, t, c in rows[:28]:
    print(f"{t:9.1f} us  x{c:4.1f}  {k[:110]}")
