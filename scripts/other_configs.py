"""Timings of the other BASELINE configs (parity-test cases, not bench lines) on one B200:
config 3 (rade-features, 500k Gaussians, 64 feature channels, 960x540, fwd+bwd),
config 4 slice (3M Gaussians, C views of 1080p in ONE call on one GPU, fwd+bwd),
config 5 (2M Gaussians, forward-only RGB+ED sweep, views/s)."""
import json, math, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "collab-splats_b200")); sys.path.insert(0, str(ROOT))
import torch
from radegs_b200 import backend as be, scenes
from gsplat.rendering import rasterization

dev = torch.device("cuda:0")
lib = be.load()
out = {}


def timed(fn, reps=8, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def spans(fn, n=3):
    lib.rs_timing_enable(1)
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    s = be.timing_collect()
    lib.rs_timing_enable(0)
    return {k: round(v[0] / n, 4) for k, v in sorted(s.items(), key=lambda kv: -kv[1][0])}


# ---- config 3: rade-features
cfg = scenes.BASELINE_CONFIGS[3]
gs, vm, Ks = scenes.make_scene(cfg, n_views=1)
p = [t.to(dev).requires_grad_(True) for t in scenes.activate(gs, None)]
vmd, Kd = vm.to(dev), Ks.to(dev)
def step3():
    for t in p: t.grad = None
    o = rasterization(*p, vmd, Kd, cfg.width, cfg.height, packed=False, render_mode="RGB+ED",
                      rasterize_mode="antialiased", return_depth_normal=True)
    (o[0].square().mean() + o[2].mean() + o[3].mean() + o[4].mean()).backward()
    return o
o = step3()
out["config3_features67_960x540_500k"] = {"fwd_bwd_ms": round(timed(step3), 3), "n_isects": int(o[5]["flatten_ids"].numel()),
                                         "channels": int(o[0].shape[-1]), "kernels_ms": spans(step3)}
del p, o
torch.cuda.empty_cache()

# ---- config 5: forward-only sweep, 2M Gaussians
cfg = scenes.BASELINE_CONFIGS[5]
gs, vm, Ks = scenes.make_scene(cfg, n_views=16)
p = [t.to(dev) for t in scenes.activate(gs, 3)]
vmd, Kd = vm.to(dev), Ks.to(dev)
state = {"i": 0}
def view5():
    i = state["i"] % 16; state["i"] += 1
    with torch.no_grad():
        return rasterization(*p, vmd[i:i + 1], Kd[i:i + 1], cfg.width, cfg.height, packed=False, sh_degree=3,
                             render_mode="RGB+ED", rasterize_mode="antialiased", return_depth_normal=True)
o = view5()
ms = timed(view5, reps=16)
out["config5_sweep_2M_1080p_fwd_only"] = {"ms_per_view": round(ms, 3), "views_per_s": round(1000.0 / ms, 1),
                                          "n_isects": int(o[5]["flatten_ids"].numel()), "kernels_ms": spans(view5, 4)}
# with the device->host copy of the median depth + rgb that the reference's sweep does per frame (mesh.py:1612-1620)
pin_d = torch.empty(cfg.height, cfg.width, dtype=torch.float32).pin_memory()
pin_c = torch.empty(cfg.height, cfg.width, 4, dtype=torch.float32).pin_memory()
def view5_d2h():
    o = view5()
    pin_d.copy_(o[3][0, ..., 0], non_blocking=True); pin_c.copy_(o[0][0], non_blocking=True)
    torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(16): view5_d2h()
out["config5_sweep_2M_1080p_fwd_only"]["views_per_s_with_d2h"] = round(16 / (time.perf_counter() - t0), 1)
del p, o
torch.cuda.empty_cache()

# ---- config 4 slice: 3M Gaussians, C views in one call on ONE GPU
cfg = scenes.BASELINE_CONFIGS[4]
gs, vm, Ks = scenes.make_scene(cfg, n_views=8)
p = [t.to(dev).requires_grad_(True) for t in scenes.activate(gs, 3)]
for C in (1, 2, 4):
    vmd, Kd = vm[:C].to(dev), Ks[:C].to(dev)
    def step4():
        for t in p: t.grad = None
        o = rasterization(*p, vmd, Kd, cfg.width, cfg.height, packed=False, sh_degree=3, render_mode="RGB+ED",
                          rasterize_mode="antialiased", return_depth_normal=True)
        (o[0].square().mean() + o[2].mean() + o[3].mean() + o[4].mean()).backward()
        return o
    o = step4()
    ms = timed(step4, reps=5)
    out[f"config4_3M_{C}views_one_gpu"] = {"fwd_bwd_ms": round(ms, 3), "ms_per_view": round(ms / C, 3),
                                          "n_isects": int(o[5]["flatten_ids"].numel())}
    del o
print(json.dumps(out, indent=1))
