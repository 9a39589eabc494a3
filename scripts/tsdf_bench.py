"""Config-5 sweep with TSDF fusion on the device: per-view time of render (2M Gaussians, 1080p, RGB+ED fwd) +
rs_tsdf_integrate, next to the reference-shaped loop (frame copied to the host every view)."""
import json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "collab-splats_b200")); sys.path.insert(0, str(ROOT))
import torch
from radegs_b200 import backend as be, scenes, tsdf

dev = torch.device("cuda:0")
lib = be.load()
n_views = int(sys.argv[1]) if len(sys.argv) > 1 else 32
cfg = scenes.BASELINE_CONFIGS[5]
gs, vm, Ks = scenes.make_scene(cfg, n_views=n_views)
p = [t.to(dev) for t in scenes.activate(gs, 3)]
vmd, Kd = vm.to(dev), Ks.to(dev)
out = {}
for voxel, trunc, vpl in ((0.01, 0.03, 1), (0.01, 0.03, 4), (0.005, 0.015, 1)):
    vol = tsdf.ScalableTSDFVolume(voxel, trunc, max_units=262144 if voxel < 0.01 else 65536, device=dev)
    tsdf.fuse_render_sweep(vol, p, vmd[:4], Kd[:4], cfg.width, cfg.height, depth_trunc=20.0, views_per_launch=vpl)   # warm-up
    vol.reset()
    torch.cuda.synchronize()
    lib.rs_timing_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    tsdf.fuse_render_sweep(vol, p, vmd, Kd, cfg.width, cfg.height, depth_trunc=20.0, views_per_launch=vpl)
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    spans = be.timing_collect()
    lib.rs_timing_enable(0)
    U = vol.n_units()
    touched_last = int(vol.counters[1])
    out[f"voxel_{voxel}_views_per_launch_{vpl}"] = {
        "views": n_views, "ms_per_view_device": round(e0.elapsed_time(e1) / n_views, 3),
        "ms_per_view_wall": round(wall * 1e3 / n_views, 3), "units_allocated": U,
        "units_touched_last_frame": touched_last,
        "tsdf_integrate_ms_per_view": round(spans["rs_tsdf_integrate"][0] / n_views, 4),
        "tsdf_algorithmic_GBps_last_frame_units": round(touched_last * 4096 * 40 / (spans["rs_tsdf_integrate"][0] / n_views * 1e-3) / 1e9, 1),
        "kernels_ms_per_view": {k: round(v[0] / n_views, 4) for k, v in sorted(spans.items(), key=lambda kv: -kv[1][0])[:8]},
    }
    del vol
    torch.cuda.empty_cache()
print(json.dumps(out, indent=1))
