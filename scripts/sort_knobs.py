"""Sweep of the radix-sort knobs (items per thread, look-back window) on the two sorts of the presorted isect path."""
import math, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "collab-splats_b200")); sys.path.insert(0, str(ROOT))
import torch
from radegs_b200 import backend as be, scenes
import gsplat.cuda._wrapper as wr
lib = be.load(); dev = torch.device("cuda:0")
cfg = scenes.BASELINE_CONFIGS[2]
gs, vm, Ks = scenes.make_scene(cfg, n_views=1)
means, quats, scales, _, _ = [t.to(dev) for t in scenes.activate(gs, 3)]
W, H = cfg.width, cfg.height
radii, m2, depths = wr.fully_fused_projection(means, None, quats, scales, vm.to(dev), Ks.to(dev), W, H)[:3]
tw, th = math.ceil(W / 16), math.ceil(H / 16)
for items in (8, 16):
    for window in (4, 8, 16):
        lib.rs_sort_set_items(items); lib.rs_sort_set_window(window)
        for _ in range(3): out = wr.isect_tiles(m2, radii, depths, 16, tw, th)
        lib.rs_timing_enable(1)
        for _ in range(20): out = wr.isect_tiles(m2, radii, depths, 16, tw, th)
        torch.cuda.synchronize()
        s = be.timing_collect(); lib.rs_timing_enable(0)
        print(f"items {items:2d} window {window:2d}: argsort_u32 {s['rs_argsort_u32'][0] / 20:.4f}  sort_pairs {s['rs_sort_pairs'][0] / 20:.4f}")
lib.rs_sort_set_items(8); lib.rs_sort_set_window(4)
