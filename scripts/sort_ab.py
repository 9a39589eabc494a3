"""A/B of the radix sort tuning knob on the real keys of BASELINE config 2 (run on a B200)."""
import math, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "collab-splats_b200")); sys.path.insert(0, str(ROOT))
import torch
from radegs_b200 import backend as be, scenes
from gsplat.cuda._wrapper import fully_fused_projection, isect_tiles

lib = be.load()
dev = torch.device("cuda:0")
cfg = scenes.BASELINE_CONFIGS[2]
gs, vm, Ks = scenes.make_scene(cfg, n_views=1)
means, quats, scales, _, _ = [t.to(dev) for t in scenes.activate(gs, 3)]
radii, m2, depths = fully_fused_projection(means, None, quats, scales, vm.to(dev), Ks.to(dev), cfg.width, cfg.height)[:3]
tw, th = math.ceil(cfg.width / 16), math.ceil(cfg.height / 16)
_, ids, flat = isect_tiles(m2, radii, depths, 16, tw, th, sort=False)
M = ids.numel()
end_bit = 32 + lib.rs_tile_bits(tw, th) + 1
tb = lib.rs_sort_pairs_temp_bytes(M, 0, end_bit)
temp = torch.empty(tb, device=dev, dtype=torch.uint8)
ref = None
for items in (8, 16, 8, 16):
    lib.rs_sort_set_items(items)
    ts = []
    for rep in range(12):
        ka, va = ids.clone(), flat.clone()
        kb, vb = torch.empty_like(ka), torch.empty_like(va)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        where = be.check(lib.rs_sort_pairs(be.ptr(ka), be.ptr(va), be.ptr(kb), be.ptr(vb), M, 0, end_bit, be.ptr(temp), tb,
                                           be.stream_ptr(dev)), "sort")
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    out = (kb, vb) if where == 0 else (ka, va)
    if ref is None:
        ref = (out[0].clone(), out[1].clone())
        srt = torch.sort(ids, stable=True)
        assert torch.equal(ref[0], srt.values) and torch.equal(ref[1], flat[srt.indices])
    assert torch.equal(out[0], ref[0]) and torch.equal(out[1], ref[1])
    ts.sort()
    gbs = M * (8 + 24 * math.ceil(end_bit / 8)) / (ts[len(ts) // 2] * 1e-3) / 1e9
    print(f"items={items:2d}  M={M}  median {ts[len(ts)//2]:.4f} ms  min {ts[0]:.4f} ms  algorithmic {gbs:.0f} GB/s")
