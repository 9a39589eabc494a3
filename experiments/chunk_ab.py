"""A/B of the intersection pipelines (isect_tiles + offsets) on BASELINE configs 2, 5 and 4 (one view, and 4 views):
"radix" (presorted depth + radix sort of the tile bits + offset encode) vs "chunk" (chunked counting sort)."""
import math, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "collab-splats_b200")); sys.path.insert(0, str(ROOT))
import torch
from radegs_b200 import backend as be, scenes
import gsplat.cuda._wrapper as wr

lib = be.load()
dev = torch.device("cuda:0")
for cfg_id, views in ((2, 1), (5, 1), (4, 1), (4, 4)):
    cfg = scenes.BASELINE_CONFIGS[cfg_id]
    gs, vm, Ks = scenes.make_scene(cfg, n_views=views)
    means, quats, scales, _, _ = [t.to(dev) for t in scenes.activate(gs, 3)]
    W, H = cfg.width, cfg.height
    radii, m2, depths = wr.fully_fused_projection(means, None, quats, scales, vm.to(dev), Ks.to(dev), W, H)[:3]
    tw, th = math.ceil(W / 16), math.ceil(H / 16)
    ref = None
    for method in ("radix", "compact", "radix", "compact"):
        for _ in range(3): out = wr.isect_tiles_and_offsets(m2, radii, depths, 16, tw, th, method=method)
        lib.rs_timing_enable(1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): out = wr.isect_tiles_and_offsets(m2, radii, depths, 16, tw, th, method=method)
        e1.record(); torch.cuda.synchronize()
        s = be.timing_collect(); lib.rs_timing_enable(0)
        if ref is None: ref = out
        same = all(torch.equal(a, b) for a, b in zip(out, ref))
        ks = "  ".join(f"{k[3:]} {v[0] / 10:.4f}" for k, v in sorted(s.items(), key=lambda kv: -kv[1][0]))
        print(f"cfg {cfg_id} views {views} {method:6s}: wall {e0.elapsed_time(e1) / 10:.4f} ms  kernels {sum(v[0] for v in s.values()) / 10:.4f} ms  "
              f"M={out[1].numel()}  identical={same}\n    {ks}", flush=True)
    del means, quats, scales, radii, m2, depths, out, ref
    torch.cuda.empty_cache()
