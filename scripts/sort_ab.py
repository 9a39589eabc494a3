"""A/B of the radix scatter kernels on the real keys of BASELINE config 2 (run on a B200): the tile-bit sort of the
depth-ordered intersections (rs_sort_pairs, bits [32, 32 + tile bits)) and the depth argsort (rs_argsort_u32)."""
import math, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "collab-splats_b200")); sys.path.insert(0, str(ROOT))
import torch
from radegs_b200 import backend as be, scenes
from gsplat.cuda._wrapper import fully_fused_projection, isect_tiles

lib = be.load()
dev = torch.device("cuda:0")
cfg = scenes.BASELINE_CONFIGS[int(sys.argv[1]) if len(sys.argv) > 1 else 2]
gs, vm, Ks = scenes.make_scene(cfg, n_views=1)
means, quats, scales, _, _ = [t.to(dev) for t in scenes.activate(gs, 3)]
radii, m2, depths = fully_fused_projection(means, None, quats, scales, vm.to(dev), Ks.to(dev), cfg.width, cfg.height)[:3]
tw, th = math.ceil(cfg.width / 16), math.ceil(cfg.height / 16)
_, ids, flat = isect_tiles(m2, radii, depths, 16, tw, th, sort=False)
# the pipeline's input order: depth-sorted, so that only the tile bits remain to be sorted
order = torch.sort(ids & 0xFFFFFFFF, stable=True).indices
ids, flat = ids[order].contiguous(), flat[order].contiguous()
M = ids.numel()
b0, b1 = 32, 32 + lib.rs_tile_bits(tw, th)
tb = lib.rs_sort_pairs_temp_bytes(M, b0, b1)
temp = torch.empty(tb, device=dev, dtype=torch.uint8)
srt = torch.sort(ids >> 32, stable=True)
ref = (ids[srt.indices], flat[srt.indices])
dk = depths.reshape(-1).contiguous().view(torch.int32)
n = dk.numel()
tb2 = lib.rs_sort_pairs_temp_bytes(n, 0, 32)
temp2 = torch.empty(tb2, device=dev, dtype=torch.uint8)
dref = torch.sort(dk.to(torch.int64) & 0xFFFFFFFF, stable=True).indices.to(torch.int32)


def timed(fn, reps=15):
    ts = []
    for _ in range(reps):
        args = fn(None)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(args); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0], out


def pairs(args):
    if args is None:
        ka, va = ids.clone(), flat.clone()
        return ka, va, torch.empty_like(ka), torch.empty_like(va)
    ka, va, kb, vb = args
    w = be.check(lib.rs_sort_pairs(be.ptr(ka), be.ptr(va), be.ptr(kb), be.ptr(vb), M, b0, b1, be.ptr(temp), tb,
                                   be.stream_ptr(dev)), "sort")
    return (kb, vb) if w == 0 else (ka, va)


def argsort(args):
    if args is None:
        ka = dk.clone()
        return ka, torch.empty_like(ka), torch.empty_like(ka), torch.empty_like(ka)
    ka, va, kb, vb = args
    w = be.check(lib.rs_argsort_u32(be.ptr(ka), be.ptr(va), be.ptr(kb), be.ptr(vb), n, 0, 32, be.ptr(temp2), tb2,
                                    be.stream_ptr(dev)), "argsort")
    return vb if w == 0 else va


for rep in range(2):
    med, mn, out = timed(pairs)
    assert torch.equal(out[0], ref[0]) and torch.equal(out[1], ref[1]), "pairs differ"
    med2, mn2, out2 = timed(argsort)
    assert torch.equal(out2, dref), "argsort differs"
    print(f"sort_pairs M={M} bits[{b0},{b1}) median {med:.4f} min {mn:.4f} ms | "
          f"argsort n={n} median {med2:.4f} min {mn2:.4f} ms")
