"""Camera-sharded step on G GPUs (torchrun): gradients of ShGradExchange (p2p and allgather) + small-bucket
all-reduce against rank 0 rendering ALL cameras in one batch, and the time of the three gradient-exchange schemes.

    python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 scripts/multi_gpu_check.py
"""
import os, sys, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "collab-splats_b200")); sys.path.insert(0, str(ROOT))
import torch
import torch.distributed as dist
from datetime import timedelta

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device(f"cuda:{local}")
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev, timeout=timedelta(seconds=120))
from radegs_b200 import backend as be, scenes
from radegs_b200.multiview import ShGradExchange
from gsplat.rendering import rasterization
lib = be.load()

from radegs_b200.multiview import shard_views
N = int(os.environ.get("CHECK_N", "200000"))
VIEWS = int(os.environ.get("CHECK_VIEWS", str(world)))     # e.g. 3 views on 2 ranks: shards of unequal size (2 + 1)
mine = shard_views(VIEWS, rank, world)
cfg = scenes.SceneConfig("check", N, 640, 360, VIEWS, 3, 0, 77)
gs, vm, Ks = scenes.make_scene(cfg, n_views=VIEWS)
params = scenes.activate(gs, 3)
names = ["means", "quats", "scales", "opacities", "sh"]


def step(cams, exchange, small_allreduce):
    leaves = [p.detach().to(dev).requires_grad_(True) for p in params]
    out = rasterization(*leaves, vm[cams].to(dev), Ks[cams].to(dev), cfg.width, cfg.height, sh_degree=3, packed=False,
                        render_mode="RGB+ED", rasterize_mode="antialiased", return_depth_normal=True)
    # a fixed function of (pixel, channel) only, the same for every camera and for any position in the batch
    loss = sum((o * torch.sin(0.37 * torch.arange(o[0].numel(), device=dev).reshape(o.shape[1:]) + 0.1 * c0)[None]).sum()
               for o, c0 in zip(out[:5], range(5)))
    if exchange is None:
        loss.backward()
    else:
        exchange.begin_step()
        with exchange:
            loss.backward()
        leaves[4].grad = exchange.finish()
    if small_allreduce:
        flat = torch.cat([l.grad.reshape(-1) for l in leaves[:4]])
        dist.all_reduce(flat)
        o = 0
        for l in leaves[:4]:
            l.grad = flat[o:o + l.numel()].view(l.shape); o += l.numel()
    return [l.grad for l in leaves]


# the loss of camera c must not depend on which rank renders it: the cotangent above is a fixed function of the pixel
ref = None
if rank == 0:
    ref = [torch.zeros_like(p, device=dev) for p in params]
    for c in range(VIEWS):
        for r, g in zip(ref, step([c], None, False)):
            r += g
report = {}
for mode, engine in (("push", "dma"), ("push", "sm"), ("p2p", "dma"), ("allgather", "dma")):
    try:
        ex = ShGradExchange(N, len(mine), dev, mode=mode, push_engine=engine)
    except Exception as e:  # noqa: BLE001
        report[mode] = f"setup failed: {e}"
        continue
    mode = mode + ("/sm" if engine == "sm" else "")
    got = step(mine, ex, True)
    got2 = step(mine, ex, True)
    ex.check()
    if rank == 0:
        errs = {}
        for a, b, nm in zip(got, ref, names):
            errs[nm] = float((a - b).abs().max() / (b.abs().max() + 1e-30))
        errs["step2_vs_step1"] = max(float((a - b).abs().max()) for a, b in zip(got, got2))
        report[mode] = errs
    # every rank must hold bit-identical coefficient gradients
    h = got[4].double().sum().reshape(1)
    hs = [torch.zeros_like(h) for _ in range(world)]
    dist.all_gather(hs, h)
    if rank == 0:
        report[mode]["replicas_identical"] = bool(all(torch.equal(x, hs[0]) for x in hs))
    # timing of the exchange alone: local region is already filled; time signal+wait+gather / all-gather+gather
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    means_d = params[0].to(dev)
    reps = 20
    e0.record()
    for _ in range(reps):
        ex.begin_step()
        ex.published(means_d, 3, 16, be.stream_ptr(dev))     # the regions still hold the last step's rows
        out = ex.finish()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        report[mode]["exchange_ms"] = float(t.item())
    ex.check(); ex.close()
# the small-gradient all-reduce: two-shot peer kernel (rs_peer_allreduce) against NCCL, values and time
from radegs_b200.multiview import PeerAllReduce
n_small = 11 * int(os.environ.get("CHECK_AR_N", "1000000"))
try:
    ar = PeerAllReduce(n_small, dev)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    x = torch.randn(n_small, device=dev, generator=g)
    ref_sum = x.clone()
    dist.all_reduce(ref_sum)
    ar.flat.zero_(); ar.flat[:n_small].copy_(x)
    got = ar.all_reduce()[:n_small].clone()
    ar.check()
    err = float((got - ref_sum).abs().max() / ref_sum.abs().max())
    h = got.double().sum().reshape(1); hs = [torch.zeros_like(h) for _ in range(world)]; dist.all_gather(hs, h)
    times = {}
    for name, fn in (("peer", lambda: ar.all_reduce()), ("nccl", lambda: dist.all_reduce(ref_sum))):
        for _ in range(3): fn()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 20], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times[name] = float(t.item())
    ar.check(); ar.close()
    if rank == 0:
        report["small_allreduce"] = {"floats": n_small, "max_rel_diff_vs_nccl": err,
                                     "replicas_identical": bool(all(torch.equal(v, hs[0]) for v in hs)),
                                     "peer_ms": times["peer"], "nccl_ms": times["nccl"]}
except Exception as e:  # noqa: BLE001
    if rank == 0:
        report["small_allreduce"] = f"failed: {e}"
# the all-reduce it replaces
buf = torch.zeros(N, 16, 3, device=dev)
for _ in range(3): dist.all_reduce(buf)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): dist.all_reduce(buf)
e1.record(); torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / 20], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    report["allreduce_sh_coeffs_ms"] = float(t.item())
    report["n_gaussians"], report["world"], report["views"] = N, world, VIEWS
    print(json.dumps(report, indent=1))
dist.barrier(); dist.destroy_process_group()
