"""Raw transfer rates for the gradient exchange at G GPUs (torchrun): every rank sends 16 MB to each peer at once
(copy engines, 1/4/7 streams) vs NCCL all-gather of the same payload."""
import os, sys, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "collab-splats_b200")); sys.path.insert(0, str(ROOT))
import torch, ctypes as ct
import torch.distributed as dist
from datetime import timedelta
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device(f"cuda:{local}"); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev, timeout=timedelta(seconds=120))
from radegs_b200 import backend as be
from radegs_b200.multiview import ShGradExchange
lib = be.load()
N = 1_000_000
ex = ShGradExchange(N, 1, dev, mode="push")
nbytes = ex.region_bytes
rep = {}
def timed(fn, reps=10):
    for _ in range(2): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
peers = [g for g in range(world) if g != rank]
for ns in (1, 4, 7):
    streams = [torch.cuda.Stream(dev) for _ in range(min(ns, len(peers)))]
    def push():
        main = torch.cuda.current_stream(dev)
        ev = torch.cuda.Event(); ev.record(main)
        for s in streams: s.wait_event(ev)
        for i in range(len(peers)):
            g = peers[(i + rank) % len(peers)]
            with torch.cuda.stream(streams[i % len(streams)]):
                be.check(lib.rs_peer_copy(ct.c_void_p(ex.region_base[g] + (2 * rank) * ex.region_stride),
                                          ct.c_void_p(ex.region_base[rank] + (2 * rank) * ex.region_stride), nbytes,
                                          be.stream_ptr(dev)), "copy")
        for s in streams:
            e = torch.cuda.Event(); e.record(s); main.wait_event(e)
    ms = timed(push)
    rep[f"dma_{ns}_streams_ms"] = ms
    rep[f"dma_{ns}_streams_ingress_GBps"] = (world - 1) * nbytes / ms / 1e6
local_buf = torch.zeros(nbytes, device=dev, dtype=torch.uint8)
gathered = torch.zeros(world * nbytes, device=dev, dtype=torch.uint8)
ms = timed(lambda: dist.all_gather_into_tensor(gathered, local_buf))
rep["nccl_all_gather_ms"] = ms
rep["nccl_all_gather_ingress_GBps"] = (world - 1) * nbytes / ms / 1e6
small = torch.zeros(N * 11, device=dev)
rep["nccl_all_reduce_44MB_ms"] = timed(lambda: dist.all_reduce(small))
big = torch.zeros(N * 48, device=dev)
rep["nccl_all_reduce_192MB_ms"] = timed(lambda: dist.all_reduce(big))
# gather kernel alone, pulling over NVLink (p2p) vs local inboxes (push)
means = torch.rand(N, 3, device=dev)
out = torch.empty(N, 16, 3, device=dev)
cams = (ct.c_int * world)(*([1] * world))
for name, bases in (("pull", [ex.region_base[g] + 2 * g * ex.region_stride for g in range(world)]),
                    ("local", [ex.region_base[rank] + 2 * g * ex.region_stride for g in range(world)])):
    regions = (ct.c_void_p * world)(*bases)
    rep[f"gather_kernel_{name}_ms"] = timed(lambda: be.check(lib.rs_sh_coeffs_gather(
        3, 16, N, be.ptr(means), regions, cams, world, be.ptr(out), be.stream_ptr(dev)), "gather"))
if rank == 0:
    rep["world"], rep["payload_MB_per_rank"] = world, nbytes / 1e6
    print(json.dumps(rep, indent=1))
ex.close()
dist.barrier(); dist.destroy_process_group()
