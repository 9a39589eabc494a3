"""Time rs_sh_coeffs_gather on one GPU with 8 local source regions (1 M Gaussians, degree 3): the per-rank cost of
rebuilding the SH coefficient gradient over 8 cameras in the N=8 camera-sharded step."""
import ctypes as ct, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "collab-splats_b200")); sys.path.insert(0, str(ROOT))
import torch
from radegs_b200 import backend as be
lib = be.load(); dev = torch.device("cuda:0")
N, K, S = 1_000_000, 16, int(sys.argv[1]) if len(sys.argv) > 1 else 8
means = torch.rand(N, 3, device=dev) * 2 - 1
regions = []
for s in range(S):
    r = torch.zeros(lib.rs_sh_region_bytes(1, N), device=dev, dtype=torch.uint8)
    r[:16].view(torch.float32)[:4] = torch.tensor([3.0 * (s + 1), 0.1 * s, -2.0, 1.0], device=dev)
    body = r[1024:1024 + N * 12].view(torch.float32).view(N, 3)
    body[:] = torch.randn(N, 3, device=dev) * (torch.rand(N, 1, device=dev) > 0.2)
    regions.append(r)
v = torch.empty(N, K, 3, device=dev)
ptrs = (ct.c_void_p * S)(*[r.data_ptr() for r in regions]); cams = (ct.c_int * S)(*[1] * S)
st = be.stream_ptr(dev)
for _ in range(3): be.check(lib.rs_sh_coeffs_gather(3, K, N, be.ptr(means), ptrs, cams, S, be.ptr(v), st), "g")
ts = []
for _ in range(10):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); be.check(lib.rs_sh_coeffs_gather(3, K, N, be.ptr(means), ptrs, cams, S, be.ptr(v), st), "g"); e1.record()
    torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ts.sort()
print(f"sources {S}: gather median {ts[5]:.4f} ms, {(N * (12 * S + 12 + 192)) / (ts[5] * 1e-3) / 1e9:.0f} GB/s algorithmic; checksum {float(v.sum()):.3f}")
