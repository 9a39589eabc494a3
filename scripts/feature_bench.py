"""Feature decode + cosine loss at BASELINE config 3 sizes: fused kernels vs the reference's
framework formulation (F.interpolate + 1x1 convs + cosine_similarity + autograd) run on the same GPU."""
import json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "collab-splats_b200")); sys.path.insert(0, str(ROOT))
import torch
import torch.nn.functional as F
from radegs_b200 import backend as be, feature_decode as fd

dev = torch.device("cuda:0")
lib = be.load()
H, W, NF = 540, 960, 64
dims = {"clip": (768, 38, 68), "dino": (384, 38, 68)}
torch.manual_seed(0)
dec = fd.TwoLayerMLP(NF, 64, dims).to(dev)
render = torch.randn(H, W, 3 + NF + 1, device=dev).requires_grad_(True)
gt = {k: torch.randn(*v, device=dev) for k, v in dims.items()}


def ours():
    render.grad = None
    loss = fd.features_loss(render, dec, dims, "clip", gt, ch0=3, n_features=NF)
    loss.backward()
    return loss


def framework():
    render.grad = None
    feats = render[..., 3:3 + NF].permute(2, 0, 1)
    x = F.interpolate(feats.unsqueeze(0), size=dims["clip"][1:], mode="bilinear", align_corners=False)
    h = F.relu(dec.hidden_conv(x))
    loss = 0.0
    for k, conv in dec.feature_branch_dict.items():
        y = conv(h)
        if k != "clip":
            y = F.interpolate(y, size=dims[k][1:], mode="bilinear", align_corners=False)
        loss = loss + (1 - F.cosine_similarity(y.squeeze(0), gt[k], dim=0)).mean() * (1.0 if k == "clip" else 0.1)
    loss = loss * 1e-3
    loss.backward()
    return loss


def timed(fn, reps=20, warm=3):
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


a, b = ours(), framework()
out = {"loss_ours": float(a), "loss_framework": float(b), "ours_ms": round(timed(ours), 4),
       "framework_ms": round(timed(framework), 4)}
lib.rs_timing_enable(1)
for _ in range(5):
    ours()
torch.cuda.synchronize()
out["kernels_ms"] = {k: round(v[0] / 5, 4) for k, v in be.timing_collect().items()}
lib.rs_timing_enable(0)
print(json.dumps(out, indent=1))
