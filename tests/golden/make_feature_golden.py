"""Generates tests/golden/feature_decoder.npz from the REFERENCE's own decoder class.

``collab_splats/utils/features.py`` (TwoLayerMLP, :408-478) is loaded from /root/reference as a standalone module;
the packages its top-level imports name but TwoLayerMLP never touches (maskclip_onnx, huggingface_hub, torchvision)
are replaced by empty stubs when missing.  The fixture holds a seeded input, the class's own randomly initialised
weights, its forward outputs, its ``per_gaussian_forward`` outputs, and the decode/loss pipeline of
rade_features_model.py:149-189,564-582 evaluated with the reference decoder instance and torch.nn.functional
(forward + autograd gradients).  Only runs in the build container (the GPU box has no /root/reference).
Re-generate with:  python tests/golden/make_feature_golden.py
"""
import importlib.util
import sys
import types
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

REF = Path("/root/reference/collab_splats/utils/features.py")
OUT = Path(__file__).resolve().parent / "feature_decoder.npz"

for name in ("maskclip_onnx", "huggingface_hub", "torchvision", "torchvision.transforms"):
    try:
        __import__(name)
    except Exception:
        m = types.ModuleType(name)
        m.hf_hub_download = lambda *a, **k: None
        sys.modules[name] = m
if not hasattr(sys.modules["torchvision"], "transforms"):
    sys.modules["torchvision"].transforms = sys.modules["torchvision.transforms"]

spec = importlib.util.spec_from_file_location("ref_features", REF)
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

torch.manual_seed(20261018)
Fin, Hd = 13, 64                                           # reference defaults (rade_features_model.py:61,64)
dims = {"clip": (96, 9, 12), "dino": (40, 7, 10)}           # (C, H, W) per branch; "clip" is the main branch
main = "clip"
H, W = 45, 60
dec = ref.TwoLayerMLP(input_dim=Fin, hidden_dim=Hd, features_dim_dict=dims)
feats = (torch.randn(H, W, Fin) * 0.5).requires_grad_(True)
gt = {k: torch.randn(*v) for k, v in dims.items()}

# --- decode_features (rade_features_model.py:149-189) with the reference decoder
x = feats.permute(2, 0, 1)
x = F.interpolate(x.unsqueeze(0), size=dims[main][1:], mode="bilinear", align_corners=False)
dd = dec(x)
for k, v in dims.items():
    if k != main:
        dd[k] = F.interpolate(dd[k], size=v[1:], mode="bilinear", align_corners=False)
    dd[k] = dd[k].squeeze(0)
# --- get_loss_dict features term (rade_features_model.py:564-582)
loss = torch.tensor(0.0)
for k, pred in dd.items():
    weight = 1.0 if k == main else 0.1
    loss = loss + (1 - F.cosine_similarity(pred, gt[k], dim=0)).mean() * weight
loss = loss * 1e-3
loss.backward()

per_gauss_in = torch.randn(50, Fin)
pg = dec.per_gaussian_forward(per_gauss_in)

d = {"features": feats.detach(), "v_features": feats.grad, "loss": loss.detach(),
     "w_hidden": dec.hidden_conv.weight.detach().view(Hd, Fin), "b_hidden": dec.hidden_conv.bias.detach(),
     "v_w_hidden": dec.hidden_conv.weight.grad.view(Hd, Fin), "v_b_hidden": dec.hidden_conv.bias.grad,
     "per_gauss_in": per_gauss_in}
for k in dims:
    conv = dec.feature_branch_dict[k]
    d[f"w_{k}"] = conv.weight.detach().view(dims[k][0], Hd)
    d[f"b_{k}"] = conv.bias.detach()
    d[f"v_w_{k}"] = conv.weight.grad.view(dims[k][0], Hd)
    d[f"v_b_{k}"] = conv.bias.grad
    d[f"gt_{k}"] = gt[k]
    d[f"decoded_{k}"] = dd[k].detach()
    d[f"per_gauss_{k}"] = pg[k]
np.savez_compressed(OUT, **{k: v.numpy() for k, v in d.items()})
print("wrote", OUT, OUT.stat().st_size, "bytes; loss", float(loss))
