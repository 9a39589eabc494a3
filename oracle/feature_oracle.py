"""CPU oracle for the feature decode + cosine loss of the rade-features model.
TEST INFRASTRUCTURE ONLY.

Restates, with plain torch on the CPU (fp32 or fp64, autograd for the backward), the reference's in-tree code:

* ``TwoLayerMLP``                 -- collab_splats/utils/features.py:408-456 (1x1 conv F->hidden, ReLU, one 1x1 conv
  per feature branch);
* ``RadegsFeaturesModel.decode_features`` -- collab_splats/models/rade_features_model.py:149-189 (rendered features
  [H,W,F] -> bilinear resize, align_corners=False, to the main branch's feature-map size -> decoder -> every other
  branch bilinearly resized to its own feature-map size);
* the features term of ``get_loss_dict``   -- rade_features_model.py:564-582 (per branch
  ``(1 - cosine_similarity(pred, gt, dim=0)).mean() * weight``, weight 1 for the main branch and
  ``features_regularization_lambda`` (0.1) otherwise, the sum scaled by ``features_loss_lambda`` (1e-3)).

Pinned: ``tests/golden/feature_decoder.npz`` holds outputs of the reference's own ``TwoLayerMLP`` class (its source
file loaded from /root/reference by tests/golden/make_feature_golden.py with the heavyweight imports it does not use
stubbed out) and of ``torch.nn.functional`` for the resize / cosine steps; tests/test_feature_decode.py checks this
oracle against them.  The two model methods cannot be imported (the model module needs nerfstudio and gsplat), so
they are restated line by line.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this module.
"""

from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor

FEATURES_LOSS_LAMBDA = 1e-3              # rade_features_model.py:39
FEATURES_REGULARIZATION_LAMBDA = 0.1     # rade_features_model.py:42


def mlp_forward(x_bchw: Tensor, w_hidden: Tensor, b_hidden: Tensor, branches: Dict[str, Tuple[Tensor, Tensor]]):
    """features.py:447-449.  ``w_hidden`` [Hd,F], branch weights [C,Hd] (the 1x1 kernels flattened)."""
    h = F.relu(F.conv2d(x_bchw, w_hidden[:, :, None, None], b_hidden))
    return {name: F.conv2d(h, w[:, :, None, None], b) for name, (w, b) in branches.items()}


def decode_features(features_hwf: Tensor, w_hidden: Tensor, b_hidden: Tensor,
                    branches: Dict[str, Tuple[Tensor, Tensor]], feature_dims: Dict[str, Tuple[int, int, int]],
                    main: str, resize_factor: float = 1.0) -> Dict[str, Tensor]:
    """rade_features_model.py:149-189: [H,W,F] -> {branch: [C_b, H_b, W_b]}."""
    x = features_hwf.permute(2, 0, 1)
    size = (int(feature_dims[main][1] * resize_factor), int(feature_dims[main][2] * resize_factor))
    x = F.interpolate(x.unsqueeze(0), size=size, mode="bilinear", align_corners=False)
    out = mlp_forward(x, w_hidden, b_hidden, branches)
    for name, dims in feature_dims.items():
        if name != main:
            out[name] = F.interpolate(out[name], size=tuple(dims[1:]), mode="bilinear", align_corners=False)
        out[name] = out[name].squeeze(0)
    return out


def features_loss(features_hwf: Tensor, w_hidden: Tensor, b_hidden: Tensor,
                  branches: Dict[str, Tuple[Tensor, Tensor]], feature_dims: Dict[str, Tuple[int, int, int]],
                  main: str, gt: Dict[str, Tensor], reg_lambda: float = FEATURES_REGULARIZATION_LAMBDA,
                  loss_lambda: float = FEATURES_LOSS_LAMBDA) -> Tensor:
    """rade_features_model.py:564-582."""
    decoded = decode_features(features_hwf, w_hidden, b_hidden, branches, feature_dims, main)
    total = torch.zeros((), dtype=features_hwf.dtype)
    for name, pred in decoded.items():
        weight = 1.0 if name == main else reg_lambda
        total = total + (1 - F.cosine_similarity(pred, gt[name], dim=0)).mean() * weight
    return total * loss_lambda
