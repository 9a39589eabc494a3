"""Feature decode + cosine loss of the rade-features model on the device.

Mirrors, with the same names and argument meaning:

* ``TwoLayerMLP``  -- collab_splats/utils/features.py:408-478 (``forward`` on a [1,F,H,W] map, ``per_gaussian_forward``
  on [N,F]); parameters are held the way ``nn.Conv2d(kernel_size=1)`` holds them so a reference ``state_dict`` loads;
* ``decode_features`` -- collab_splats/models/rade_features_model.py:149-189;
* ``features_loss``   -- the features term of ``get_loss_dict``, rade_features_model.py:564-582.

Three kernels (csrc/feature_decode.cu) replace the ~25 framework kernels + autograd graph of the reference; the
rendered features are read in place from ``rasterization()``'s colour output (``render[..., 3:3+F]``,
rade_features_model.py:360) and the gradient is written straight into that tensor's gradient image.
There is no CPU path: CPU tensors raise ``RuntimeError``.
"""

from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import torch
from torch import Tensor, nn

from radegs_b200 import backend as be

MAX_DIM = 128


def _check(t: Tensor, what: str) -> Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{what} must be a CUDA tensor: there is no CPU path")
    return t.detach().to(torch.float32).contiguous()


def _hidden_fwd(render: Tensor, ch0: int, F: int, size: Tuple[int, int], w1: Tensor, b1: Tensor):
    H, W, ld = render.shape
    Hm, Wm = size
    Hd = w1.shape[0]
    x = torch.empty(Hm * Wm, F, device=render.device, dtype=torch.float32)
    h = torch.empty(Hm * Wm, Hd, device=render.device, dtype=torch.float32)
    lib = be.load()
    with torch.cuda.device(render.device):
        be.check(lib.rs_feature_hidden_fwd(be.ptr(render), H, W, ld, ch0, F, Hm, Wm, be.ptr(w1), be.ptr(b1), Hd,
                                           be.ptr(x), be.ptr(h), be.stream_ptr(render.device)),
                 "rs_feature_hidden_fwd")
    return x, h


def _validate(render: Tensor, ch0: int, F: int, w1: Tensor):
    if render.dim() != 3:
        raise ValueError(f"rendered features must be [H,W,channels], got {tuple(render.shape)}")
    if ch0 < 0 or ch0 + F > render.shape[-1]:
        raise ValueError(f"feature channels {ch0}..{ch0 + F} do not fit in {render.shape[-1]} rendered channels")
    if w1.shape[1] != F:
        raise ValueError(f"decoder expects {w1.shape[1]} input channels, got {F}")
    if F > MAX_DIM or w1.shape[0] > MAX_DIM:
        raise NotImplementedError(f"feature / hidden width above {MAX_DIM} is not supported")


class _FeaturesLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, render, w1, b1, meta, gts, *branch_params):
        ch0, F, main_size, sizes, weights = meta
        dev = render.device
        r = _check(render, "render")
        w1c, b1c = _check(w1, "decoder weight"), _check(b1, "decoder bias")
        x, h = _hidden_fwd(r, ch0, F, main_size, w1c, b1c)
        Hm, Wm = main_size
        Hd = w1c.shape[0]
        lib = be.load()
        loss = torch.zeros(1, device=dev, dtype=torch.float32)
        v_h = torch.zeros_like(h)
        grads = []
        with torch.cuda.device(dev):
            for b in range(len(sizes)):
                w2, b2 = _check(branch_params[2 * b], "branch weight"), _check(branch_params[2 * b + 1], "branch bias")
                C, (Hb, Wb) = w2.shape[0], sizes[b]
                gt = _check(gts[b], "ground-truth features")
                if tuple(gt.shape) != (C, Hb, Wb):
                    raise ValueError(f"ground-truth features of branch {b} must be {(C, Hb, Wb)}, got {tuple(gt.shape)}")
                v_w2, v_b2 = torch.zeros_like(w2), torch.zeros_like(b2)
                psum = torch.zeros(Hb * Wb, 3, device=dev, dtype=torch.float32)
                be.check(lib.rs_feature_branch(be.ptr(h), Hm, Wm, Hd, be.ptr(w2), be.ptr(b2), C, be.ptr(gt), Hb, Wb,
                                               float(weights[b]) / (Hb * Wb), be.ptr(loss), be.ptr(psum), be.ptr(v_h),
                                               be.ptr(v_w2), be.ptr(v_b2), None, be.stream_ptr(dev)),
                         "rs_feature_branch")
                grads += [v_w2, v_b2]
        ctx.save_for_backward(x, h, v_h, w1c, *grads)
        ctx.meta = (ch0, F, main_size, tuple(render.shape), [w.shape for w in (w1, b1, *branch_params)])
        return loss[0]

    @staticmethod
    def backward(ctx, v_loss):
        x, h, v_h, w1c, *grads = ctx.saved_tensors
        ch0, F, (Hm, Wm), (H, W, ld), shapes = ctx.meta
        dev = x.device
        need_render = ctx.needs_input_grad[0]
        v_render = torch.zeros(H, W, ld, device=dev, dtype=torch.float32) if need_render else None
        v_w1 = torch.zeros_like(w1c)
        v_b1 = torch.zeros(w1c.shape[0], device=dev, dtype=torch.float32)
        lib = be.load()
        # every saved gradient was computed for d(loss) = 1; the incoming scalar is applied inside the kernel
        # (no extra pass over the gradient image) and to the small branch gradients below
        gs = v_loss.detach().to(torch.float32).reshape(1).contiguous()
        with torch.cuda.device(dev):
            be.check(lib.rs_feature_hidden_bwd(be.ptr(x), be.ptr(h), be.ptr(v_h), Hm, Wm, w1c.shape[0], F, be.ptr(w1c),
                                               be.ptr(v_w1), be.ptr(v_b1), be.ptr(v_render), H, W, ld, ch0,
                                               be.ptr(gs), be.stream_ptr(dev)), "rs_feature_hidden_bwd")
        out = [v_render, v_w1.view(shapes[0]), v_b1.view(shapes[1]), None, None]
        for g, shp in zip(grads, shapes[2:]):
            out.append((g * v_loss).view(shp))
        return tuple(out)


class TwoLayerMLP(nn.Module):
    """collab_splats/utils/features.py:408-478 with the same constructor, parameter names and shapes
    (``hidden_conv.weight`` [Hd,F,1,1], ``feature_branch_dict.<name>.weight`` [C,Hd,1,1])."""

    def __init__(self, input_dim: int, hidden_dim: int, features_dim_dict: Dict[str, Tuple[int, int, int]]):
        super().__init__()
        self.hidden_conv = nn.Conv2d(input_dim, hidden_dim, kernel_size=1)
        self.feature_branch_dict = nn.ModuleDict(
            {model: nn.Conv2d(hidden_dim, shape[0], kernel_size=1) for model, shape in features_dim_dict.items()})

    def _flat(self):
        w1 = self.hidden_conv.weight.view(self.hidden_conv.out_channels, -1)
        return w1, self.hidden_conv.bias, {k: (c.weight.view(c.out_channels, -1), c.bias)
                                           for k, c in self.feature_branch_dict.items()}

    def forward(self, x: Tensor) -> Dict[str, Tensor]:
        """[1,F,H,W] -> {branch: [1,C,H,W]} (inference; training goes through ``features_loss``, and a call that
        would need gradients raises)."""
        if x.dim() != 4 or x.shape[0] != 1:
            raise ValueError("TwoLayerMLP.forward takes a [1,F,H,W] map")
        _refuse_autograd("TwoLayerMLP.forward", x, self)
        _, F, H, W = x.shape
        hwf = x[0].permute(1, 2, 0).contiguous()
        out = decode_features(hwf, self, {k: (c.out_channels, H, W) for k, c in self.feature_branch_dict.items()},
                              main=next(iter(self.feature_branch_dict)))
        return {k: v.unsqueeze(0) for k, v in out.items()}

    @torch.no_grad()
    def per_gaussian_forward(self, x: Tensor) -> Dict[str, Tensor]:
        """[N,F] -> {branch: [N,C]} (features.py:452-478)."""
        N = x.shape[0]
        out = decode_features(x.view(N, 1, -1), self, {k: (c.out_channels, N, 1)
                                                       for k, c in self.feature_branch_dict.items()},
                              main=next(iter(self.feature_branch_dict)))
        return {k: v.view(v.shape[0], N).t() for k, v in out.items()}


def _refuse_autograd(what: str, render: Tensor, decoder: "TwoLayerMLP"):
    if torch.is_grad_enabled() and (render.requires_grad or any(p.requires_grad for p in decoder.parameters())):
        raise NotImplementedError(
            f"{what} is the inference path (the reference calls it under no_grad in get_outputs_for_camera, "
            "rade_features_model.py:509-511) and would silently drop gradients here; for training use "
            "radegs_b200.feature_decode.features_loss, or call it under torch.no_grad()")


def decode_features(render: Tensor, decoder: TwoLayerMLP, feature_dims: Dict[str, Sequence[int]], main: str,
                    resize_factor: float = 1.0, ch0: int = 0, n_features: Optional[int] = None) -> Dict[str, Tensor]:
    """rade_features_model.py:149-189: rendered features [H,W,F] (or the whole colour output with ``ch0`` /
    ``n_features`` selecting the feature columns) -> {branch: [C_b, H_b, W_b]}; the main branch at
    ``int(dims * resize_factor)``, every other branch at its own feature-map size.  Inference only: with gradients
    enabled and differentiable inputs it raises instead of returning tensors that are cut off from the graph."""
    _refuse_autograd("decode_features", render, decoder)
    with torch.no_grad():
        return _decode_features(render, decoder, feature_dims, main, resize_factor, ch0, n_features)


def _decode_features(render, decoder, feature_dims, main, resize_factor, ch0, n_features):
    w1, b1, branches = decoder._flat()
    F = n_features if n_features is not None else render.shape[-1] - ch0
    _validate(render, ch0, F, w1)
    r = _check(render, "render")
    main_size = (int(feature_dims[main][1] * resize_factor), int(feature_dims[main][2] * resize_factor))
    w1c, b1c = _check(w1, "decoder weight"), _check(b1, "decoder bias")
    _, h = _hidden_fwd(r, ch0, F, main_size, w1c, b1c)
    lib = be.load()
    out = {}
    with torch.cuda.device(r.device):
        for name, (w2, b2) in branches.items():
            Hb, Wb = main_size if name == main else (int(feature_dims[name][1]), int(feature_dims[name][2]))
            w2c, b2c = _check(w2, "branch weight"), _check(b2, "branch bias")
            dec = torch.empty(w2c.shape[0], Hb, Wb, device=r.device, dtype=torch.float32)
            be.check(lib.rs_feature_branch(be.ptr(h), main_size[0], main_size[1], w1c.shape[0], be.ptr(w2c), be.ptr(b2c),
                                           w2c.shape[0], None, Hb, Wb, 0.0, None, None, None, None, None, be.ptr(dec),
                                           be.stream_ptr(r.device)), "rs_feature_branch")
            out[name] = dec
    return out


def features_loss(render: Tensor, decoder: TwoLayerMLP, feature_dims: Dict[str, Sequence[int]], main: str,
                  gt: Dict[str, Tensor], features_regularization_lambda: float = 0.1,
                  features_loss_lambda: float = 1e-3, ch0: int = 0, n_features: Optional[int] = None) -> Tensor:
    """rade_features_model.py:564-582: ``features_loss_lambda * sum_b w_b * mean(1 - cos(pred_b, gt_b))`` with
    gradients to ``render`` (all its columns; zero outside the feature columns) and to the decoder parameters."""
    w1, b1, branches = decoder._flat()
    F = n_features if n_features is not None else render.shape[-1] - ch0
    _validate(render, ch0, F, w1)
    if not render.is_cuda:
        raise RuntimeError("render must be a CUDA tensor: there is no CPU path")
    names = list(branches)
    main_size = (int(feature_dims[main][1]), int(feature_dims[main][2]))
    sizes = [main_size if n == main else (int(feature_dims[n][1]), int(feature_dims[n][2])) for n in names]
    weights = [(1.0 if n == main else features_regularization_lambda) * features_loss_lambda for n in names]
    flat = []
    for n in names:
        flat += list(branches[n])
    return _FeaturesLoss.apply(render, w1, b1, (ch0, F, main_size, sizes, weights), [gt[n] for n in names], *flat)
