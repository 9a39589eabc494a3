"""``RadegsModel.get_outputs`` after the rasterization call (collab_splats/models/rade_gs_model.py:200-271, SURVEY.md
row a14), fused: two launches forward (four global maxima, one pass over the pixels), one backward -- instead of ~25
elementwise / reduction kernels and the host-side camera rebuild of ``depth_double_to_normal``
(collab_splats/utils/camera_utils.py:176-279).  Same keys, shapes and values as the reference's dict for one camera.
"""

from __future__ import annotations

from typing import Dict, Optional

import torch
from torch import Tensor


class _RadeOutputs(torch.autograd.Function):
    @staticmethod
    def forward(ctx, render, alpha, exp_d, med_d, normals, background, fx, fy, depth_channel, use_dn):
        from . import backend as _be
        lib = _be.load()
        H, W, D = render.shape
        dev = render.device
        f32 = dict(device=dev, dtype=torch.float32)
        maxima = torch.zeros(4, device=dev, dtype=torch.int32)
        rgb = torch.empty(H, W, 3, **f32)
        depth = torch.empty(H, W, 1, **f32)
        median = torch.empty(H, W, 1, **f32)
        depth_im = torch.empty(H, W, 1, **f32) if depth_channel >= 0 else None
        nrm_out = torch.empty(H, W, 3, **f32)
        err = torch.empty(2, H, W, 1, **f32)
        with torch.cuda.device(dev):
            _be.check(lib.rs_rade_outputs_fwd(
                _be.ptr(render), _be.ptr(alpha), _be.ptr(exp_d), _be.ptr(med_d), _be.ptr(normals), _be.ptr(background),
                fx, fy, W, H, D, depth_channel, int(use_dn), _be.ptr(maxima), _be.ptr(rgb), _be.ptr(depth),
                _be.ptr(median), _be.ptr(depth_im), _be.ptr(nrm_out), _be.ptr(err), _be.stream_ptr(dev)),
                "rs_rade_outputs_fwd")
        ctx.save_for_backward(render, alpha, exp_d, med_d, normals, background)
        ctx.cfg = (fx, fy, depth_channel, use_dn)
        ctx.set_materialize_grads(False)
        if depth_im is None:
            depth_im = torch.empty(0, **f32)
            ctx.mark_non_differentiable(depth_im)
        return rgb, depth, median, depth_im, nrm_out, err[0], err[1]

    @staticmethod
    def backward(ctx, v_rgb, v_depth, v_median, v_depth_im, v_nrm, v_e0, v_e1):
        from . import backend as _be
        lib = _be.load()
        render, alpha, exp_d, med_d, normals, background = ctx.saved_tensors
        fx, fy, depth_channel, use_dn = ctx.cfg
        H, W, D = render.shape
        dev = render.device
        c = lambda t: None if t is None else t.contiguous()
        v_err = None
        if v_e0 is not None or v_e1 is not None:
            z = torch.zeros(H, W, 1, device=dev)
            v_err = torch.stack([z if v_e0 is None else v_e0, z if v_e1 is None else v_e1]).contiguous()
        g_render = torch.empty_like(render)
        g_alpha = torch.empty_like(alpha)
        g_exp, g_med = torch.zeros_like(exp_d), torch.zeros_like(med_d)
        g_nrm = torch.empty_like(normals)
        with torch.cuda.device(dev):
            _be.check(lib.rs_rade_outputs_bwd(
                _be.ptr(render), _be.ptr(alpha), _be.ptr(exp_d), _be.ptr(med_d), _be.ptr(normals), _be.ptr(background),
                fx, fy, W, H, D, depth_channel, int(use_dn), _be.ptr(c(v_rgb)), _be.ptr(c(v_depth)),
                _be.ptr(c(v_median)), _be.ptr(c(v_depth_im)) if depth_channel >= 0 else None, _be.ptr(c(v_nrm)),
                _be.ptr(v_err), None, _be.ptr(g_render), _be.ptr(g_alpha), _be.ptr(g_exp), _be.ptr(g_med),
                _be.ptr(g_nrm), _be.stream_ptr(dev)), "rs_rade_outputs_bwd")
        return g_render, g_alpha, g_exp, g_med, g_nrm, None, None, None, None, None


def rade_get_outputs(render: Tensor, alpha: Tensor, expected_depths: Tensor, median_depths: Tensor,
                     expected_normals: Tensor, background: Tensor, fx: float, fy: float,
                     render_mode: str = "RGB+ED", use_depth_normal: bool = True) -> Dict[str, Optional[Tensor]]:
    """One camera, as the reference calls its rasterizer (C = 1): render [1,H,W,D] or [H,W,D], alpha / depths
    [1,H,W,1] or [H,W,1], expected_normals [1,H,W,3] or [H,W,3], background [3] -> the reference's output dict
    (rgb [H,W,3], depth / median_depth / depth_im / accumulation [H,W,1], normals [H,W,3],
    depth_normal_error_map / middepth_normal_error_map [H,W,1], background).  ``use_depth_normal=False`` is the
    reference's branch before ``regularization_from_iter``: the error maps are zero.  `fx, fy` are host floats (the
    principal point is the image centre, rade_gs_model.py:327-334)."""
    if not render.is_cuda:
        raise RuntimeError("rade_get_outputs: tensors must live on a CUDA device; there is no CPU path")
    sq = lambda t: t[0] if t.dim() == 4 else t
    render, alpha, exp_d, med_d, nrm = (sq(t).contiguous().float() for t in
                                        (render, alpha, expected_depths, median_depths, expected_normals))
    H, W, D = render.shape
    assert D >= 3 and alpha.shape == (H, W, 1) and exp_d.shape == (H, W, 1) and med_d.shape == (H, W, 1)
    assert nrm.shape == (H, W, 3) and background.shape == (3,)
    ed = render_mode == "RGB+ED"
    assert not ed or D >= 4
    rgb, depth, median, depth_im, normals, e0, e1 = _RadeOutputs.apply(
        render, alpha, exp_d, med_d, nrm, background.contiguous().float(), float(fx), float(fy), 3 if ed else -1,
        bool(use_depth_normal))
    return {"rgb": rgb, "depth": depth, "median_depth": median, "depth_im": depth_im if ed else None,
            "accumulation": alpha, "normals": normals, "depth_normal_error_map": e0,
            "middepth_normal_error_map": e1, "background": background}
