"""Model-side glue that runs right after ``rasterization`` every training step.

Mirrors, on device tensors and without the per-step host work of the reference (camera matrices rebuilt on
the CPU and pixel grids uploaded each step, collab_splats/utils/camera_utils.py:54-71,228-238):

* ``depth_double_to_normal``  -- collab_splats/utils/camera_utils.py:176-279
* ``depth_normal_loss``       -- collab_splats/models/rade_gs_model.py:202-219 and :292-307
  (lambda 0.05, depth_ratio 0.6: collab_splats/configs/rade_gs_method.py:38-40)

SURVEY.md section 8(f) row f1 lists this stencil as the first "next" component (a fused kernel); this module is
its plain device-side form and the reference for that fusion.
"""

from __future__ import annotations

import torch
from torch import Tensor


def depth_double_to_normal(Ks_c: Tensor, width: int, height: int, depth1: Tensor, depth2: Tensor) -> Tensor:
    """Two z-depth maps [H,W] of one pinhole camera with centred principal point -> normals [2,H,W,3];
    central differences, border rows/columns are zero (camera_utils.py:253-279)."""
    dt, dev = depth1.dtype, depth1.device
    fx, fy = Ks_c[0, 0], Ks_c[1, 1]
    gx = (torch.arange(width, dtype=dt, device=dev) + 0.5)[None, :].expand(height, width)
    gy = (torch.arange(height, dtype=dt, device=dev) + 0.5)[:, None].expand(height, width)
    rays = torch.stack([gx / fx - width / (2 * fx), gy / fy - height / (2 * fy), torch.ones_like(gx)], dim=0)
    pts = torch.stack([depth1[None] * rays, depth2[None] * rays], dim=0)      # [2,3,H,W]
    out = torch.zeros_like(pts)
    d_row = pts[..., 2:, 1:-1] - pts[..., :-2, 1:-1]
    d_col = pts[..., 1:-1, 2:] - pts[..., 1:-1, :-2]
    out[..., 1:-1, 1:-1] = torch.nn.functional.normalize(torch.cross(d_row, d_col, dim=1), dim=1)
    return out.permute(0, 2, 3, 1)


def depth_normal_loss(Ks_c: Tensor, width: int, height: int, expected_depth: Tensor, median_depth: Tensor,
                      rendered_normals: Tensor, lam: float = 0.05, depth_ratio: float = 0.6):
    """expected/median depth [H,W], rendered normals [H,W,3] -> (loss, error maps [2,H,W])."""
    n_d = depth_double_to_normal(Ks_c, width, height, expected_depth, median_depth)
    err = 1.0 - (rendered_normals[None] * n_d).sum(dim=-1)
    loss = lam * ((1.0 - depth_ratio) * err[0].mean() + depth_ratio * err[1].mean())
    return loss, err


# ------------------------------------------------------------------------------------------------ fused kernel
class _FusedRadeLoss(torch.autograd.Function):
    """L1(RGB) + depth-normal consistency for ONE camera in one kernel (csrc/loss.cu): the forward pass also
    produces the gradients, so backward is a multiply by the upstream scalar (a no-op when it is 1)."""

    @staticmethod
    def forward(ctx, render, alphas, exp_depth, med_depth, normals, gt_u8, K, background, lam, depth_ratio,
                use_depth_normal, l1_weight):
        from . import backend as _be
        lib = _be.load()
        H, W, D = render.shape
        dev = render.device
        P = H * W
        sums = torch.zeros(4, device=dev, dtype=torch.float32)
        v_render = torch.empty_like(render)
        v_alphas = torch.empty_like(alphas)
        v_exp = torch.zeros_like(exp_depth)
        v_med = torch.zeros_like(med_depth)
        v_nrm = torch.empty_like(normals)
        w_l1 = l1_weight / (3.0 * P)
        w_exp = lam * (1.0 - depth_ratio) / P if use_depth_normal else 0.0
        w_med = lam * depth_ratio / P if use_depth_normal else 0.0
        fx, fy = K
        with torch.cuda.device(dev):
            _be.check(lib.rs_rade_loss_fwd_bwd(
                _be.ptr(render), _be.ptr(alphas), _be.ptr(exp_depth), _be.ptr(med_depth), _be.ptr(normals),
                _be.ptr(gt_u8), _be.ptr(background), fx, fy, W, H, D, w_l1, w_exp, w_med, int(use_depth_normal),
                _be.ptr(sums), _be.ptr(v_render), _be.ptr(v_alphas), _be.ptr(v_exp), _be.ptr(v_med), _be.ptr(v_nrm),
                _be.stream_ptr(dev)), "rs_rade_loss_fwd_bwd")
        ctx.save_for_backward(v_render, v_alphas, v_exp, v_med, v_nrm)
        ctx.set_materialize_grads(False)             # an unused output's gradient arrives as None (see backward)
        return sums[3], sums                         # loss, (l1, dn_expected, dn_median, loss)

    @staticmethod
    def backward(ctx, g_loss, g_terms):
        # the kernel wrote d(loss)/d(input) for an upstream gradient of 1; scale by the real one in place with one
        # launch that returns immediately when it IS 1 (tested on the device: no device->host read, and no 166 MB
        # pass in the common case).  The buffers are private to this node (a second backward through the same graph
        # would need retain_graph and is not supported).  Gradients w.r.t. the individual terms are not supported:
        # differentiate the total.
        from . import backend as _be
        if g_terms is not None:
            raise NotImplementedError("fused_rade_loss: the individual terms are reported, not differentiable -- "
                                      "differentiate the total loss")
        if g_loss is None:
            return (None,) * 12
        if getattr(ctx, "_consumed", False):
            raise RuntimeError("fused_rade_loss: a second backward through the same graph is not supported "
                               "(the gradient buffers are scaled in place)")
        ctx._consumed = True
        out = list(ctx.saved_tensors)
        _be.scale_unless_one(out, g_loss)
        return (*out, None, None, None, None, None, None, None)


def fused_rade_loss(render: Tensor, alphas: Tensor, expected_depth: Tensor, median_depth: Tensor,
                    rendered_normals: Tensor, gt_rgb_u8: Tensor, fx: float, fy: float,
                    background: Tensor = None, lam: float = 0.05, depth_ratio: float = 0.6,
                    use_depth_normal: bool = True, l1_weight: float = 1.0):
    """One camera: render [H,W,D>=3], alphas / depths [H,W], normals [H,W,3], gt uint8 [H,W,3] ->
    (loss, terms[4] = (L1, dn_expected, dn_median, loss)).  Same arithmetic as ``(clamp(rgb)-gt).abs().mean() +
    depth_normal_loss(...)`` above, fused with its own backward.  `fx, fy` are host floats (the principal point
    is the image centre, rade_gs_model.py:327-334) so no device->host read is needed.

    ``l1_weight`` scales the L1 term: the reference's rgb loss is ``(1 - ssim_lambda) * L1 + ssim_lambda * (1 - SSIM)``
    with ssim_lambda = 0.2 (nerfstudio's SplatfactoModel.get_loss_dict, reached through rade_gs_model.py:289); pass
    0.8 and add the SSIM term (host-framework code, not part of this path) to reproduce that objective."""
    assert render.dim() == 3 and render.shape[-1] >= 3 and gt_rgb_u8.dtype == torch.uint8
    c = lambda t: t.contiguous()
    return _FusedRadeLoss.apply(c(render), c(alphas), c(expected_depth), c(median_depth), c(rendered_normals),
                                c(gt_rgb_u8), (float(fx), float(fy)), None if background is None else c(background),
                                float(lam), float(depth_ratio), bool(use_depth_normal), float(l1_weight))
