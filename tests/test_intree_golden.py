"""Golden vectors produced by the REFERENCE's own in-tree code (tests/golden/make_intree_golden.py):
``depth_double_to_normal`` (collab_splats/utils/camera_utils.py:176-279) and
``project_gaussians`` (collab_splats/utils/utils.py:13-40, row f4).

CPU: the oracle's restatements against them (this pins those two oracle functions).
GPU: the fused loss kernel (csrc/loss.cu) and the lookup kernel (csrc/stats.cu) against them."""

from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import rade_oracle as O

G = np.load(Path(__file__).resolve().parent / "golden" / "intree_glue.npz")


def _t(k):
    return torch.from_numpy(G[k])


def test_oracle_depth_double_to_normal_matches_reference():
    W, H = int(G["W"]), int(G["H"])
    got = O.depth_double_to_normal(_t("K")[0], W, H, _t("d1"), _t("d2"))
    ref = _t("normals")
    assert got.shape == ref.shape == (2, H, W, 3)
    # the reference goes focal -> fov -> focal through atan/tan in float64 and builds K^-1 in fp32: ~1e-6 relative
    torch.testing.assert_close(got, ref, atol=2e-5, rtol=0)
    assert float(ref[:, 0].abs().max()) == 0 and float(ref[:, :, -1].abs().max()) == 0      # zero border


def test_oracle_project_gaussians_matches_reference():
    W, H = int(G["W"]), int(G["H"])
    r = _t("radii").squeeze()
    valid = (r > 1.0).sum(dim=1) > 0
    xy = torch.round(_t("means2d")).squeeze().long()
    flat = torch.clamp(xy[:, 0], 0, W - 1) + torch.clamp(xy[:, 1], 0, H - 1) * W
    assert torch.equal(flat, _t("pg_proj_flattened"))
    assert torch.equal(valid, _t("pg_valid_mask"))
    assert torch.equal(valid.nonzero(as_tuple=False).squeeze(), _t("pg_gaussian_ids"))


@pytest.mark.gpu
def test_project_gaussians_kernel_matches_reference_golden(cuda_dev):
    from radegs_b200.meta_utils import project_gaussians
    meta = {"width": int(G["W"]), "height": int(G["H"]), "radii": _t("radii").to(cuda_dev),
            "means2d": _t("means2d").to(cuda_dev), "depths": _t("depths").to(cuda_dev)}
    got = project_gaussians(meta, to_cpu=True)
    for k in ("proj_flattened", "valid_mask", "gaussian_ids", "proj_depths"):
        assert torch.equal(got[k], _t("pg_" + k)), k


@pytest.mark.gpu
def test_fused_loss_error_maps_match_reference_golden(cuda_dev):
    """The depth-normal term of csrc/loss.cu on the golden depth maps: with rendered normals n the loss is
    lambda * ((1-r) * mean(1 - <n, N1>) + r * mean(1 - <n, N2>)), N1/N2 = the reference's depth normals."""
    from radegs_b200.losses import fused_rade_loss
    W, H = int(G["W"]), int(G["H"])
    K = _t("K")[0]
    g = torch.Generator().manual_seed(3)
    nrm = torch.nn.functional.normalize(torch.randn(H, W, 3, generator=g), dim=-1)
    N = _t("normals")
    err = 1.0 - (nrm[None] * N).sum(-1)
    ref_dn = 0.05 * (0.4 * err[0].mean() + 0.6 * err[1].mean())
    render = torch.rand(H, W, 3, generator=g)
    gt = (render * 255).round().to(torch.uint8)            # L1 term ~ quantisation only; subtract it below
    loss, terms = fused_rade_loss(render.to(cuda_dev), torch.ones(H, W, device=cuda_dev), _t("d1").to(cuda_dev),
                                  _t("d2").to(cuda_dev), nrm.to(cuda_dev), gt.to(cuda_dev), float(K[0, 0]),
                                  float(K[1, 1]), use_depth_normal=True)
    l1 = (render - gt.float() / 255.0).abs().mean()
    assert abs(loss.item() - l1.item() - ref_dn.item()) <= 2e-6, (loss.item(), l1.item(), ref_dn.item())
