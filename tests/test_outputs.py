"""get_outputs epilogue: golden vectors produced by EXECUTING the reference's own lines
(collab_splats/models/rade_gs_model.py:200-271, tests/golden/make_outputs_golden.py).

CPU: the oracle's restatement against them (this pins ``oracle.rade_oracle.get_outputs_glue``).
GPU: the fused kernels (csrc/outputs.cu, radegs_b200.outputs.rade_get_outputs) against them, and their backward
     against the oracle's autograd."""

from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import rade_oracle as O

G = np.load(Path(__file__).resolve().parent / "golden" / "rade_outputs.npz")
KEYS = ("rgb", "depth", "median_depth", "depth_im", "accumulation", "normals", "depth_normal_error_map",
        "middepth_normal_error_map")


def _t(k):
    return torch.from_numpy(G[k])


def _gold(mode, key):
    name = f"{mode}__{key}".replace("+", "p")
    return torch.from_numpy(G[name]) if name in G.files else None


def _inputs(mode):
    render = _t("render")[0] if mode == "RGB+ED" else _t("render")[0, ..., :3]
    return render, _t("alpha")[0], _t("exp_d")[0], _t("med_d")[0], _t("normals")[0], _t("background")


@pytest.mark.parametrize("mode", ["RGB+ED", "RGB"])
def test_oracle_get_outputs_matches_reference(mode):
    W, H = int(G["W"]), int(G["H"])
    render, alpha, exp_d, med_d, nrm, bg = _inputs(mode)
    got = O.get_outputs_glue(_t("K")[0], W, H, render, alpha, exp_d, med_d, nrm, bg, render_mode=mode)
    n_masked = int((alpha <= 0).sum())
    assert n_masked > 100, "the golden must exercise the masked fills"
    for k in KEYS:
        ref = _gold(mode, k)
        if ref is None:
            assert got[k] is None, k
            continue
        assert got[k].shape == ref.shape, (k, got[k].shape, ref.shape)
        # exact for the elementwise outputs; the error maps go through the reference's fov round trip (~1e-6)
        tol = 2e-5 if "error_map" in k else 0.0
        torch.testing.assert_close(got[k], ref, atol=tol, rtol=0)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["RGB+ED", "RGB"])
def test_outputs_kernel_matches_reference_golden(cuda_dev, mode):
    from radegs_b200.outputs import rade_get_outputs
    render, alpha, exp_d, med_d, nrm, bg = _inputs(mode)
    K = _t("K")[0]
    got = rade_get_outputs(render.to(cuda_dev)[None], alpha.to(cuda_dev)[None], exp_d.to(cuda_dev)[None],
                           med_d.to(cuda_dev)[None], nrm.to(cuda_dev)[None], bg.to(cuda_dev), float(K[0, 0]),
                           float(K[1, 1]), render_mode=mode)
    for k in KEYS:
        ref = _gold(mode, k)
        if ref is None:
            assert got[k] is None, k
            continue
        assert got[k].shape == ref.shape, (k, got[k].shape, ref.shape)
        tol = 2e-5 if "error_map" in k else 0.0
        torch.testing.assert_close(got[k].cpu(), ref, atol=tol, rtol=0)


@pytest.mark.gpu
@pytest.mark.parametrize("mode,use_dn", [("RGB+ED", True), ("RGB", True), ("RGB+ED", False)])
def test_outputs_kernel_backward_matches_oracle_autograd(cuda_dev, mode, use_dn):
    from radegs_b200.outputs import rade_get_outputs
    from tests.util import grad_close_report
    W, H = int(G["W"]), int(G["H"])
    K = _t("K")[0]
    ins = _inputs(mode)
    cpu = [t.clone().requires_grad_(True) for t in ins[:5]]
    ref = O.get_outputs_glue(K, W, H, *cpu, ins[5], render_mode=mode, use_depth_normal=use_dn)
    gpu = [t.clone().to(cuda_dev).requires_grad_(True) for t in ins[:5]]
    got = rade_get_outputs(*gpu, ins[5].to(cuda_dev), float(K[0, 0]), float(K[1, 1]), render_mode=mode,
                           use_depth_normal=use_dn)
    gen = torch.Generator().manual_seed(1)
    lr, lg = 0.0, 0.0
    for k in KEYS:
        if ref[k] is None or k == "accumulation":
            continue
        w = torch.randn(ref[k].shape, generator=gen)
        lr = lr + (ref[k] * w).sum()
        lg = lg + (got[k] * w.to(cuda_dev)).sum()
        torch.testing.assert_close(got[k].detach().cpu(), ref[k].detach(), atol=1e-5 if "error_map" in k else 0.0, rtol=0)
    lr.backward()
    lg.backward()
    for nm, a, b in zip(("render", "alpha", "expected_depths", "median_depths", "normals"), gpu, cpu):
        ok, msg = grad_close_report("v_" + nm, a.grad, b.grad, rel=1e-5, floor=1e-7)
        assert ok, msg
