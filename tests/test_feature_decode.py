"""Feature decode + cosine loss.

CPU: oracle/feature_oracle.py against tests/golden/feature_decoder.npz, which was produced by the reference's own
``TwoLayerMLP`` class (collab_splats/utils/features.py:408-478, loaded from /root/reference by
tests/golden/make_feature_golden.py) -- this pins the oracle.
GPU: csrc/feature_decode.cu through ``radegs_b200.feature_decode`` against the golden vectors and against the
oracle on other shapes (F = 13 and 64, ragged channel counts, up- and down-sampling branches, resize_factor 8).
Tolerance (floating point; fp32 accumulation order differs): max-abs 1e-5 + rel 1e-4 on decoded features, loss to
1e-6 relative, gradients to 1e-4 of their scale."""

from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import feature_oracle as fo

GOLD = Path(__file__).resolve().parent / "golden" / "feature_decoder.npz"
DIMS = {"clip": (96, 9, 12), "dino": (40, 7, 10)}


def _gold():
    g = np.load(GOLD)
    return {k: torch.from_numpy(g[k]) for k in g.files}


def _branches(g):
    return {k: (g[f"w_{k}"], g[f"b_{k}"]) for k in DIMS}


# ----------------------------------------------------------------------------- oracle vs the reference's decoder (CPU)
def test_oracle_decode_matches_reference_decoder():
    g = _gold()
    out = fo.decode_features(g["features"], g["w_hidden"], g["b_hidden"], _branches(g), DIMS, "clip")
    for k in DIMS:
        assert out[k].shape == g[f"decoded_{k}"].shape
        torch.testing.assert_close(out[k], g[f"decoded_{k}"], atol=1e-6, rtol=1e-5)


def test_oracle_loss_and_gradients_match_reference_decoder():
    g = _gold()
    feats = g["features"].clone().requires_grad_(True)
    w1, b1 = g["w_hidden"].clone().requires_grad_(True), g["b_hidden"].clone().requires_grad_(True)
    br = {k: (w.clone().requires_grad_(True), b.clone().requires_grad_(True)) for k, (w, b) in _branches(g).items()}
    loss = fo.features_loss(feats, w1, b1, br, DIMS, "clip", {k: g[f"gt_{k}"] for k in DIMS})
    loss.backward()
    torch.testing.assert_close(loss.detach(), g["loss"], atol=1e-9, rtol=1e-6)
    torch.testing.assert_close(feats.grad, g["v_features"], atol=1e-10, rtol=1e-4)
    torch.testing.assert_close(w1.grad, g["v_w_hidden"], atol=1e-9, rtol=1e-4)
    torch.testing.assert_close(b1.grad, g["v_b_hidden"], atol=1e-9, rtol=1e-4)
    for k in DIMS:
        torch.testing.assert_close(br[k][0].grad, g[f"v_w_{k}"], atol=1e-9, rtol=1e-4)
        torch.testing.assert_close(br[k][1].grad, g[f"v_b_{k}"], atol=1e-9, rtol=1e-4)


def test_oracle_per_gaussian_equals_1x1_conv():
    g = _gold()
    x = g["per_gauss_in"]
    out = fo.mlp_forward(x.t()[None, :, :, None], g["w_hidden"], g["b_hidden"], _branches(g))
    for k in DIMS:
        torch.testing.assert_close(out[k][0, :, :, 0].t(), g[f"per_gauss_{k}"], atol=1e-6, rtol=1e-5)


# ----------------------------------------------------------------------------- CUDA (GPU)
def _decoder(dev, Fin, Hd, dims, g=None, seed=0):
    from radegs_b200 import feature_decode as fd
    torch.manual_seed(seed)
    dec = fd.TwoLayerMLP(Fin, Hd, dims)
    if g is not None:
        sd = {"hidden_conv.weight": g["w_hidden"][:, :, None, None], "hidden_conv.bias": g["b_hidden"]}
        for k in dims:
            sd[f"feature_branch_dict.{k}.weight"] = g[f"w_{k}"][:, :, None, None]
            sd[f"feature_branch_dict.{k}.bias"] = g[f"b_{k}"]
        dec.load_state_dict(sd)               # the reference's state_dict layout loads unchanged
    return dec.to(dev)


def _scale_close(name, got, ref, rel=1e-4):
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    assert got.shape == ref.shape, (name, got.shape, ref.shape)
    tol = rel * ref.abs().max().item() + 1e-12
    err = (got - ref).abs().max().item()
    assert err <= tol, f"{name}: max err {err:.3e} > {tol:.3e}"


@pytest.mark.gpu
def test_cuda_matches_golden_from_reference_decoder(cuda_dev):
    from radegs_b200 import feature_decode as fd
    g = _gold()
    dec = _decoder(cuda_dev, 13, 64, DIMS, g)
    # features live inside a wider render row (rgb | features | depth), as rasterization() returns them
    H, W, F = g["features"].shape
    render = torch.zeros(H, W, 3 + F + 1)
    render[..., 3:3 + F] = g["features"]
    render = render.to(cuda_dev).requires_grad_(True)
    with pytest.raises(NotImplementedError, match="features_loss"):     # would silently cut the graph: refused
        fd.decode_features(render, dec, DIMS, "clip", ch0=3, n_features=F)
    with torch.no_grad():
        out = fd.decode_features(render, dec, DIMS, "clip", ch0=3, n_features=F)
    for k in DIMS:
        torch.testing.assert_close(out[k].cpu(), g[f"decoded_{k}"], atol=1e-5, rtol=1e-4)
    loss = fd.features_loss(render, dec, DIMS, "clip", {k: g[f"gt_{k}"].to(cuda_dev) for k in DIMS}, ch0=3,
                            n_features=F)
    (loss * 2.0).backward()                                   # the incoming gradient scales everything
    assert abs(loss.item() - g["loss"].item()) <= 1e-6 * abs(g["loss"].item())
    _scale_close("v_features", render.grad[..., 3:3 + F], g["v_features"] * 2)
    assert float(render.grad[..., :3].abs().max()) == 0 and float(render.grad[..., 3 + F:].abs().max()) == 0
    _scale_close("v_w_hidden", dec.hidden_conv.weight.grad.view(64, F), g["v_w_hidden"] * 2)
    _scale_close("v_b_hidden", dec.hidden_conv.bias.grad, g["v_b_hidden"] * 2)
    for k in DIMS:
        conv = dec.feature_branch_dict[k]
        _scale_close(f"v_w_{k}", conv.weight.grad.view(DIMS[k][0], 64), g[f"v_w_{k}"] * 2)
        _scale_close(f"v_b_{k}", conv.bias.grad, g[f"v_b_{k}"] * 2)
    pg = dec.per_gaussian_forward(g["per_gauss_in"].to(cuda_dev))
    for k in DIMS:
        torch.testing.assert_close(pg[k].cpu(), g[f"per_gauss_{k}"], atol=1e-5, rtol=1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("Fin,Hd,H,W,dims,main", [
    (64, 64, 135, 240, {"clip": (200, 9, 17), "dino": (70, 19, 34)}, "clip"),       # config-3-like, up-sampled branch
    (13, 32, 50, 70, {"a": (33, 50, 70), "b": (5, 3, 4), "c": (129, 60, 80)}, "b"),   # ragged counts, tiny main map
    (7, 128, 31, 45, {"only": (64, 16, 23)}, "only"),
])
def test_cuda_matches_oracle(cuda_dev, Fin, Hd, H, W, dims, main):
    from radegs_b200 import feature_decode as fd
    dec = _decoder(cuda_dev, Fin, Hd, dims, seed=3)
    torch.manual_seed(7)
    feats = torch.randn(H, W, Fin) * 0.5
    gt = {k: torch.randn(*v) for k, v in dims.items()}
    w1, b1, br = dec._flat()
    cw1, cb1 = w1.detach().cpu().requires_grad_(True), b1.detach().cpu().requires_grad_(True)
    cbr = {k: (w.detach().cpu().requires_grad_(True), b.detach().cpu().requires_grad_(True)) for k, (w, b) in br.items()}
    cf = feats.clone().requires_grad_(True)
    ref_dec = fo.decode_features(cf, cw1, cb1, cbr, dims, main)
    ref_loss = fo.features_loss(cf, cw1, cb1, cbr, dims, main, gt, reg_lambda=0.25, loss_lambda=0.5)
    ref_loss.backward()
    df = feats.to(cuda_dev).requires_grad_(True)
    with torch.no_grad():
        out = fd.decode_features(df, dec, dims, main)
    for k in dims:
        torch.testing.assert_close(out[k].cpu(), ref_dec[k].detach(), atol=1e-5, rtol=1e-4)
    loss = fd.features_loss(df, dec, dims, main, {k: v.to(cuda_dev) for k, v in gt.items()},
                            features_regularization_lambda=0.25, features_loss_lambda=0.5)
    loss.backward()
    assert abs(loss.item() - ref_loss.item()) <= 2e-6 * abs(ref_loss.item())
    _scale_close("v_features", df.grad, cf.grad)
    _scale_close("v_w1", dec.hidden_conv.weight.grad.view(Hd, Fin), cw1.grad)
    _scale_close("v_b1", dec.hidden_conv.bias.grad, cb1.grad)
    for k in dims:
        conv = dec.feature_branch_dict[k]
        _scale_close(f"v_w_{k}", conv.weight.grad.view(dims[k][0], Hd), cbr[k][0].grad)
        _scale_close(f"v_b_{k}", conv.bias.grad, cbr[k][1].grad)


@pytest.mark.gpu
def test_cuda_decode_with_resize_factor(cuda_dev):
    """The viewer path: decode_features(outs["features"], resize_factor=8.0), rade_features_model.py:509-511."""
    from radegs_b200 import feature_decode as fd
    dims = {"clip": (48, 6, 9), "dino": (24, 11, 13)}
    dec = _decoder(cuda_dev, 13, 64, dims, seed=1)
    torch.manual_seed(2)
    feats = torch.randn(60, 90, 13)
    w1, b1, br = dec._flat()
    ref = fo.decode_features(feats, w1.detach().cpu(), b1.detach().cpu(),
                             {k: (w.detach().cpu(), b.detach().cpu()) for k, (w, b) in br.items()}, dims, "clip",
                             resize_factor=8.0)
    with torch.no_grad():
        out = fd.decode_features(feats.to(cuda_dev), dec, dims, "clip", resize_factor=8.0)
        mlp = dec(feats.to(cuda_dev).permute(2, 0, 1)[None].contiguous())      # TwoLayerMLP.forward on a [1,F,H,W] map
    ref_mlp = fo.mlp_forward(feats.permute(2, 0, 1)[None], w1.detach().cpu(), b1.detach().cpu(),
                             {k: (w.detach().cpu(), b.detach().cpu()) for k, (w, b) in br.items()})
    for k in dims:
        torch.testing.assert_close(mlp[k].cpu(), ref_mlp[k], atol=1e-5, rtol=1e-4)
    assert out["clip"].shape == (48, 48, 72) and out["dino"].shape == (24, 11, 13)
    for k in dims:
        torch.testing.assert_close(out[k].cpu(), ref[k], atol=1e-5, rtol=1e-4)


@pytest.mark.gpu
def test_feature_decode_errors_are_loud(cuda_dev):
    from radegs_b200 import feature_decode as fd
    dims = {"clip": (8, 4, 4)}
    dec = _decoder(cuda_dev, 13, 16, dims)
    with torch.no_grad():
        with pytest.raises(RuntimeError, match="no CPU path"):
            fd.decode_features(torch.zeros(8, 8, 13), dec, dims, "clip")
        with pytest.raises(ValueError):
            fd.decode_features(torch.zeros(8, 8, 12, device=cuda_dev), dec, dims, "clip")
    with pytest.raises(ValueError):
        fd.features_loss(torch.zeros(8, 8, 13, device=cuda_dev), dec, dims, "clip",
                         {"clip": torch.zeros(8, 5, 4, device=cuda_dev)})
    with torch.no_grad():
        with pytest.raises(NotImplementedError):
            fd.decode_features(torch.zeros(8, 8, 200, device=cuda_dev), _decoder(cuda_dev, 200, 16, dims), dims, "clip")
