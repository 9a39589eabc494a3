"""CPU checks of the per-element device math (csrc/rade_math.cuh compiled for the host, tests/hostmath):

* the exact-order section of the projection reproduces the oracle bit for bit in fp32
  (radii, means2d, depths, conics, compensations) -- the contract that makes tile lists reproducible;
* the hand-derived VJP of the projection (incl. the RaDe ray-plane / normal terms and the camera gradient)
  and of the SH evaluation match oracle autograd in fp64.
"""

import ctypes

import numpy as np
import pytest
import torch

from oracle import rade_oracle as O
from radegs_b200 import scenes


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _scene(n=20000, w=256, h=192, views=2, spread=1.5, seed=1235):
    cfg = scenes.SceneConfig("t", n, w, h, views, 3, 0, seed)
    gs, vm, Ks = scenes.make_scene(cfg)
    means, quats, scales, _, _ = scenes.activate(gs, 3)
    return cfg, means * spread, quats, scales, vm, Ks


def _run_fwd(lib, dt, means, quats, scales, vm, K, W, H, eps=0.3, near=0.01, far=1e10, clip=0.0):
    N = means.shape[0]
    npdt = np.float32 if dt == "f32" else np.float64
    cf = ctypes.c_float if dt == "f32" else ctypes.c_double
    arrs = [np.ascontiguousarray(a.detach().numpy().astype(npdt)) for a in (means, quats, scales, vm, K)]
    radii = np.zeros((N, 2), np.int32)
    outs = [np.zeros(s, npdt) for s in ((N, 2), (N,), (N, 3), (N,), (N,), (N, 2), (N, 3))]
    getattr(lib, "hm_project_fwd_" + dt)(*[P(a) for a in arrs], N, W, H, cf(eps), cf(near), cf(far), cf(clip),
                                         P(radii), *[P(o) for o in outs])
    return [radii] + outs


def test_projection_exact_section_is_bit_exact_fp32(hostmath):
    cfg, means, quats, scales, vm, Ks = _scene()
    names = ["radii", "means2d", "depths", "conics", "compensations"]
    for c in range(2):
        ref = O.fully_fused_projection(means, quats, scales, vm[c:c + 1], Ks[c:c + 1], cfg.width, cfg.height,
                                       calc_compensations=True)
        out = _run_fwd(hostmath, "f32", means, quats, scales, vm[c], Ks[c], cfg.width, cfg.height)
        n_valid = int((out[0] > 0).all(-1).sum())
        assert 1000 < n_valid < cfg.n_gaussians      # the scene exercises both culled and visible Gaussians
        for nm, r, o in zip(names, ref, out):
            assert np.array_equal(r[0].numpy(), o), f"camera {c}: {nm} differs from the oracle"
        for r, o in zip(ref[5:], out[5:]):           # RaDe terms: ordinary arithmetic
            np.testing.assert_allclose(o, r[0].numpy(), rtol=2e-5, atol=2e-6)


def test_projection_vjp_matches_autograd_fp64(hostmath):
    cfg, means, quats, scales, vm, Ks = _scene(n=5000)
    md, qd, sd, vmd, Kd = [t.double() for t in (means, quats, scales, vm, Ks)]
    for t in (md, qd, sd):
        t.requires_grad_(True)
    vm0 = vmd[0:1].clone().requires_grad_(True)
    ref = O.fully_fused_projection(md, qd, sd, vm0, Kd[0:1], cfg.width, cfg.height, calc_compensations=True)
    out = _run_fwd(hostmath, "f64", md, qd, sd, vmd[0], Kd[0], cfg.width, cfg.height)
    for r, o in zip(ref, out):
        np.testing.assert_allclose(o, r[0].detach().numpy(), rtol=1e-11, atol=1e-12)
    g = torch.Generator().manual_seed(1)
    vs = [torch.randn(r.shape, generator=g, dtype=torch.float64) for r in ref[1:]]
    sum((r * v).sum() for r, v in zip(ref[1:], vs)).backward()
    N = md.shape[0]
    v_means, v_quats, v_scales = np.zeros((N, 3)), np.zeros((N, 4)), np.zeros((N, 3))
    v_W, v_t = np.zeros(9), np.zeros(3)
    ins = [np.ascontiguousarray(a.detach().numpy()) for a in (md, qd, sd, vmd[0], Kd[0])]
    gv = [np.ascontiguousarray(v[0].numpy()) for v in vs]
    cd = ctypes.c_double
    hostmath.hm_project_bwd_f64(*[P(a) for a in ins], N, cfg.width, cfg.height, cd(0.3), cd(0.01), cd(1e10), cd(0.0),
                                *[P(a) for a in gv], P(v_means), P(v_quats), P(v_scales), P(v_W), P(v_t))

    def rel(a, b):
        return np.abs(a - b).max() / (np.abs(b).max() + 1e-300)

    assert rel(v_means, md.grad.numpy()) < 1e-10
    assert rel(v_quats, qd.grad.numpy()) < 1e-10
    assert rel(v_scales, sd.grad.numpy()) < 1e-10
    assert rel(v_W.reshape(3, 3), vm0.grad[0, :3, :3].numpy()) < 1e-9
    assert rel(v_t, vm0.grad[0, :3, 3].numpy()) < 1e-9


@pytest.mark.parametrize("degree", [0, 1, 2, 3])
def test_sh_forward_and_vjp_fp64(hostmath, degree):
    g = torch.Generator().manual_seed(degree)
    N, K = 2000, 16
    dirs = torch.randn(N, 3, generator=g, dtype=torch.float64).requires_grad_(True)
    coeffs = torch.randn(N, K, 3, generator=g, dtype=torch.float64).requires_grad_(True)
    v = torch.randn(N, 3, generator=g, dtype=torch.float64)
    ref = O.spherical_harmonics(degree, dirs, coeffs)
    (ref * v).sum().backward()
    d, c = np.ascontiguousarray(dirs.detach().numpy()), np.ascontiguousarray(coeffs.detach().numpy())
    out = np.zeros((N, 3))
    hostmath.hm_sh_fwd_f64(degree, K, N, P(d), P(c), P(out))
    np.testing.assert_allclose(out, ref.detach().numpy(), rtol=1e-12, atol=1e-13)
    vc, vd, vv = np.zeros((N, K, 3)), np.zeros((N, 3)), np.ascontiguousarray(v.numpy())
    hostmath.hm_sh_bwd_f64(degree, K, N, P(d), P(c), P(vv), P(vc), P(vd))
    np.testing.assert_allclose(vc, coeffs.grad.numpy(), rtol=1e-11, atol=1e-12)
    ref_vd = np.zeros((N, 3)) if dirs.grad is None else dirs.grad.numpy()   # degree 0 does not depend on dirs
    np.testing.assert_allclose(vd, ref_vd, rtol=1e-9, atol=1e-11)
