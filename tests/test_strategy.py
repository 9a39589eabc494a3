"""Host-side densification strategy (gsplat.strategy): the interface collab-splats relies on
(rade_gs_model.py:191-198, 456-458) exercised on CPU tensors."""

import torch

from gsplat.strategy import DefaultStrategy, MCMCStrategy, duplicate, remove, split


def _params(n=50):
    g = torch.Generator().manual_seed(0)
    p = torch.nn.ParameterDict({
        "means": torch.nn.Parameter(torch.randn(n, 3, generator=g)),
        "scales": torch.nn.Parameter(torch.randn(n, 3, generator=g) - 4.0),
        "quats": torch.nn.Parameter(torch.randn(n, 4, generator=g)),
        "opacities": torch.nn.Parameter(torch.randn(n, generator=g)),
        "features_dc": torch.nn.Parameter(torch.randn(n, 3, generator=g)),
    })
    opts = {k: torch.optim.Adam([v], lr=1e-3) for k, v in p.items()}
    for k, v in p.items():          # populate Adam state
        v.grad = torch.ones_like(v)
        opts[k].step()
        v.grad = None
    return p, opts


def test_interface_used_by_the_reference():
    s = DefaultStrategy(absgrad=True)
    assert isinstance(s, DefaultStrategy) and s.absgrad is True
    state = s.initialize_state(scene_scale=2.0)
    assert state["scene_scale"] == 2.0 and state["grad2d"] is None
    m2 = torch.zeros(1, 10, 2, requires_grad=True) * 1.0
    info = {"means2d": m2}
    s.step_pre_backward({}, {}, state, 0, info)
    m2.sum().backward()
    assert info["means2d"].grad is not None        # retain_grad() on a non-leaf
    assert MCMCStrategy is not None


def test_param_surgery_keeps_optimizer_state_aligned():
    p, opts = _params(50)
    state = {"grad2d": torch.arange(50.0), "count": torch.ones(50), "scene_scale": 1.0}
    mask = torch.zeros(50, dtype=torch.bool)
    mask[:5] = True
    duplicate(p, opts, state, mask)
    assert p["means"].shape[0] == 55 and torch.equal(p["means"][50:], p["means"][:5])
    st = opts["means"].state[p["means"]]
    assert st["exp_avg"].shape[0] == 55 and float(st["exp_avg"][50:].abs().max()) == 0.0
    assert state["grad2d"].shape[0] == 55
    mask = torch.zeros(55, dtype=torch.bool)
    mask[10:14] = True
    split(p, opts, state, mask)
    assert p["means"].shape[0] == 55 - 4 + 8 and p["quats"].shape[0] == 59
    keep_n = 59
    mask = torch.zeros(keep_n, dtype=torch.bool)
    mask[::2] = True
    remove(p, opts, state, mask)
    assert p["means"].shape[0] == keep_n - int(mask.sum())
    for k, v in p.items():
        assert opts[k].param_groups[0]["params"][0] is v
        assert opts[k].state[v]["exp_avg"].shape == v.shape


def test_step_post_backward_accumulates_and_refines():
    p, opts = _params(40)
    s = DefaultStrategy(refine_start_iter=0, refine_every=1, grow_grad2d=1e-6, reset_every=1000, prune_opa=0.0)
    s.check_sanity(p, opts)
    state = s.initialize_state()
    m2 = torch.zeros(1, 40, 2, requires_grad=True)
    m2v = m2 * 1.0
    m2v.retain_grad()
    (m2v * torch.linspace(0, 1, 80).reshape(1, 40, 2)).sum().backward()
    radii = torch.ones(1, 40, 2, dtype=torch.int32)
    radii[0, :3] = 0                                   # culled Gaussians do not accumulate statistics
    info = {"means2d": m2v, "radii": radii, "width": 64, "height": 48, "n_cameras": 1}
    s.step_post_backward(p, opts, state, 5, info)
    assert p["means"].shape[0] > 40                    # high-gradient Gaussians were cloned / split
    assert float(state["grad2d"].abs().max()) == 0.0   # statistics reset after a refinement
