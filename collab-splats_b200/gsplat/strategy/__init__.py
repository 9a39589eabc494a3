"""Densification strategies with the interface collab-splats relies on.

The reference only touches ``DefaultStrategy`` through ``isinstance(self.strategy, DefaultStrategy)`` and
``.absgrad`` (collab_splats/models/rade_gs_model.py:456-458) and through
``strategy.step_pre_backward(params, optimizers, state, step, info)`` (rade_gs_model.py:191-198); nerfstudio's
``SplatfactoModel`` additionally calls ``initialize_state`` / ``check_sanity`` / ``step_post_backward`` and
imports ``MCMCStrategy``.  These are host-side bookkeeping over torch tensors (out of scope for
kernels); they are restated here device-agnostically from the published gsplat 1.5 behaviour so that the
shim is a complete ``gsplat.strategy`` namespace.  The per-Gaussian statistics they consume
(``info["means2d"].grad`` / ``.absgrad``, ``info["radii"]``) are produced by the CUDA kernels.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Callable, Dict, Union

import torch
from torch import Tensor

Params = Union[Dict[str, torch.nn.Parameter], torch.nn.ParameterDict]


@dataclass
class Strategy:
    def check_sanity(self, params: Params, optimizers: Dict[str, torch.optim.Optimizer]):
        trainable = {k for k, v in params.items() if v.requires_grad}
        assert trainable == set(optimizers.keys()), (trainable, set(optimizers.keys()))
        for opt in optimizers.values():
            assert len(opt.param_groups) == 1, "each optimizer must hold exactly one param group"

    def step_pre_backward(self, *args, **kwargs):
        pass

    def step_post_backward(self, *args, **kwargs):
        pass


# ------------------------------------------------------------------------------------------------ param surgery
@torch.no_grad()
def _update_param_with_optimizer(param_fn: Callable[[str, Tensor], Tensor], state_fn: Callable[[str, Tensor], Tensor],
                                 params: Params, optimizers: Dict[str, torch.optim.Optimizer], names=None):
    names = list(params.keys()) if names is None else names
    for name in names:
        old = params[name]
        new = torch.nn.Parameter(param_fn(name, old), requires_grad=old.requires_grad)
        if name in optimizers:
            opt = optimizers[name]
            for group in opt.param_groups:
                for i, p in enumerate(group["params"]):
                    if p is old:
                        st = opt.state.pop(p, {})
                        for k, v in list(st.items()):
                            if isinstance(v, Tensor) and v.dim() > 0 and v.shape[0] == old.shape[0]:
                                st[k] = state_fn(k, v)
                        group["params"][i] = new
                        opt.state[new] = st
        params[name] = new


@torch.no_grad()
def duplicate(params, optimizers, state, mask: Tensor):
    sel = torch.where(mask)[0]
    _update_param_with_optimizer(lambda n, p: torch.cat([p, p[sel]]),
                                 lambda k, v: torch.cat([v, torch.zeros((len(sel), *v.shape[1:]), device=v.device,
                                                                        dtype=v.dtype)]), params, optimizers)
    for k, v in state.items():
        if isinstance(v, Tensor) and v.dim() > 0:
            state[k] = torch.cat([v, v[sel]])


@torch.no_grad()
def split(params, optimizers, state, mask: Tensor, revised_opacity: bool = False):
    dev = mask.device
    sel, rest = torch.where(mask)[0], torch.where(~mask)[0]
    scales = torch.exp(params["scales"][sel])
    q = torch.nn.functional.normalize(params["quats"][sel], dim=-1)
    w, x, y, z = q.unbind(-1)
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                     2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                     2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], dim=-1).reshape(-1, 3, 3)
    samples = torch.einsum("nij,nj,bnj->bni", R, scales, torch.randn(2, len(sel), 3, device=dev))

    def param_fn(name, p):
        if name == "means":
            new = (p[sel] + samples).reshape(-1, 3)
        elif name == "scales":
            new = torch.log(scales / 1.6).repeat(2, 1)
        elif name == "opacities" and revised_opacity:
            new = torch.logit(1.0 - torch.sqrt(1.0 - torch.sigmoid(p[sel]))).repeat(2, *([1] * (p.dim() - 1)))
        else:
            new = p[sel].repeat(2, *([1] * (p.dim() - 1)))
        return torch.cat([p[rest], new])

    def state_fn(k, v):
        return torch.cat([v[rest], torch.zeros((2 * len(sel), *v.shape[1:]), device=v.device, dtype=v.dtype)])

    _update_param_with_optimizer(param_fn, state_fn, params, optimizers)
    for k, v in state.items():
        if isinstance(v, Tensor) and v.dim() > 0:
            state[k] = torch.cat([v[rest], v[sel].repeat(2, *([1] * (v.dim() - 1)))])


@torch.no_grad()
def remove(params, optimizers, state, mask: Tensor):
    keep = torch.where(~mask)[0]
    _update_param_with_optimizer(lambda n, p: p[keep], lambda k, v: v[keep], params, optimizers)
    for k, v in state.items():
        if isinstance(v, Tensor) and v.dim() > 0:
            state[k] = v[keep]


@torch.no_grad()
def reset_opa(params, optimizers, state, value: float):
    cap = torch.logit(torch.tensor(value)).item()
    _update_param_with_optimizer(lambda n, p: torch.clamp(p, max=cap), lambda k, v: torch.zeros_like(v), params,
                                 optimizers, names=["opacities"])


# ------------------------------------------------------------------------------------------------ strategies
@dataclass
class DefaultStrategy(Strategy):
    """Adaptive density control of 3DGS (clone / split / prune / opacity reset)."""
    prune_opa: float = 0.005
    grow_grad2d: float = 0.0002
    grow_scale3d: float = 0.01
    grow_scale2d: float = 0.05
    prune_scale3d: float = 0.1
    prune_scale2d: float = 0.15
    refine_scale2d_stop_iter: int = 0
    refine_start_iter: int = 500
    refine_stop_iter: int = 15_000
    reset_every: int = 3000
    refine_every: int = 100
    pause_refine_after_reset: int = 0
    absgrad: bool = False
    revised_opacity: bool = False
    verbose: bool = False
    key_for_gradient: str = "means2d"

    def initialize_state(self, scene_scale: float = 1.0) -> Dict[str, Any]:
        state = {"grad2d": None, "count": None, "scene_scale": scene_scale}
        if self.refine_scale2d_stop_iter > 0:
            state["radii"] = None
        return state

    def check_sanity(self, params, optimizers):
        super().check_sanity(params, optimizers)
        for key in ("means", "scales", "quats", "opacities"):
            assert key in params, f"{key} is required in params but missing"

    def step_pre_backward(self, params, optimizers, state, step: int, info: Dict[str, Any]):
        assert self.key_for_gradient in info, f"{self.key_for_gradient} missing from the rasterization meta"
        info[self.key_for_gradient].retain_grad()

    def step_post_backward(self, params, optimizers, state, step: int, info: Dict[str, Any], packed: bool = False):
        if step >= self.refine_stop_iter:
            return
        self._update_state(params, state, info, packed=packed)
        if (step > self.refine_start_iter and step % self.refine_every == 0
                and step % self.reset_every >= self.pause_refine_after_reset):
            n_dupli, n_split = self._grow_gs(params, optimizers, state, step)
            n_prune = self._prune_gs(params, optimizers, state, step)
            if self.verbose:
                print(f"step {step}: {n_dupli} duplicated, {n_split} split, {n_prune} pruned, "
                      f"{len(params['means'])} GSs")
            state["grad2d"].zero_()
            state["count"].zero_()
            if self.refine_scale2d_stop_iter > 0:
                state["radii"].zero_()
        if step % self.reset_every == 0 and step > 0:
            reset_opa(params, optimizers, state, value=self.prune_opa * 2.0)

    @torch.no_grad()
    def _update_state(self, params, state, info, packed: bool = False):
        for key in ("width", "height", "n_cameras", "radii", self.key_for_gradient):
            assert key in info, f"{key} is required in the rasterization meta but missing"
        if packed:
            raise NotImplementedError("packed=True is not on the collab-splats path")
        m2 = info[self.key_for_gradient]
        raw = m2.absgrad if self.absgrad else m2.grad
        n = len(list(params.values())[0])
        dev = raw.device
        if state["grad2d"] is None:
            state["grad2d"] = torch.zeros(n, device=dev)
        if state["count"] is None:
            state["count"] = torch.zeros(n, device=dev)
        if self.refine_scale2d_stop_iter > 0 and state["radii"] is None:
            state["radii"] = torch.zeros(n, device=dev)
        radii = info["radii"]
        sx = info["width"] / 2.0 * info["n_cameras"]
        sy = info["height"] / 2.0 * info["n_cameras"]
        if raw.is_cuda and radii.dim() == 3:
            # one fused pass over [C,N] (csrc/stats.cu: rs_densify_stats) instead of the clone / scale / mask / gather /
            # norm / index_add chain below
            from radegs_b200 import backend as _be
            lib = _be.load()
            C_, N_ = radii.shape[:2]
            g = raw.to(torch.float32).contiguous()
            r32 = radii.to(torch.int32).contiguous()
            with torch.cuda.device(dev):
                _be.check(lib.rs_densify_stats(
                    _be.ptr(g), _be.ptr(r32), C_, N_, sx, sy, 1.0 / float(max(info["width"], info["height"])),
                    _be.ptr(state["grad2d"]), _be.ptr(state["count"]),
                    _be.ptr(state["radii"]) if self.refine_scale2d_stop_iter > 0 else None, _be.stream_ptr(dev)),
                    "rs_densify_stats")
            return
        # host tensors (CPU tests of the bookkeeping): the published gsplat sequence
        grads = raw.clone()
        grads[..., 0] *= sx
        grads[..., 1] *= sy
        sel = (radii > 0).all(dim=-1) if radii.dim() == 3 else radii > 0      # [C,N]
        gs_ids = torch.where(sel)[1]
        state["grad2d"].index_add_(0, gs_ids, grads[sel].norm(dim=-1))
        state["count"].index_add_(0, gs_ids, torch.ones_like(gs_ids, dtype=torch.float32))
        if self.refine_scale2d_stop_iter > 0:
            r = radii[sel].float()
            r = (r.max(dim=-1).values if r.dim() == 2 else r) / float(max(info["width"], info["height"]))
            state["radii"][gs_ids] = torch.maximum(state["radii"][gs_ids], r)

    @torch.no_grad()
    def _grow_gs(self, params, optimizers, state, step: int):
        grads = state["grad2d"] / state["count"].clamp_min(1)
        is_grad_high = grads > self.grow_grad2d
        is_small = torch.exp(params["scales"]).max(dim=-1).values <= self.grow_scale3d * state["scene_scale"]
        is_dupli = is_grad_high & is_small
        n_dupli = int(is_dupli.sum().item())
        is_split = is_grad_high & ~is_small
        if step < self.refine_scale2d_stop_iter:
            is_split |= state["radii"] > self.grow_scale2d
        n_split = int(is_split.sum().item())
        if n_dupli > 0:
            duplicate(params, optimizers, state, is_dupli)
        is_split = torch.cat([is_split, torch.zeros(n_dupli, dtype=torch.bool, device=is_split.device)])
        if n_split > 0:
            split(params, optimizers, state, is_split, revised_opacity=self.revised_opacity)
        return n_dupli, n_split

    @torch.no_grad()
    def _prune_gs(self, params, optimizers, state, step: int):
        is_prune = torch.sigmoid(params["opacities"].flatten()) < self.prune_opa
        if step > self.reset_every:
            is_too_big = torch.exp(params["scales"]).max(dim=-1).values > self.prune_scale3d * state["scene_scale"]
            if step < self.refine_scale2d_stop_iter:
                is_too_big |= state["radii"] > self.prune_scale2d
            is_prune = is_prune | is_too_big
        n_prune = int(is_prune.sum().item())
        if n_prune > 0:
            remove(params, optimizers, state, is_prune)
        return n_prune


@dataclass
class MCMCStrategy(Strategy):
    """Interface stub: nerfstudio's splatfacto imports the name; collab-splats' method configs never select it
    (collab_splats/configs/rade_gs_method.py:23-89 use the default strategy), and its relocation kernels are not
    on the hot path (last rows)."""
    cap_max: int = 1_000_000
    noise_lr: float = 5e5
    refine_start_iter: int = 500
    refine_stop_iter: int = 25_000
    refine_every: int = 100
    min_opacity: float = 0.005
    verbose: bool = False

    def initialize_state(self) -> Dict[str, Any]:
        raise NotImplementedError("MCMCStrategy is not on the collab-splats path")

    def step_post_backward(self, *args, **kwargs):
        raise NotImplementedError("MCMCStrategy is not on the collab-splats path")
