"""ctypes + autograd wrapper of the C compositing oracle (oracle/raster_oracle.c).  TEST INFRASTRUCTURE ONLY.

``rasterize_to_pixels`` has the signature and return values of ``oracle.rade_oracle.rasterize_to_pixels`` (the
PyTorch restatement of SURVEY.md rows a10/a11) but runs the per-pixel loops in C on ``threads`` host threads, in
fp32 or fp64 (the dtype of ``means2d``).  ``rade_oracle.rasterization(..., compositor="c")`` routes through it, so a
complete BASELINE-size view can be rendered and differentiated on the CPU in seconds.  The product path never
imports this module.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path
from typing import Optional

import torch
from torch import Tensor

_DIR = Path(__file__).resolve().parent
_SRC = [_DIR / "raster_oracle.c", _DIR / "raster_oracle_impl.h"]
_OUT = _DIR / "_build" / "libraster_oracle.so"
_lib = None
THREADS = os.cpu_count() or 1   # default worker count; tests and the bench override it per call


def build(force: bool = False) -> Path:
    """gcc -O2, strict IEEE (no -ffast-math, no FMA contraction: the fp32 build must round like the PyTorch oracle)."""
    if not force and _OUT.exists() and _OUT.stat().st_mtime > max(p.stat().st_mtime for p in _SRC):
        return _OUT
    _OUT.parent.mkdir(parents=True, exist_ok=True)
    cmd = ["gcc", "-O2", "-std=c11", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-pthread",
           str(_SRC[0]), "-o", str(_OUT), "-lm"]
    subprocess.run(cmd, check=True)
    return _OUT


def load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(build()))
        p, i, ll = C.c_void_p, C.c_int, C.c_longlong
        for sfx in ("_f32", "_f64"):
            getattr(_lib, "ro_rasterize_fwd" + sfx).argtypes = [i] * 7 + [p] * 11 + [ll] + [p] * 9 + [C.c_double, i]
            getattr(_lib, "ro_rasterize_fwd" + sfx).restype = i
            getattr(_lib, "ro_rasterize_bwd" + sfx).argtypes = [i] * 7 + [p] * 11 + [ll] + [p] * 15 + [i]
            getattr(_lib, "ro_rasterize_bwd" + sfx).restype = i
    return _lib


def _ptr(t: Optional[Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class _CRasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means2d, conics, colors, opac, ray_ts, ray_planes, normals, backgrounds, Ks, width, height,
                offsets, flatten_ids, threads, want_fragile, fragile_scale=1.0):
        lib = load()
        dt = means2d.dtype
        assert dt in (torch.float32, torch.float64)
        sfx = "_f32" if dt == torch.float32 else "_f64"
        Cn, N = opac.shape
        D = colors.shape[-1]
        th, tw = offsets.shape[1:]
        M = flatten_ids.numel()
        ins = [t.detach().to(dt).contiguous() for t in (means2d, conics, colors, opac, ray_ts, ray_planes, normals, Ks)]
        bg = None if backgrounds is None else backgrounds.detach().to(dt).contiguous()
        offs = offsets.to(torch.int32).contiguous()
        flat = flatten_ids.to(torch.int32).contiguous()
        out_c = torch.empty(Cn, height, width, D, dtype=dt)
        out_a = torch.empty(Cn, height, width, 1, dtype=dt)
        out_de = torch.empty(Cn, height, width, 1, dtype=dt)
        out_dm = torch.empty(Cn, height, width, 1, dtype=dt)
        out_n = torch.empty(Cn, height, width, 3, dtype=dt)
        last = torch.empty(Cn, height, width, dtype=torch.int32)
        med = torch.empty(Cn, height, width, dtype=torch.int32)
        frag = torch.zeros(Cn, height, width, dtype=torch.uint8) if want_fragile else None
        counters = torch.zeros(2, dtype=torch.int64)
        rc = getattr(lib, "ro_rasterize_fwd" + sfx)(
            Cn, N, D, width, height, tw, th, *[_ptr(t) for t in ins], _ptr(bg), _ptr(offs), _ptr(flat), M,
            _ptr(out_c), _ptr(out_a), _ptr(out_de), _ptr(out_dm), _ptr(out_n), _ptr(last), _ptr(med), _ptr(frag),
            _ptr(counters), float(fragile_scale), int(threads))
        assert rc == 0
        ctx.save_for_backward(*ins, bg, offs, flat, last, med)
        ctx.cfg = (Cn, N, D, width, height, tw, th, M, sfx, int(threads), backgrounds is not None)
        ctx.mark_non_differentiable(last, med, counters)
        if frag is None:
            frag = torch.zeros(0, dtype=torch.uint8)
        ctx.mark_non_differentiable(frag)
        return out_c, out_a, out_de, out_dm, out_n, last, med, frag, counters

    @staticmethod
    def backward(ctx, v_c, v_a, v_de, v_dm, v_n, *_):
        lib = load()
        *ins, bg, offs, flat, last, med = ctx.saved_tensors
        Cn, N, D, width, height, tw, th, M, sfx, threads, has_bg = ctx.cfg
        dt = ins[0].dtype

        def z(g, shape):
            return torch.zeros(shape, dtype=dt) if g is None else g.to(dt).contiguous()

        v_c, v_a = z(v_c, (Cn, height, width, D)), z(v_a, (Cn, height, width, 1))
        v_de, v_dm = z(v_de, (Cn, height, width, 1)), z(v_dm, (Cn, height, width, 1))
        v_n = z(v_n, (Cn, height, width, 3))
        grads = [torch.zeros_like(t) for t in ins[:7]]
        g_bg = torch.zeros_like(bg) if has_bg else None
        rc = getattr(lib, "ro_rasterize_bwd" + sfx)(
            Cn, N, D, width, height, tw, th, *[_ptr(t) for t in ins], _ptr(bg), _ptr(offs), _ptr(flat), M, _ptr(last),
            _ptr(med), _ptr(v_c), _ptr(v_a), _ptr(v_de), _ptr(v_dm), _ptr(v_n), *[_ptr(g) for g in grads], _ptr(g_bg),
            threads)
        assert rc == 0
        return (*grads, g_bg, None, None, None, None, None, None, None, None)


def rasterize_to_pixels(means2d: Tensor, conics: Tensor, colors: Tensor, opacities: Tensor, ray_ts: Tensor,
                        ray_planes: Tensor, normals: Tensor, Ks: Tensor, width: int, height: int, tile_size: int,
                        isect_offsets: Tensor, flatten_ids: Tensor, backgrounds: Optional[Tensor] = None,
                        return_aux: bool = False, tile_window=None, threads: Optional[int] = None,
                        fragile_scale: float = 1.0):
    """Same contract as ``rade_oracle.rasterize_to_pixels`` (colours may be [N,D] or [C,N,D]; ``tile_window`` is not
    supported -- the point of the C oracle is that complete views are affordable).  ``fragile_scale`` widens the margins
    of the ``fragile`` mask (1 = the margins of the PyTorch oracle, right for two fp32 implementations of the same
    arithmetic; an fp32-against-fp64 comparison needs ~10: a transmittance that is a product of hundreds of fp32
    factors is ~1e-5 off, which is the width of the median-crossing margin)."""
    assert tile_size == 16 and tile_window is None
    Cn, N = opacities.shape
    if colors.dim() == 2:
        colors = colors[None].expand(Cn, N, colors.shape[-1])
    out = _CRasterize.apply(means2d, conics, colors, opacities, ray_ts, ray_planes, normals, backgrounds, Ks,
                            int(width), int(height), isect_offsets, flatten_ids, int(threads or THREADS),
                            bool(return_aux), float(fragile_scale))
    res = tuple(out[:5])
    if return_aux:
        aux = dict(last_ids=out[5], median_ids=out[6], fragile=out[7].bool(), n_tested=int(out[8][0]),
                   n_contrib=int(out[8][1]))
        return res + (aux,)
    return res
