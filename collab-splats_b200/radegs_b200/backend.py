"""ctypes binding of librade_b200.so (include/rade_b200.h).

This is the *only* compute backend: there is no CPU path and no fallback.  If the library has not
been built (``python __graft_entry__.py build``) loading fails loudly, and every call that returns a
non-zero status raises ``RuntimeError`` with the library's own message.
"""

from __future__ import annotations

import ctypes as C
from pathlib import Path

import torch

_LIB_PATH = Path(__file__).resolve().parent.parent / "lib" / "librade_b200.so"
_lib = None

_p = C.c_void_p
_i = C.c_int
_ll = C.c_longlong
_f = C.c_float

_SIGNATURES = {
    "rs_version": (C.c_int, []),
    "rs_error_string": (C.c_char_p, [_i]),
    "rs_last_cuda_error": (C.c_int, []),
    "rs_launch_count": (C.c_ulonglong, []),
    "rs_timing_enable": (None, [_i]),
    "rs_fma_peak_probe": (_i, [_i, _i, _p, _p]),
    "rs_timing_collect": (_i, [C.c_char_p, _p, _p, _i]),
    "rs_project_fwd": (_i, [_p] * 5 + [_i] * 4 + [_f] * 4 + [_i] + [_p] * 8 + [_p]),
    "rs_project_bwd": (_i, [_p] * 5 + [_i] * 4 + [_f] * 4 + [_p] * 7 + [_p] * 4 + [_p]),
    "rs_sh_fwd": (_i, [_i, _i, _ll, _ll, _p, _p, _p, _p, _p]),
    "rs_sh_bwd": (_i, [_i, _i, _ll, _ll, _p, _p, _p, _p, _p, _p, _p]),
    "rs_tile_bits": (_i, [_i, _i]),
    "rs_isect_count": (_i, [_p, _p, _ll, _i, _i, _p, _p]),
    "rs_cumsum_temp_bytes": (_ll, [_ll]),
    "rs_cumsum_i32_i64": (_i, [_p, _p, _ll, _p, _ll, _p]),
    "rs_isect_emit": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p]),
    "rs_offset_encode": (_i, [_p, _ll, _i, _i, _i, _p, _p]),
    "rs_cumsum_gather_i32_i64": (_i, [_p, _p, _p, _ll, _p, _ll, _p]),
    "rs_isect_emit_ordered": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p]),
    "rs_argsort_u32": (_i, [_p, _p, _p, _p, _ll, _i, _i, _p, _ll, _p]),
    "rs_scale_unless_one": (_i, [_p, _p, _i, _p, _p]),
    "rs_zero_bytes": (_i, [_p, _ll, _p]),
    "rs_isect_emit_ordered32": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p]),
    "rs_sort_pairs_u32": (_i, [_p, _p, _p, _p, _ll, _i, _i, _p, _ll, _p]),
    "rs_isect_finish32": (_i, [_p, _p, _p, _ll, _i, _i, _i, _p, _p, _p]),
    "rs_sort_pairs_temp_bytes": (_ll, [_ll, _i, _i]),
    "rs_sort_pairs": (_i, [_p, _p, _p, _p, _ll, _i, _i, _p, _ll, _p]),
    "rs_raster_padded_channels": (_i, [_i]),
    "rs_pack_geom": (_i, [_p, _p, _p, _i, _p, _i, _i, _p, _p, _p, _p, _p, _i, _p]),
    "rs_pack_colors": (_i, [_p, _ll, _i, _i, _p, _p]),
    "rs_rasterize_fwd": (_i, [_p, _p, _i, _i, _i, _p, _p] + [_i] * 6 + [_p, _p, _ll] + [_p] * 8 + [_i, _p, _p, _p]),
    "rs_rasterize_bwd": (_i, [_p, _p, _i, _i, _i, _p, _p] + [_i] * 6 + [_p, _p, _ll] + [_p] * 4 + [_p] * 5
                         + [_p, _p, _p] + [_i, _p, _p]),
    "rs_isect_emit_ordered_bounded": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _ll, _p, _p, _p]),
    "rs_sort_pairs_dev": (_i, [_p, _p, _p, _p, _ll, _p, _i, _i, _p, _ll, _p]),
    "rs_offset_encode_dev": (_i, [_p, _ll, _p, _i, _i, _i, _p, _p, _p]),
    "rs_unpack_geom_grad": (_i, [_p, _p, _p, _i, _i, _p, _i, _p] + [_p] * 9 + [_i, _i, _p]),
    "rs_sh_colors_fwd": (_i, [_i, _i, _i, _i] + [_p] * 6 + [_p]),
    "rs_sh_colors_bwd": (_i, [_i, _i, _i, _i] + [_p] * 5 + [_i] + [_p] * 3 + [_p]),
    "rs_unpack_colors_grad": (_i, [_p, _ll, _i, _i, _p, _p]),
    "rs_sh_region_bytes": (_ll, [_i, _i]),
    "rs_sh_colors_bwd_local": (_i, [_i, _i, _i, _i] + [_p] * 5 + [_i] + [_p] * 3 + [_p]),
    "rs_sh_coeffs_gather": (_i, [_i, _i, _i, _p, _p, _p, _i, _p, _p]),
    "rs_peer_alloc": (_i, [_ll, _p]),
    "rs_peer_free": (_i, [_p]),
    "rs_peer_handle_bytes": (_i, []),
    "rs_peer_export": (_i, [_p, _p]),
    "rs_peer_import": (_i, [_p, _p]),
    "rs_peer_unimport": (_i, [_p]),
    "rs_peer_copy": (_i, [_p, _p, _ll, _p]),
    "rs_peer_push": (_i, [_p, _i, _p, _ll, _i, _p]),
    "rs_peer_allreduce": (_i, [_p, _i, _i, _ll, _i, _p]),
    "rs_peer_signal": (_i, [_p, _i, _i, C.c_ulonglong, _p]),
    "rs_peer_wait": (_i, [_p, _i, C.c_ulonglong, _i, _p, _p]),
    "rs_densify_stats": (_i, [_p, _p, _i, _i, _f, _f, _f, _p, _p, _p, _p]),
    "rs_project_lookup": (_i, [_p, _p, _i, _i, _i, _p, _p, _p]),
    "rs_tsdf_integrate": (_i, [_p, _p, _p, _i, _i, _f, _f, _f, _f, _p, _p, _f, _f, _f, _i, _i, _p, _p, _ll, _p, _i]
                          + [_p] * 6 + [_p]),
    "rs_feature_hidden_fwd": (_i, [_p] + [_i] * 7 + [_p, _p, _i, _p, _p, _p]),
    "rs_feature_branch": (_i, [_p, _i, _i, _i, _p, _p, _i, _p, _i, _i, _f] + [_p] * 6 + [_p]),
    "rs_feature_hidden_bwd": (_i, [_p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p]),
    "rs_rade_outputs_fwd": (_i, [_p] * 6 + [_f, _f] + [_i] * 5 + [_p] * 7 + [_p]),
    "rs_rade_outputs_bwd": (_i, [_p] * 6 + [_f, _f] + [_i] * 5 + [_p] * 12 + [_p]),
    "rs_rade_loss_fwd_bwd": (_i, [_p] * 7 + [_f, _f, _i, _i, _i, _f, _f, _f, _i] + [_p] * 6 + [_p]),
}

# per-call compositing options (include/rade_b200.h)
RS_RASTER_CULL_BBOX = 0x1
RS_RASTER_ONE_PIXEL = 0x2
RS_RASTER_NO_COLOR_MMA = 0x4
RS_RASTER_BWD_MMA = 0x8
RS_RASTER_FWD_RING = 0x10
RS_RASTER_BWD_BARRIER = 0x20


def RS_RASTER_BWD_TUNE(x: int) -> int:
    return (x & 0xF) << 8


EXPORTED_SYMBOLS = tuple(_SIGNATURES) + ("rs_set_last_cuda_error", "rs_count_launches", "rs_timing_begin",
                                          "rs_timing_end")


def timing_collect(cap: int = 64):
    """{entry point: (total ms, calls)} accumulated since rs_timing_enable(1) / the last collect."""
    lib = load()
    names = C.create_string_buffer(48 * cap)
    ms = (C.c_float * cap)()
    calls = (C.c_int * cap)()
    n = lib.rs_timing_collect(names, C.cast(ms, C.c_void_p), C.cast(calls, C.c_void_p), cap)
    out = {}
    for i in range(n):
        out[names.raw[48 * i:48 * (i + 1)].split(b"\0", 1)[0].decode()] = (float(ms[i]), int(calls[i]))
    return out


def lib_path() -> Path:
    return _LIB_PATH


def load():
    """Load the shared library (once).  Raises if it is missing -- never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise RuntimeError(
            f"{_LIB_PATH} not found: the CUDA extension has not been built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` at the repo root. "
            "There is no CPU fallback for this path.")
    lib = C.CDLL(str(_LIB_PATH))
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, what: str) -> int:
    if status < 0:
        lib = load()
        msg = lib.rs_error_string(status).decode()
        raise RuntimeError(f"librade_b200 {what} failed: status {status}: {msg}")
    return status


def ptr(t):
    """Device pointer of a tensor (None -> NULL).  The tensor must be a dense CUDA tensor."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("librade_b200 needs CUDA tensors: there is no CPU path")
    if not t.is_contiguous():
        raise RuntimeError("librade_b200 needs contiguous tensors")
    return C.c_void_p(t.data_ptr())


def zeros(shape, device, dtype=torch.float32):
    """torch.zeros through cudaMemsetAsync on the current stream (rs_zero_bytes) instead of a fill kernel."""
    t = torch.empty(shape, device=device, dtype=dtype)
    if t.numel():
        with torch.cuda.device(t.device):
            check(load().rs_zero_bytes(C.c_void_p(t.data_ptr()), t.numel() * t.element_size(), stream_ptr(t.device)),
                  "rs_zero_bytes")
    return t


def scale_unless_one(tensors, scale: torch.Tensor):
    """In place t *= scale (a device scalar) for up to 8 fp32 tensors in one launch that exits at once when
    scale == 1 (tested on the device)."""
    ts = [t for t in tensors if t is not None and t.numel()]
    if not ts:
        return
    assert len(ts) <= 8 and all(t.dtype == torch.float32 and t.is_contiguous() for t in ts)
    bufs = (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
    counts = (C.c_longlong * len(ts))(*[t.numel() for t in ts])
    s = scale.detach().to(torch.float32).reshape(1)
    with torch.cuda.device(ts[0].device):
        check(load().rs_scale_unless_one(C.cast(bufs, C.c_void_p), C.cast(counts, C.c_void_p), len(ts), ptr(s),
                                         stream_ptr(ts[0].device)), "rs_scale_unless_one")


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
