"""The C-ABI library loads on a CPU-only box and exports every symbol include/rade_b200.h declares; entry
points validate their arguments before touching a device; the Python layer refuses CPU tensors (there is
no CPU fallback).  No compute call is made here."""

import ctypes
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from radegs_b200 import backend, build
    build.build()           # no-op when the in-tree library is current
    return backend.load()


def _declared_symbols():
    text = (ROOT / "include" / "rade_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rs_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    from radegs_b200 import backend
    names = _declared_symbols()
    assert len(names) >= 20
    raw = ctypes.CDLL(str(backend.lib_path()))
    for n in names:
        assert hasattr(raw, n), f"{n} is declared in include/rade_b200.h but not exported"
    assert set(backend.EXPORTED_SYMBOLS) == set(names), set(backend.EXPORTED_SYMBOLS) ^ set(names)


def test_pure_host_entry_points(lib):
    assert lib.rs_version() >= 100
    assert lib.rs_error_string(0) == b"ok"
    assert b"bad argument" in lib.rs_error_string(-1)
    assert lib.rs_tile_bits(120, 68) == (120 * 68).bit_length() == 13
    assert lib.rs_tile_bits(16, 16) == 9 and lib.rs_tile_bits(1, 1) == 1
    assert [lib.rs_raster_padded_channels(d) for d in (1, 3, 4, 5, 16, 17, 67, 68, 72, 73)] == \
        [4, 4, 4, 8, 16, 20, 68, 68, 72, -1]
    assert lib.rs_cumsum_temp_bytes(1) >= 24 and lib.rs_cumsum_temp_bytes(10_000_000) > 8 * 4000
    a, b = lib.rs_sort_pairs_temp_bytes(1_000_000, 0, 46), lib.rs_sort_pairs_temp_bytes(1_000_000, 0, 40)
    assert a > b > 0


def test_argument_validation_before_any_launch(lib):
    null = None
    # negative sizes / null pointers are rejected with RS_ERR_BAD_ARG, nothing is launched
    assert lib.rs_project_fwd(null, null, null, null, null, 1, -1, 64, 64, 0.3, 0.01, 1e10, 0.0, 0, null, null, null,
                              null, null, null, null, null, null) == -1
    assert lib.rs_project_fwd(null, null, null, null, null, 1, 10, 64, 64, 0.3, 0.01, 1e10, 0.0, 0, null, null, null,
                              null, null, null, null, null, null) == -1
    assert lib.rs_sh_fwd(4, 25, 10, 10, null, null, null, null, null) == -1          # degree > 3
    assert lib.rs_sort_pairs(null, null, null, null, 10, 0, 46, null, 0, null) == -1
    assert lib.rs_sort_pairs(null, null, null, null, 0, 0, 46, null, 0, null) == 1    # empty input: nothing to do
    assert lib.rs_sort_pairs(null, null, null, null, 1 << 31, 0, 46, null, 0, null) == -3
    assert lib.rs_isect_count(null, null, 5, 4, 4, null, null) == -1
    assert lib.rs_rasterize_fwd(null, null, 0, 3, -1, null, null, 1, 1, 16, 16, 1, 1, null, null, 0, null, null, null,
                                null, null, null, null, null, 0, null, null, null) == -1
    assert lib.rs_sh_colors_fwd(4, 25, 1, 10, null, null, null, null, null, null, null) == -1   # degree > 3
    # zero-sized problems are fine
    assert lib.rs_project_fwd(null, null, null, null, null, 1, 0, 64, 64, 0.3, 0.01, 1e10, 0.0, 0, null, null, null,
                              null, null, null, null, null, null) == 0
    assert lib.rs_isect_count(null, null, 0, 4, 4, null, null) == 0


def test_python_layer_has_no_cpu_path():
    from gsplat import rasterization
    from gsplat.cuda._wrapper import fully_fused_projection, spherical_harmonics
    N = 8
    means, quats, scales = torch.zeros(N, 3), torch.ones(N, 4), torch.ones(N, 3)
    vm, Ks = torch.eye(4)[None], torch.eye(3)[None]
    with pytest.raises(RuntimeError, match="CUDA"):
        fully_fused_projection(means, None, quats, scales, vm, Ks, 32, 32)
    with pytest.raises(RuntimeError, match="CUDA"):
        spherical_harmonics(0, torch.ones(N, 3), torch.ones(N, 1, 3))
    with pytest.raises(RuntimeError, match="CUDA"):
        rasterization(means, quats, scales, torch.ones(N), torch.ones(N, 3), vm, Ks, 32, 32, packed=False)
    with pytest.raises(NotImplementedError):
        rasterization(means, quats, scales, torch.ones(N), torch.ones(N, 3), vm, Ks, 32, 32)   # packed=True default
    with pytest.raises(NotImplementedError):
        fully_fused_projection(means, torch.ones(N, 3, 3), None, None, vm, Ks, 32, 32)


def test_header_is_plain_c_and_cxx():
    """include/rade_b200.h is the drop-in boundary: it must compile as C99 and as C++ with no torch / CUDA headers."""
    import shutil
    import subprocess
    hdr = ROOT / "include" / "rade_b200.h"
    for cc, args in (("gcc", ["-std=c99", "-x", "c"]), ("g++", ["-std=c++17", "-x", "c++"])):
        if shutil.which(cc) is None:
            pytest.skip(f"{cc} not installed")
        r = subprocess.run([cc, *args, "-Wall", "-Werror", "-fsyntax-only", str(hdr)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    text = hdr.read_text()
    assert "torch" not in text.lower().replace("pytorch", "") or "#include <torch" not in text
    assert "#include <cuda" not in text and "at::" not in text


def test_every_entry_point_cites_the_reference():
    """Each block of the header names the reference interface it replaces (file:line) or the SURVEY row."""
    text = (ROOT / "include" / "rade_b200.h").read_text()
    assert text.count(".py:") >= 8, "reference file:line citations are missing from the header"
    for needle in ("rade_gs_model.py", "rade_features_model.py", "camera_utils.py", "mesh.py", "features.py"):
        assert needle in text, needle
