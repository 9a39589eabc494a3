"""Builds librade_b200.so (all CUDA kernels + the C ABI) in-tree with nvcc for sm_100a.

No torch, no pybind: the library is plain ``extern "C"`` and is loaded with ctypes
(``radegs_b200.backend``).  Each .cu is compiled to an object in parallel, then linked.
"""

from __future__ import annotations

import hashlib
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_ROOT = Path(__file__).resolve().parent.parent            # collab-splats_b200/
CSRC = PKG_ROOT / "csrc"
LIB_DIR = PKG_ROOT / "lib"
LIB_PATH = LIB_DIR / "librade_b200.so"
OBJ_DIR = LIB_DIR / "obj"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    cand = os.environ.get("NVCC") or "/usr/local/cuda/bin/nvcc"
    return cand if os.path.exists(cand) else "nvcc"


def _sources():
    return sorted(CSRC.glob("*.cu"))


def _signature() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    stamp = LIB_DIR / "build.sig"
    return LIB_PATH.exists() and stamp.exists() and stamp.read_text().strip() == _signature()


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and is_current():
        return LIB_PATH
    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src: Path) -> Path:
        obj = OBJ_DIR / (src.stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-I", str(CSRC), "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    # --cudart shared: the library resolves the CUDA runtime torch has already loaded (libcudart.so.12) instead of
    # embedding a static copy of every runtime entry point
    cmd = [nvcc, "-shared", "--cudart", "shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
           *[str(o) for o in objs], "-o", str(LIB_PATH)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    (LIB_DIR / "build.sig").write_text(_signature())
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=False))
