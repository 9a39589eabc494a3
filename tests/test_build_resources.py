"""Occupancy guard for the hot kernels (CPU, reads the objects the build left in collab-splats_b200/lib/obj).

The compositing kernels are issue-bound and tuned to a resident-CTA count: the forward to 72 registers (7 CTAs of 128
threads per SM), the ring backward to 90 (5 CTAs), the radix scatter pass to 64 (4 CTAs of 256).  A harmless-looking edit
can move the allocation (round 2: a run-time test in the forward's prologue pushed it to 80 registers, -1 CTA per SM,
+5 % time, found only in an ncu capture).  This test makes that a build failure instead.
"""

import re
import shutil
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "collab-splats_b200"))

LIMITS = [
    # object, demangled-name pattern (regex on the mangled symbol), max registers, what it buys
    ("rasterize.o", r"rasterize_fwd2_kernelILi4ELi128ELb0ELi4ELi2ELb[01]E", 72, "7 CTAs of 128 threads per SM"),
    ("rasterize.o", r"rasterize_bwd2_kernelILi128ELb0ELi5ELi4ELi3ELb[01]E", 90, "5 CTAs of 128 threads per SM"),
    ("radix_sort.o", r"radix_scatter_kernelI[yj]Li8ELi8EE", 64, "4 CTAs of 256 threads per SM"),
    ("projection.o", r"project_bwd_kernel", 128, "2 CTAs of 256 threads per SM"),
]


def _resource_usage(obj: Path):
    out = subprocess.run(["cuobjdump", "-res-usage", str(obj)], capture_output=True, text=True, check=True).stdout
    usage, name = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            name = m.group(1)
            continue
        m = re.search(r"REG:(\d+)\s+STACK:(\d+)", line)
        if m and name:
            usage[name] = (int(m.group(1)), int(m.group(2)))
            name = None
    return usage


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_hot_kernels_keep_their_register_budget():
    from radegs_b200 import build
    build.build()
    cache = {}
    for obj, pattern, max_regs, why in LIMITS:
        path = build.OBJ_DIR / obj
        assert path.exists(), f"{path} missing: build() did not leave its objects"
        usage = cache.setdefault(obj, _resource_usage(path))
        hits = {k: v for k, v in usage.items() if re.search(pattern, k)}
        assert hits, f"no kernel matching {pattern} in {obj}: renamed? update tests/test_build_resources.py"
        for name, (regs, stack) in hits.items():
            assert regs <= max_regs, f"{name}: {regs} registers > {max_regs} ({why})"
            assert stack <= 64, f"{name}: {stack} B of stack (spills) in a hot kernel"
