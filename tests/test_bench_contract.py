"""The bench lines committed under profiles/ (written by bench.py on a B200) carry every key the measurement contract
names (task statement, section 4): a guard against renaming a key in bench.py without noticing (CPU test; it reads the
committed evidence, it does not run the bench)."""

import json
from pathlib import Path

import pytest

PROFILES = Path(__file__).resolve().parent.parent / "profiles"


def _line(name):
    path = PROFILES / name
    if not path.exists():
        pytest.skip(f"{name} not committed")
    return json.loads(path.read_text().strip().splitlines()[-1])


def test_own_arm_line_has_the_contract_keys():
    d = _line("r02_bench_n1.json")
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in d, key
    assert d["n_gpus"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0
    assert abs(d["value"] - 1e3 / d["ms_per_step"]) < 1e-6 * d["value"]              # one view per step and GPU
    for key in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert key in d["e2e"], key
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["e2e"]["value"] != d["value"]                                           # measured separately
    r = d["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in r, key
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 2e-3
    c = d["cpu_baseline"]
    for key in ("value", "unit", "cores", "kind", "sample"):
        assert key in c, key
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1
    for key in ("sm_mhz", "sm_max_mhz", "reasons"):
        assert key in d["clocks"], key
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_reference_arm_line_has_the_contract_keys():
    d = _line("r02_bench_reference_arm.json")
    assert d["impl"] == "reference"
    own = _line("r02_bench_n1.json")
    for key in ("metric", "unit", "higher_is_better"):
        assert d[key] == own[key], key
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] in ("reference", "port")


@pytest.mark.parametrize("name,n", [("r02_bench_n2.json", 2), ("r02_bench_n4.json", 4), ("r02_bench_n8_final.json", 8)])
def test_multi_gpu_lines(name, n):
    d = _line(name)
    assert d["n_gpus"] == n and d["scaling"] == "weak"
    assert abs(d["value"] - n * 1e3 / d["ms_per_step"]) < 1e-6 * d["value"]          # whole-job aggregate
    assert d["multi_gpu"]["replica_gradients_bit_identical"] is True
    assert d["config4"]["replica_gradients_bit_identical"] is True and d["config4"]["scaling"] == "strong"
