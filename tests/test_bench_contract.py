"""CPU: the bench lines committed under profiles/ carry every key the bench contract asks for, and their numbers are consistent with
each other (value = work / time, roofline.frac = achieved / peak, e2e slower than resident, reference arm has zero copy bytes)."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config")


def _line(name):
    p = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(p):
        pytest.skip(name + " not committed")
    return json.loads(open(p).read().strip().splitlines()[-1])


@pytest.mark.parametrize("name,n", [("r02_bench_1gpu.json", 1), ("r02_bench_2gpu.json", 2), ("r02_bench_8gpu.json", 8)])
def test_our_arm_lines(name, n):
    d = _line(name)
    for k in BASE + ("e2e", "gpu_launches", "roofline", "clocks"):
        assert k in d, k
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))["metric"]      # "fragments.EM-iters/sec and wall-time to converge, ..."
    assert d["n_gpus"] == n and d["higher_is_better"] is True and d["scaling"] == "weak" and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["metric"] == "fragments*EM-iters/sec" and base.startswith("fragments") and "EM-iters/sec" in base and "wall_ms_to_converge" in d
    assert d["vs_baseline"] is None and "workload" in d["config"] and d["warmup"] >= 3
    assert d["gpu_launches"] > 0
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["unit"] == "GB/s"
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if n == 1:
        c = d["cpu_baseline"]
        assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
        g = d["roofline_giant"]
        assert g["traffic"] and g["traffic"] < g["alg_bytes_per_launch"] and 0.5 < g["real_bytes_frac"] < g["frac"] < 1.0
    else:
        inv = d["strong"]["partition_invariance"]
        assert inv["arrays_differing_bitwise"] == 0


def test_reference_arm_line():
    d = _line("r02_bench_reference_arm.json")
    for k in BASE + ("impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    ours = _line("r02_bench_1gpu.json")
    assert d["metric"] == ours["metric"] and d["unit"] == ours["unit"] and ours["e2e"]["value"] / d["value"] > 50
