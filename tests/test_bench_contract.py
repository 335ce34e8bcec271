"""The bench lines kept under profiles/ carry every key the measurement contract asks for (CPU-only check of the
committed JSON produced by bench.py on the B200 box)."""
import json
import os

PROF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")
BASE = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "e2e")


def _line(name):
    with open(os.path.join(PROF, name)) as f:
        return json.loads(f.readline())


def test_own_arm_line():
    d = _line("r1_bench_v6.json")
    for k in BASE + ("gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["unit"] == "steps/s" and d["higher_is_better"] is True and d["dtype"] == "f64"
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
    assert abs(d["value"] - 1e3 / d["ms_per_step"]) / d["value"] < 1e-3
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic", "kernel"):
        assert k in r, k
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    c = d["cpu_baseline"]
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in c, k
    assert c["kind"] in ("reference", "port") and c["value"] < d["value"]
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_reference_arm_line():
    d = _line("r1_bench_reference_v6.json")
    assert d["impl"] == "reference"
    for k in BASE + ("cpu_baseline",):
        assert k in d, k
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"] == d["cpu_baseline"]["value"]
    own = _line("r1_bench_v6.json")
    assert d["metric"] == own["metric"] and d["unit"] == own["unit"] and d["config"]["workload"] == own["config"]["workload"]


def test_scaling_lines():
    vals = {}
    for n in (2, 4, 8):
        d = _line(f"r1_bench_n{n}_v6.json")
        assert d["n_gpus"] == n and d["scaling"] == "weak"
        vals[n] = d["value"]
    assert vals[2] < vals[4] < vals[8]
