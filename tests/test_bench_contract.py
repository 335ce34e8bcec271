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


def test_round2_lines():
    own = _line("r2_bench_n1.json")
    for k in BASE + ("gpu_launches", "clocks", "roofline", "cpu_baseline", "parity"):
        assert k in own, k
    assert own["n_gpus"] == 1 and own["warmup"] >= 3 and own["gpu_launches"] > 0 and own["vs_baseline"] is None
    assert own["parity"]["status"] == "ok" and own["parity"]["ref_pairs_missing_from_ours"] == 0
    assert own["parity"]["gamma_rel_err"] < 1e-8 and own["parity"]["velocity_rel_err"] < 1e-8
    e = own["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < own["value"]
    r = own["roofline"]
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3 and 0 < r["frac"] < 1
    loop = r["in_loop"]  # the two kernels of one BBPGD iteration, stamped inside the loop
    assert set(loop) >= {"k_force_vel_rec", "k_bb_tail"}
    for k in loop.values():
        assert 0 < k["frac"] < 1 and k["us"] > 0
    ref = _line("r2_bench_reference.json")
    assert ref["impl"] == "reference" and ref["metric"] == own["metric"] and ref["unit"] == own["unit"]
    assert ref["config"]["workload"] == own["config"]["workload"]
    assert ref["e2e"]["value"] == ref["value"] == ref["cpu_baseline"]["value"] and ref["value"] < e["value"]
    weak, strong = {1: own["value"]}, {}
    for n in (2, 4, 8):
        d = _line(f"r2_scale_n{n}.json")
        assert d["n_gpus"] == n and d["scaling"] == "weak" and d["parity"]["status"] == "ok"
        assert d["parity"]["rows"] == d["parity"]["rows_single_gpu"]
        weak[n] = d["value"]
        s = _line(f"r2_scale_strong_n{n}.json")
        assert s["n_gpus"] == n and s["scaling"] == "strong" and s["parity"]["status"] == "ok"
        strong[n] = s["value"]
    assert weak[1] < weak[2] < weak[4] < weak[8] and own["value"] < strong[2] < strong[4] < strong[8]
