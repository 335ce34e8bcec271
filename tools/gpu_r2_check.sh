#!/bin/bash
# what the driver runs at round end, on one GPU: the GPU test suite, smoke(), the default bench line and the reference arm
cd "$(dirname "$0")/.."
O=gpurun_out/r2check
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -25 > $O/pytest.txt; tail -8 $O/pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/smoke.txt
timeout 900 python bench.py 2> $O/bench_err.txt | tail -1 > $O/bench.json; cut -c1-300 $O/bench.json; tail -2 $O/bench_err.txt
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2check/bench.json"))
print({k:d[k] for k in ("value","ms_per_step","steps","warmup","gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"], d["roofline"].get("in_loop"), (d.get("parity") or {}).get("status"), d.get("iteration_breakdown_us",{}).get("iteration_us"))
PY
