#!/bin/bash
# round 2, call Q (1 GPU): rod headers (start + live bits in one load): tests + stamps
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x --durations=3 2>&1 | tail -30 > gpurun_out/r2q_pytest.txt; tail -8 gpurun_out/r2q_pytest.txt
for o in "rec_mode=2" "rec_mode=0"; do
  ALENS_OPTIONS="$o" timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --stamps 2> gpurun_out/r2q_err.txt | tail -1 > "gpurun_out/r2q_$o.json"
  python - "$o" <<'PY'
import json,sys
o=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/r2q_{o}.json")); b=d["iteration_breakdown_us"]; r=d["roofline"]["all_kernels"]
    print(o, "ms", d["ms_per_step"], "iters", d["config"]["bbpgd_iterations"], {k:v["rank0"] for k,v in b.items()}, {k:v["avg_us"] for k,v in r.items()}, d["config"]["phase_ms_per_step"])
except Exception as e: print(o, "ERR", e); print(open("gpurun_out/r2q_err.txt").read()[-800:])
PY
done
