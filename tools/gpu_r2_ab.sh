#!/bin/bash
# A/B of the library's measured alternatives at the bench size (1 GPU), with per-iteration stamps:
#   tools/gpu_r2_ab.sh "rec_mode=2" "rec_mode=1" "rec_mode=0" "force_kernel=1"
# Multi-GPU timing experiments that break results (remote stores / fence off) use ALENS_LATE_OPTIONS="halo_debug=1|2|3".
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for o in "$@"; do
  ALENS_OPTIONS="$o" timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --stamps 2> gpurun_out/ab_err.txt | tail -1 > "gpurun_out/ab_$o.json"
  python - "$o" <<'PY'
import json, sys
o = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/ab_{o}.json")); b = d["iteration_breakdown_us"]; r = d["roofline"]["all_kernels"]
    print(o, "ms", d["ms_per_step"], "iters", d["config"]["bbpgd_iterations"], {k: v["rank0"] for k, v in b.items()},
          {k: v["avg_us"] for k, v in r.items()}, d["config"]["phase_ms_per_step"])
except Exception as e:
    print(o, "ERR", e); print(open("gpurun_out/ab_err.txt").read()[-800:])
PY
done
