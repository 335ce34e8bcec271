#!/bin/bash
# round 2, call S (1 GPU): 256-bit stores in the record emission; HALO instantiation of the force kernel without neighbours
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_solver.py tests/test_gpu_reference.py -m gpu -q -x 2>&1 | tail -4
for o in "stamps=0" "halo_debug=4"; do
  ALENS_OPTIONS="$o" timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --stamps 2> gpurun_out/r2s_err.txt | tail -1 > "gpurun_out/r2s_$o.json"
  python - "$o" <<'PY'
import json,sys
o=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/r2s_{o}.json")); b=d["iteration_breakdown_us"]; r=d["roofline"]["all_kernels"]
    print(o, "ms", d["ms_per_step"], "iters", d["config"]["bbpgd_iterations"], {k:v["rank0"] for k,v in b.items()}, {k:v["avg_us"] for k,v in r.items()}, d["config"]["phase_ms_per_step"])
except Exception as e: print(o, "ERR", e); print(open("gpurun_out/r2s_err.txt").read()[-800:])
PY
done
