#!/bin/bash
# polydisperse variant of S1 (log-normal lengths of the same mean): collect phase with the long-rod pass at several cell sizes
# (long_rods = percent of the mean bounding radius the cells are sized for; 0 = cells sized for the longest rod)
#   tools/gpu_r2_poly.sh "0.3 0.5" "200 150 130"
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for sg in ${1:-0.3 0.5}; do
for lr in ${2:-200 0}; do
ALENS_LENGTH_SIGMA=$sg ALENS_OPTIONS="long_rods=$lr" timeout 600 python bench.py --rods 300000 --steps 3 --warmup 3 --no-cpu --no-stamps --relax 2 2> gpurun_out/poly_err.txt | tail -1 > gpurun_out/poly.json
python - "$sg" "$lr" <<'PY'
import json,sys
try:
    d=json.load(open("gpurun_out/poly.json")); print("sigma",sys.argv[1],"long_rods",sys.argv[2], "ms", d["ms_per_step"], "nc", d["config"]["constraints"], "iters", d["config"]["bbpgd_iterations"], d["config"]["phase_ms_per_step"], d["roofline"].get("pair_search"))
except Exception as e: print("ERR", sys.argv[1:], e); print(open("gpurun_out/poly_err.txt").read()[-600:])
PY
done
done
