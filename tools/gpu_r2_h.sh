#!/bin/bash
# round 2, call H (1 GPU): is U in cudaMalloc'ed / IPC-exported memory slower?  + full GPU test suite + S2 / K=20 / small configs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for o in "u_window=0" "u_window=1"; do
  ALENS_OPTIONS="$o" timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu --stamps 2> gpurun_out/r2h_err.txt | tail -1 > gpurun_out/r2h_$o.json
  python -c "
import json; d=json.load(open('gpurun_out/r2h_$o.json')); print('$o', d['ms_per_step'], {k:v['rank0'] for k,v in d['iteration_breakdown_us'].items()})"
done
timeout 1500 python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -30 | tee gpurun_out/r2h_pytest.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --workload S2 2> gpurun_out/r2h_s2_err.txt | tail -1 > gpurun_out/r2h_s2.json; cut -c1-400 gpurun_out/r2h_s2.json; tail -3 gpurun_out/r2h_s2_err.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --relax 20 2> gpurun_out/r2h_k20_err.txt | tail -1 > gpurun_out/r2h_k20.json; cut -c1-300 gpurun_out/r2h_k20.json
