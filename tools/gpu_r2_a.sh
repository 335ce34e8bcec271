#!/bin/bash
# round 2, call A: GPU tests (incl. the new reference-code parity tests), smoke, short bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | tee gpurun_out/r2_gpu.txt
nproc | tee -a gpurun_out/r2_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -x --durations=15 2>&1 | tail -40 | tee gpurun_out/r2_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r2_smoke.txt
timeout 600 python bench.py --steps 5 --warmup 3 2> gpurun_out/r2_bench_err.txt | tee gpurun_out/r2_bench.json
tail -5 gpurun_out/r2_bench_err.txt
