#!/bin/bash
# round 2, call B: the record-based force kernel -- correctness first, then the bench line and the old kernel for comparison
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -40 | tee gpurun_out/r2b_pytest_gpu.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu 2> gpurun_out/r2b_bench_err.txt | tee gpurun_out/r2b_bench.json
tail -3 gpurun_out/r2b_bench_err.txt
