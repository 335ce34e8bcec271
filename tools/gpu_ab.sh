#!/bin/bash
# quick A/B session: solver + multirank tests, then the option sweep given in $ALENS_SWEEP (tools/kernel_sweep.py)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q ${ALENS_PYTEST_ARGS} 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.txt
timeout 900 python tools/kernel_sweep.py 2>&1 | tail -12 | tee gpurun_out/sweep.txt
