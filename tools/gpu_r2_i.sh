#!/bin/bash
# round 2, call I (1 GPU): the tests that failed in call H with full tracebacks + the two-species search
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bcqp.py tests/test_gpu_dropin.py tests/test_gpu_mix.py -m gpu -q --durations=5 2>&1 | tail -150 > gpurun_out/r2i_pytest.txt
tail -15 gpurun_out/r2i_pytest.txt
