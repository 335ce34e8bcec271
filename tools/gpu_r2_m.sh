#!/bin/bash
# round 2, call M (2 GPUs): migration tests (python harness + C++ mirror on several ranks)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_dropin.py -m gpu -q --durations=5 -k "migration or time_stepping or mirror" 2>&1 | tail -80 > gpurun_out/r2m_pytest.txt
tail -12 gpurun_out/r2m_pytest.txt
ALENS_TEST_DEVICES=0,1 timeout 300 tests/cpp/test_multirank 2 5 2>&1 | tail -3
