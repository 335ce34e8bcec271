#!/bin/bash
# round 2, call P (1 GPU): full GPU suite + smoke with rec_mode 2 as the default
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -40 > gpurun_out/r2p_pytest.txt; tail -12 gpurun_out/r2p_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
