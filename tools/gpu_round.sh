#!/bin/bash
# one GPU session: tests, smoke, kernel sweep, bench (both arms), ncu launch list + full capture of the hot kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv | tee gpurun_out/gpu.txt
nproc | tee -a gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.txt
timeout 600 python tools/kernel_sweep.py 2>&1 | tail -8 | tee gpurun_out/sweep.txt
timeout 900 python bench.py --steps 10 --warmup 3 2> gpurun_out/bench_err.txt | tee gpurun_out/bench.json
tail -5 gpurun_out/bench_err.txt
if [ "$1" != "quick" ]; then
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 2> gpurun_out/bench_ref_err.txt | tee gpurun_out/bench_ref.json
tail -5 gpurun_out/bench_ref_err.txt
fi
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"k_force_vel|k_bb_tail|k_pairs|k_inc_emit" -c 10 -f -o gpurun_out/prof python tools/profile_step.py > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out
