#!/bin/bash
# round 2, call C: record kernel variants A/B (rec_mode 0 / 1 / old act kernel), new bench.py, ncu of the BBPGD kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_solver.py tests/test_gpu_reference.py tests/test_gpu_multirank.py -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/r2c_pytest.txt
for mode in 1 0; do
  ALENS_OPTIONS="rec_mode=$mode" timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu 2> gpurun_out/r2c_bench_mode${mode}_err.txt | tee gpurun_out/r2c_bench_mode${mode}.json | cut -c1-300
done
ALENS_OPTIONS="force_kernel=1" timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu 2> gpurun_out/r2c_bench_act_err.txt | tee gpurun_out/r2c_bench_act.json | cut -c1-300
timeout 900 python bench.py --steps 5 --warmup 3 2> gpurun_out/r2c_bench_full_err.txt | tee gpurun_out/r2c_bench_full.json | cut -c1-200
tail -3 gpurun_out/r2c_bench_full_err.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2> gpurun_out/r2c_bench_ref_err.txt | tee gpurun_out/r2c_bench_ref.json | cut -c1-300
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"k_force_vel_rec|k_bb_tail" -c 8 -f -o gpurun_out/r2c_prof python tools/profile_step.py > gpurun_out/r2c_ncu.log 2>&1
tail -2 gpurun_out/r2c_ncu.log
ls -la gpurun_out | tail -12
