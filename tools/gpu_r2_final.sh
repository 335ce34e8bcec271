#!/bin/bash
# round 2, measurement set kept under profiles/ (1 GPU): bench (both arms), S2 workload, example configurations,
# ncu launch list of one step + full capture of the BBPGD kernels and the pair search
cd "$(dirname "$0")/.."
T=${1:-v2}
O=gpurun_out/r2final_$T
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/gpu.txt; nproc >> $O/gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 > $O/smoke.txt; tail -1 $O/smoke.txt
timeout 900 python bench.py --steps 10 --warmup 3 2> $O/bench_n1_err.txt | tail -1 > $O/bench_n1.json; cut -c1-250 $O/bench_n1.json
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu --stamps 2> $O/bench_n1_stamps_err.txt | tail -1 > $O/bench_n1_stamps.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2> $O/bench_reference_err.txt | tail -1 > $O/bench_reference.json; cut -c1-250 $O/bench_reference.json
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --workload S2 2> $O/bench_s2_err.txt | tail -1 > $O/bench_s2.json; cut -c1-200 $O/bench_s2.json
timeout 900 python tools/bench_examples.py 50 > $O/examples.jsonl 2> $O/examples_err.txt; cut -c1-400 $O/examples.jsonl; tail -3 $O/examples_err.txt
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/launches_step_1M.csv python tools/profile_step.py > $O/ncu_launch.log 2>&1; tail -1 $O/ncu_launch.log
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"k_force_vel_rec|k_bb_tail" -c 8 -f -o $O/prof_bbpgd python tools/profile_step.py > $O/ncu_full.log 2>&1; tail -1 $O/ncu_full.log
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"k_pairs_find|k_cand_narrow|k_pairs_emit2|k_inc_emit_rec|k_setup|k_cell_order" -c 6 -f -o $O/prof_collect python tools/profile_step.py > $O/ncu_full2.log 2>&1; tail -1 $O/ncu_full2.log
python tools/ncu_summary.py $O/prof_bbpgd.ncu-rep $O/ncu_full_bbpgd.csv
python tools/ncu_summary.py $O/prof_collect.ncu-rep $O/ncu_full_collect.csv
ls -la $O | tail -25
