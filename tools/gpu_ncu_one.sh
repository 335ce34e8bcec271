#!/bin/bash
# ncu --set full capture of the kernels matching $1 (regex) in one resident bench step -> gpurun_out/$2.ncu-rep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"$1" -c ${3:-2} -f -o gpurun_out/$2 python tools/profile_step.py > gpurun_out/ncu_$2.log 2>&1
tail -2 gpurun_out/ncu_$2.log
