#!/bin/bash
# round 2, call F (4 GPUs): where the multi-GPU iteration time goes (stamps), N = 1 and N = 4
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu --stamps 2> gpurun_out/r2f_n1_err.txt | tail -1 > gpurun_out/r2f_n1.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 4 --warmup 3 --no-parity --stamps 2> gpurun_out/r2f_n4_err.txt | tail -1 > gpurun_out/r2f_n4.json
ALENS_OPTIONS="late_halo=0" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 4 --warmup 3 --no-parity --stamps 2> gpurun_out/r2f_n4_nolate_err.txt | tail -1 > gpurun_out/r2f_n4_nolate.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 4 --warmup 3 --no-parity --stamps 2> gpurun_out/r2f_n2_err.txt | tail -1 > gpurun_out/r2f_n2.json
python - <<'PY'
import json
for n in ("n1","n2","n4","n4_nolate"):
    try:
        d=json.load(open(f"gpurun_out/r2f_{n}.json")); print(n, d["ms_per_step"], json.dumps(d.get("iteration_breakdown_us")))
    except Exception as e: print(n, "ERR", e)
PY
