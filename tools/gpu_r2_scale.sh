#!/bin/bash
# round 2, scaling set (8 GPUs): N = 2, 4, 8 weak (1M rods per GPU) and N = 2, 4, 8 strong (1M rods in total), with stamps
cd "$(dirname "$0")/.."
O=gpurun_out/r2scale
mkdir -p $O
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > $O/gpus.txt
run() { # name, gpus, extra args
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 2960$2 bench.py --gpus $2 --steps 5 --warmup 3 --stamps $3 2> $O/$1_err.txt | tail -1 > $O/$1.json
}
run n8 8 ""
run n4 4 ""
run n2 2 ""
run strong_n8 8 "--scaling strong"
run strong_n4 4 "--scaling strong"
run strong_n2 2 "--scaling strong"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --stamps 2> $O/n1_err.txt | tail -1 > $O/n1.json
python - <<'PY'
import json
for n in ("n1","n2","n4","n8","strong_n2","strong_n4","strong_n8"):
    try:
        d=json.load(open(f"gpurun_out/r2scale/{n}.json")); b=d.get("iteration_breakdown_us") or {}
        print(n, "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "iters", d["config"]["bbpgd_iterations"], {k:v["rank0"] for k,v in b.items()}, (d.get("parity") or {}).get("status"))
    except Exception as e: print(n, "ERR", e)
PY
