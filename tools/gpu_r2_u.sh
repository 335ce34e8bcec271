#!/bin/bash
# round 2, call U (2 GPUs): coalesced remote pushes: multi-rank tests + N=2 stamps
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_dropin.py -m gpu -q -x 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 2 --steps 4 --warmup 3 --stamps 2> gpurun_out/r2u_n2_err.txt | tail -1 > gpurun_out/r2u_n2.json
python - <<'PY'
import json
for n in ("n2",):
    try:
        d=json.load(open(f"gpurun_out/r2u_{n}.json")); b=d.get("iteration_breakdown_us") or {}
        print(n, "value", d["value"], "ms", d["ms_per_step"], "iters", d["config"]["bbpgd_iterations"], {k:v["rank0"] for k,v in b.items()}, (d.get("parity") or {}).get("status"))
    except Exception as e: print(n, "ERR", e)
PY
