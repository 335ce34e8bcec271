#!/bin/bash
# round 2, call W (2 GPUs): mirrored rows pushed by the tail kernel instead of the force kernel: tests + N=2 A/B with stamps
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_dropin.py -m gpu -q -x 2>&1 | tail -4
for o in "tail_push=1" "tail_push=0"; do
  ALENS_OPTIONS="$o" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus 2 --steps 4 --warmup 3 --stamps 2> gpurun_out/r2w_err.txt | tail -1 > "gpurun_out/r2w_$o.json"
  python - "$o" <<'PY'
import json,sys
o=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/r2w_{o}.json")); b=d.get("iteration_breakdown_us") or {}
    print(o, "value", d["value"], "ms", d["ms_per_step"], "iters", d["config"]["bbpgd_iterations"], {k:v["rank0"] for k,v in b.items()}, (d.get("parity") or {}).get("status"))
except Exception as e: print(o, "ERR", e); print(open("gpurun_out/r2w_err.txt").read()[-1500:])
PY
done
