#!/bin/bash
# round 2, call T (2 GPUs): where do the force kernel's +17 us at N = 2 come from (remote stores off / fence off)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for dbg in 0 1 3; do
  ALENS_LATE_OPTIONS="halo_debug=$dbg" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2957$dbg bench.py --gpus 2 --steps 3 --warmup 3 --stamps --no-parity 2> gpurun_out/r2t_dbg${dbg}_err.txt | tail -1 > gpurun_out/r2t_dbg$dbg.json
done
python - <<'PY'
import json
for n in ("dbg0","dbg1","dbg3"):
    try:
        d=json.load(open(f"gpurun_out/r2t_{n}.json")); b=d.get("iteration_breakdown_us") or {}
        print(n, "value", d["value"], "ms", d["ms_per_step"], "iters", d["config"]["bbpgd_iterations"], {k:v["rank0"] for k,v in b.items()}, d["config"].get("ghosts_rank0"))
    except Exception as e: print(n, "ERR", e)
PY
