#!/bin/bash
# round 2, call L (2 GPUs): rod migration tests; what makes the force kernel 11 us slower at N > 1 (halo_debug timing runs)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_multirank.py -m gpu -q --durations=5 -k "migration or time_stepping" 2>&1 | tail -80 > gpurun_out/r2l_pytest.txt
tail -6 gpurun_out/r2l_pytest.txt
for dbg in 1 3; do
  ALENS_LATE_OPTIONS="halo_debug=$dbg" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$dbg bench.py --gpus 2 --steps 3 --warmup 3 --stamps --no-parity 2> gpurun_out/r2l_dbg${dbg}_err.txt | tail -1 > gpurun_out/r2l_dbg$dbg.json
done
python - <<'PY'
import json
for n in ("dbg1","dbg3"):
    try:
        d=json.load(open(f"gpurun_out/r2l_{n}.json")); b=d.get("iteration_breakdown_us") or {}
        print(n, "value", d["value"], "ms", d["ms_per_step"], "iters", d["config"]["bbpgd_iterations"], {k:v["rank0"] for k,v in b.items()})
    except Exception as e: print(n, "ERR", e)
PY
