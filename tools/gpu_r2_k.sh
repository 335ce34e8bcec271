#!/bin/bash
# round 2, call K (4 GPUs): multi-rank tests on separate devices, then N=2 / N=4 weak and N=4 strong with stamps
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bcqp.py tests/test_gpu_multirank.py tests/test_gpu_solver.py -m gpu -q --durations=5 2>&1 | tail -60 > gpurun_out/r2k_pytest.txt
tail -4 gpurun_out/r2k_pytest.txt
run() { # name, gpus, extra args
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 2952$2 bench.py --gpus $2 --steps 4 --warmup 3 --stamps $3 2> gpurun_out/r2k_$1_err.txt | tail -1 > gpurun_out/r2k_$1.json
}
run n2 2 ""
run n4 4 ""
run n4_strong 4 "--scaling strong"
python - <<'PY'
import json
for n in ("n2","n4","n4_strong"):
    try:
        d=json.load(open(f"gpurun_out/r2k_{n}.json")); b=d.get("iteration_breakdown_us") or {}
        print(n, "value", d["value"], "ms", d["ms_per_step"], "iters", d["config"]["bbpgd_iterations"], {k:v["rank0"] for k,v in b.items()}, (d.get("parity") or {}).get("status"))
    except Exception as e: print(n, "ERR", e)
PY
