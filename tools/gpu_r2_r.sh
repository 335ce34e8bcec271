#!/bin/bash
# round 2, call R (2 GPUs): zero-row store skipping: full tests (fused multi-rank kernels included), N=1 and N=2 stamps
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x --durations=3 2>&1 | tail -30 > gpurun_out/r2r_pytest.txt; tail -6 gpurun_out/r2r_pytest.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --stamps 2> gpurun_out/r2r_n1_err.txt | tail -1 > gpurun_out/r2r_n1.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 4 --warmup 3 --stamps 2> gpurun_out/r2r_n2_err.txt | tail -1 > gpurun_out/r2r_n2.json
python - <<'PY'
import json
for n in ("n1","n2"):
    try:
        d=json.load(open(f"gpurun_out/r2r_{n}.json")); b=d.get("iteration_breakdown_us") or {}
        print(n, "value", d["value"], "ms", d["ms_per_step"], "iters", d["config"]["bbpgd_iterations"], {k:v["rank0"] for k,v in b.items()}, (d.get("parity") or {}).get("status"), d["config"]["phase_ms_per_step"])
    except Exception as e: print(n, "ERR", e)
PY
