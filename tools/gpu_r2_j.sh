#!/bin/bash
# round 2, call J (2 GPUs): fixed tests, then the N=2 weak run with stamps (ownership byte streamed, ghost slots skipped)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bcqp.py tests/test_gpu_dropin.py tests/test_gpu_mix.py tests/test_gpu_multirank.py -m gpu -q --durations=5 2>&1 | tail -60 > gpurun_out/r2j_pytest.txt
tail -5 gpurun_out/r2j_pytest.txt
ALENS_SLAB_AXIS=2 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 4 --warmup 3 --stamps 2> gpurun_out/r2j_n2_err.txt | tail -1 > gpurun_out/r2j_n2.json
python - <<'PY'
import json
for n in ("n2",):
    try:
        d=json.load(open(f"gpurun_out/r2j_{n}.json")); b=d.get("iteration_breakdown_us") or {}
        print(n, "value", d["value"], "ms", d["ms_per_step"], "iters", d["config"]["bbpgd_iterations"], {k:v["rank0"] for k,v in b.items()}, (d.get("parity") or {}).get("status"))
    except Exception as e: print(n, "ERR", e)
PY
