#!/bin/bash
# round 2, call G (4 GPUs): z-slabs vs x-slabs with stamps
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for ax in 2 0; do
  ALENS_SLAB_AXIS=$ax timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2951$ax bench.py --gpus 4 --steps 4 --warmup 3 --no-parity --stamps 2> gpurun_out/r2g_n4_ax${ax}_err.txt | tail -1 > gpurun_out/r2g_n4_ax${ax}.json
done
ALENS_SLAB_AXIS=2 ALENS_OPTIONS="late_halo=0" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 4 --warmup 3 --no-parity --stamps 2> gpurun_out/r2g_n4_ax2_nolate_err.txt | tail -1 > gpurun_out/r2g_n4_ax2_nolate.json
ALENS_SLAB_AXIS=2 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 4 --warmup 3 --stamps 2> gpurun_out/r2g_n2_ax2_err.txt | tail -1 > gpurun_out/r2g_n2_ax2.json
ALENS_SLAB_AXIS=2 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 4 --steps 4 --warmup 3 --scaling strong --stamps 2> gpurun_out/r2g_n4_strong_err.txt | tail -1 > gpurun_out/r2g_n4_strong.json
python - <<'PY'
import json
for n in ("n4_ax2","n4_ax0","n4_ax2_nolate","n2_ax2","n4_strong"):
    try:
        d=json.load(open(f"gpurun_out/r2g_{n}.json")); b=d.get("iteration_breakdown_us") or {}
        print(n, "value", d["value"], "ms", d["ms_per_step"], "iters", d["config"]["bbpgd_iterations"], {k:v["rank0"] for k,v in b.items()}, (d.get("parity") or {}).get("status"))
    except Exception as e: print(n, "ERR", e)
PY
