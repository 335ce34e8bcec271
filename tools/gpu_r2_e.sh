#!/bin/bash
# round 2, call E (4 GPUs): boundary-first force kernel + ghost-rows-last tail under pytest (one rank per GPU), bench N=2/4 weak + strong
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q 2>&1 | tail -12 | tee gpurun_out/r2e_pytest.txt
for n in 2 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 5 --warmup 3 2> gpurun_out/r2e_bench_n${n}_err.txt | tail -1 > gpurun_out/r2e_bench_n${n}.json
  cut -c1-200 gpurun_out/r2e_bench_n${n}.json
done
ALENS_OPTIONS="late_halo=0" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 5 --warmup 3 --no-parity 2> gpurun_out/r2e_bench_n4_nolate_err.txt | tail -1 > gpurun_out/r2e_bench_n4_nolate.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 5 --warmup 3 --scaling strong 2> gpurun_out/r2e_bench_strong_n4_err.txt | tail -1 > gpurun_out/r2e_bench_strong_n4.json
cut -c1-200 gpurun_out/r2e_bench_strong_n4.json
