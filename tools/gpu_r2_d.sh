#!/bin/bash
# round 2, call D (4 GPUs): fused multi-rank kernels under pytest (one rank per GPU), bench at N=1/2/4 weak + strong with parity flags
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/r2d_gpus.txt
timeout 1200 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_solver.py tests/test_gpu_bcqp.py tests/test_gpu_reference.py -m gpu -q --durations=6 2>&1 | tail -25 | tee gpurun_out/r2d_pytest.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu 2> gpurun_out/r2d_bench_n1_err.txt | tee gpurun_out/r2d_bench_n1.json | cut -c1-250
for n in 2 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 5 --warmup 3 2> gpurun_out/r2d_bench_n${n}_err.txt | tee gpurun_out/r2d_bench_n${n}.json | cut -c1-250
  tail -2 gpurun_out/r2d_bench_n${n}_err.txt
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 5 --warmup 3 --scaling strong 2> gpurun_out/r2d_bench_strong_n${n}_err.txt | tee gpurun_out/r2d_bench_strong_n${n}.json | cut -c1-250
  tail -2 gpurun_out/r2d_bench_strong_n${n}_err.txt
done
