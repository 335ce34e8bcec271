cd /root/repo
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.txt
timeout 900 python tools/kernel_sweep.py 2>&1 | tail -16 | tee gpurun_out/sweep.txt
