mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
bash tools/gpu_ab.sh 4096 1024 2>&1 | tee gpurun_out/ab_4096.txt
bash tools/gpu_ab.sh 2048 1024 2>&1 | tee gpurun_out/ab_2048.txt
