mkdir -p gpurun_out
bash tools/gpu_ab.sh 4096 1024 2>&1 | tee gpurun_out/ab_4096.txt
bash tools/gpu_ab.sh 2048 1024 2>&1 | tee gpurun_out/ab_2048.txt
bash tools/gpu_ab.sh 1024 512 2>&1 | tee gpurun_out/ab_1024.txt
