mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -30 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
bash tools/gpu_ab.sh 4096 1024 2>&1 | tee gpurun_out/ab_4096.txt
bash tools/gpu_ab.sh 2048 1024 2>&1 | tee gpurun_out/ab_2048.txt
bash tools/gpu_prof.sh r02a
python tools/ncu_phases.py gpurun_out/prof_r02a.ncu-rep > gpurun_out/phases_r02a.txt 2>&1; head -30 gpurun_out/phases_r02a.txt
