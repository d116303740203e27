#!/bin/bash
mkdir -p gpurun_out
python tools/gpu_debug2.py 2>&1 | tail -40
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_analyse -s 3 -c 1 -f -o gpurun_out/prof_k1_v2 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full_v2.log 2>&1
tail -2 gpurun_out/ncu_full_v2.log
