#!/bin/bash
# ncu full capture of k_analyse (1 launch) -> gpurun_out/prof_$1.ncu-rep
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_analyse -s 3 -c 1 -f -o gpurun_out/prof_$1 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-c5 --no-rt > gpurun_out/ncu_full_$1.log 2>&1
tail -2 gpurun_out/ncu_full_$1.log
