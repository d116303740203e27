#!/bin/bash
# One GPU-box session: parity tests, real-time latency harness, bench (ours + reference arm), optional ncu launch list.
# Usage: tools/gpu_round.sh [rt_seconds=60] [ncu]
RT=${1:-60}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt; lscpu | grep -E "Model name|NUMA|Socket" >> gpurun_out/smi.txt
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
g++ -O2 -std=c++17 tools/rt_latency.cpp -Iinclude -Lfeature-extractor_b200/lib -lfxb200 -lpthread -Wl,-rpath,$PWD/feature-extractor_b200/lib -o /tmp/rt_latency
timeout 300 /tmp/rt_latency 512 256 $RT 1 128 2048 > gpurun_out/rt_paced.json 2> gpurun_out/rt.err
timeout 300 /tmp/rt_latency 512 256 10 0 128 2048 > gpurun_out/rt_unpaced.json 2>> gpurun_out/rt.err
cat gpurun_out/rt_paced.json gpurun_out/rt_unpaced.json; tail -3 gpurun_out/rt.err
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
cat gpurun_out/bench_ref.json
if [ "$2" == "ncu" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_list.log 2>&1
fi
