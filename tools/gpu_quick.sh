#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
timeout 600 python bench.py --no-cpu --e2e-steps 1 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
cat gpurun_out/bench_quick.json; tail -3 gpurun_out/bench_quick.err
