#!/bin/bash
# The round's closing GPU session: every -m gpu test, smoke, bench (ours + reference arm), other window sizes, the launch list,
# one ncu --set full capture of k_analyse with the per-phase breakdown.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt; lscpu | grep -E "Model name|NUMA|Socket" >> gpurun_out/smi.txt
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 900 python tools/parity_report.py --out gpurun_out/parity.json 2>&1 | tail -6
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
cut -c1-400 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
cut -c1-300 gpurun_out/bench_ref.json
for w in "2048 1024" "2048 512" "1024 512"; do set -- $w; python bench.py --no-cpu --no-e2e --no-c5 --no-rt --steps 5 --warmup 3 --window $1 --hop $2 2>/dev/null; done > gpurun_out/bench_other_windows.json
cut -c1-200 gpurun_out/bench_other_windows.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-c5 --no-rt > gpurun_out/ncu_list.log 2>&1
bash tools/gpu_prof.sh final
python tools/ncu_phases.py gpurun_out/prof_final.ncu-rep 1916928 16 0 > gpurun_out/phases_final.txt 2>&1; head -22 gpurun_out/phases_final.txt | cut -c1-200
python tools/summarize_ncu.py gpurun_out/prof_final.ncu-rep gpurun_out/ncu_full_final.txt; head -30 gpurun_out/ncu_full_final.txt
