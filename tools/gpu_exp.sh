#!/bin/bash
# parity on the in-tree build, then a short device-only bench of every experiment variant under lib/exp/
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
for so in feature-extractor_b200/lib/libfxb200.so feature-extractor_b200/lib/exp/*.so; do
  n=$(basename $so .so)
  FXB200_LIB=$PWD/$so timeout 300 python bench.py --no-cpu --no-e2e --steps 3 --warmup 2 > gpurun_out/exp_$n.json 2> gpurun_out/exp_$n.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/exp_$n.json"))
    print("$n", "ms/step %.2f" % d["ms_per_step"], "frames/s %.3e" % d["value"], "kernel_ms %.2f" % d["roofline"]["kernel_ms"])
except Exception as e:
    print("$n", "failed", e)
PY
done
