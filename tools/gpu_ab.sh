#!/bin/bash
# A/B timing on one box: every library under lib/ and lib/exp/, interleaved, 3 rounds (kernel ms from bench.py's own events).
# Usage: tools/gpu_ab.sh [window=4096] [hop=1024]
W=${1:-4096}; H=${2:-1024}
for r in 1 2 3; do
for so in feature-extractor_b200/lib/libfxb200.so feature-extractor_b200/lib/exp/*.so; do
  n=$(basename $so .so)
  FXB200_LIB=$PWD/$so timeout 300 python bench.py --no-cpu --no-e2e --no-c5 --no-rt --steps 4 --warmup 3 --window $W --hop $H 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$n', $W, round(d['roofline']['kernel_ms'], 2), round(d['ms_per_step'], 2))"
done; done
