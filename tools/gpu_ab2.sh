#!/bin/bash
# A/B timing like gpu_ab.sh with a chosen number of rounds and a library filter: tools/gpu_ab2.sh window hop rounds 'glob under lib/exp'
W=${1:-4096}; H=${2:-1024}; R=${3:-2}; G=${4:-*}
for r in $(seq 1 $R); do
for so in feature-extractor_b200/lib/libfxb200.so feature-extractor_b200/lib/exp/$G.so; do
  n=$(basename $so .so)
  FXB200_LIB=$PWD/$so timeout 300 python bench.py --no-cpu --no-e2e --no-c5 --no-rt --steps 4 --warmup 3 --window $W --hop $H 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$n', $W, round(d['roofline']['kernel_ms'], 2), round(d['ms_per_step'], 2))"
done; done
