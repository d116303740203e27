#!/bin/bash
# configs[4] leg (2 minutes of audio) against the number of chunks per track, then the headline leg with the heuristic's own choice
for c in 1 2 3 4 ""; do
  FXB200_CHUNKS=$c timeout 300 python bench.py --no-cpu --no-e2e --no-rt --steps 3 --warmup 2 --c5-minutes 2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('chunks [$c]', 'c5 frames/s', round(d['c5']['value']), 'kernel_ms', round(d['roofline']['kernel_ms'], 2), 'step', round(d['ms_per_step'], 2))"
done
