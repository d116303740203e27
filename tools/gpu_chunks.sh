#!/bin/bash
# kernel time of the bench workload against the number of chunks per track (tuning aid for choose_chunks)
for r in 1 2; do for c in 1 2 3 4 5 6 8 13; do
  FXB200_CHUNKS=$c timeout 300 python bench.py --no-cpu --no-e2e --no-c5 --no-rt --steps 4 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('chunks $c', round(d['roofline']['kernel_ms'], 2), round(d['ms_per_step'], 2))"
done; done
