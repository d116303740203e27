#!/bin/bash
# K1b (k_finalize) launch time per library variant, from an ncu launch list (cold-cache, serialised): tools/gpu_fz.sh [window] [hop]
W=${1:-4096}; H=${2:-1024}
for so in feature-extractor_b200/lib/libfxb200.so feature-extractor_b200/lib/exp/*.so; do
  n=$(basename $so .so)
  FXB200_LIB=$PWD/$so ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file /tmp/l.csv python bench.py --no-cpu --no-e2e --no-c5 --no-rt --steps 2 --warmup 0 --window $W --hop $H > /dev/null 2>&1
  python - "$n" $W <<'P'
import csv,sys
rows=[r for r in csv.reader(open('/tmp/l.csv')) if len(r)>10 and r[0].isdigit()]
fz=[int(r[-1]) for r in rows if 'k_finalize' in r[4]]; ka=[int(r[-1]) for r in rows if 'k_analyse' in r[4]]
print(sys.argv[1], sys.argv[2], 'k_finalize us', [round(x/1e3,1) for x in fz], 'k_analyse ms', [round(x/1e6,2) for x in ka])
P
done
