#!/bin/bash
# compute-sanitizer over a small analysis (all three window sizes): memcheck, then racecheck on shared memory
mkdir -p gpurun_out
cat > /tmp/san_small.py <<'PY'
import sys, os
sys.path[:0] = [os.path.join(os.environ.get("GRAFT_REPO_ROOT", "."), "feature-extractor_b200"), os.path.join(os.environ.get("GRAFT_REPO_ROOT", "."), "tests")]
import numpy as np, fxb200, oracle_util as ou
for N, H, sr in ((4096, 1024, 48000.0), (2048, 512, 48000.0), (1024, 512, 44100.0)):
    T = 10
    audio = ou.make_tracks(T, 12 * H, sr)
    audio[3, 5 * H + 7] = np.nan          # the lag search's no-crossing branch (one more block barrier) runs for these frames
    with fxb200.Engine(n_tracks=T, window=N, hop=H, sample_rate=sr) as e:
        g = e.analyse_host(audio)
    print(N, H, g["frames"], float(np.nansum(g["raw"][..., 1])))
PY
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python /tmp/san_small.py > gpurun_out/sanitize_memcheck.log 2>&1
tail -5 gpurun_out/sanitize_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 20 python /tmp/san_small.py > gpurun_out/sanitize_racecheck.log 2>&1
tail -8 gpurun_out/sanitize_racecheck.log
