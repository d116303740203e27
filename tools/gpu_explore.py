"""Debug helper (GPU box): run the CUDA path and the oracle on the same signals and print error statistics."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "feature-extractor_b200"), os.path.join(ROOT, "tests")]
import fxb200
import oracle_util as ou

np.set_printoptions(precision=6, suppress=True, linewidth=220)


def run(N, H, sr, T, seconds, **kw):
    S = (int(sr * seconds) // H) * H
    audio = ou.make_tracks(T, S, sr)
    ora = ou.best_oracle()
    t0 = time.time()
    o = ora.analyse(audio, window=N, hop=H, sample_rate=sr)
    t_cpu = time.time() - t0
    with fxb200.Engine(n_tracks=T, window=N, hop=H, sample_rate=sr) as e:
        t0 = time.time()
        g = e.analyse_host(audio)
        t_gpu = time.time() - t0
    res = ou.compare(g, o)
    print(f"== N={N} H={H} sr={sr} T={T} frames/track={g['frames']} oracle={ora.kind} cpu {t_cpu:.2f}s gpu(host api) {t_gpu:.3f}s")
    print("   ", res)
    ok = ou.close(g["raw"], o["raw"])
    for name, k in ou.F.items():
        a, b = g["raw"][..., k].astype(np.float64), o["raw"][..., k].astype(np.float64)
        fin = np.isfinite(a) & np.isfinite(b)
        err = np.abs(a - b)[fin]
        nbad = int((~ok[..., k]).sum())
        print(f"    {name:9s} max|err| {err.max() if err.size else 0:.3e}  mismatches {nbad:6d}  nonfinite gpu/ora {int((~np.isfinite(a)).sum())}/{int((~np.isfinite(b)).sum())}")
        if nbad:
            idx = np.argwhere(~ok[..., k])[:4]
            for (t, f) in idx:
                print(f"        track {t} frame {f}: gpu {a[t, f]:.7g} ora {b[t, f]:.7g}   gdiag {g['diag'][t, f]}  odiag {o['diag'][t, f]}")
    lagm = g["diag"][..., 1] != o["diag"][..., 1]
    print(f"    lag mismatches {int(lagm.sum())}")
    oks = ou.close(g["smooth"], o["smooth"])
    print(f"    smooth mismatches {int((~oks).sum())} of {oks.size}")
    return res


if __name__ == "__main__":
    out = {}
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    cases = [(1024, 512, 44100.0, 2, 4.0), (2048, 512, 48000.0, 8, 4.0), (4096, 1024, 48000.0, 8, 4.0), (2048, 1024, 48000.0, 16, 6.0)]
    if quick:
        cases = cases[:1]
    for (N, H, sr, T, sec) in cases:
        out[f"{N}_{H}"] = run(N, H, sr, T, sec)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "explore.json"), "w"), indent=1)
