"""Diagnostics: list every value of a config that compare() counts as bad (mismatch without a low oracle-side margin)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "feature-extractor_b200"), os.path.join(ROOT, "tests")]
import fxb200, oracle_util as ou

N, H, sr, T, sec = (int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4]), float(sys.argv[5])) if len(sys.argv) > 5 else (2048, 512, 48000.0, 64, 60.0)
S = int(sr * sec) // H * H
audio = ou.make_tracks(T, S, sr)
with fxb200.Engine(n_tracks=T, window=N, hop=H, sample_rate=sr, ring_hops=0) as e:
    g = e.analyse_host(audio)
o = ou.best_oracle().analyse(audio, window=N, hop=H, sample_rate=sr)
res = ou.compare(g, o)
print(ou.summary(res))
ok = ou.close(g["raw"], o["raw"])
names = list(ou.F)
dn = list(ou.D)
bad = np.argwhere(~ok)
for t, f, k in bad:
    od, gd = o["diag"][t, f], g["diag"][t, f]
    print(f"track {t} frame {f} {names[k]}: gpu {g['raw'][t, f, k]!r} oracle {o['raw'][t, f, k]!r} | lag g/o {gd[ou.D['lag']]}/{od[ou.D['lag']]} "
          + " ".join(f"{n}={gd[ou.D[n]]:.3g}/{od[ou.D[n]]:.3g}" for n in ("pitch_margin", "peak_margin", "flat_margin", "gate_margin", "onset_margin", "num_peaks", "flat_count", "flat_state")))
gl = np.zeros(g["diag"].shape[:2], bool)
for n in ("pitch_margin", "peak_margin", "flat_margin", "gate_margin", "onset_margin"):
    only = (g["diag"][..., ou.D[n]] < 1e-4) & ~(np.where(o["diag"][..., ou.D[n]] < 0, np.inf, o["diag"][..., ou.D[n]]) < 1e-4)
    both = (g["diag"][..., ou.D[n]] < 1e-4) & (o["diag"][..., ou.D[n]] < 1e-4) & (o["diag"][..., ou.D[n]] >= 0)
    oonly = ~(g["diag"][..., ou.D[n]] < 1e-4) & (o["diag"][..., ou.D[n]] < 1e-4) & (o["diag"][..., ou.D[n]] >= 0)
    print(n, "gpu-only low", int(only.sum()), "both", int(both.sum()), "oracle-only", int(oonly.sum()))
