"""Debug helper (GPU box): dump the frames where GPU and port oracle disagree for one config."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "feature-extractor_b200"), os.path.join(ROOT, "tests")]
import fxb200, oracle_util as ou
np.set_printoptions(precision=7, suppress=False, linewidth=250)
N, H, sr, T, sec = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4]), float(sys.argv[5])
S = (int(sr * sec) // H) * H
audio = ou.make_tracks(T, S, sr)
o = ou.port().analyse(audio, window=N, hop=H, sample_rate=sr)
with fxb200.Engine(n_tracks=T, window=N, hop=H, sample_rate=sr) as e:
    g = e.analyse_host(audio)
print(ou.compare(g, o))
ok = ou.close(g["raw"], o["raw"])
bad = np.argwhere(~ok)
lagm = np.argwhere(g["diag"][..., 1] != o["diag"][..., 1])
print("bad raw", bad[:20].tolist(), "lag mismatch", lagm[:20].tolist())
for (t, f) in {(int(a), int(b)) for a, b, _ in bad[:20]} | {(int(a), int(b)) for a, b in lagm[:20]}:
    print(f"track {t} frame {f}")
    print("  gpu raw ", g["raw"][t, f]); print("  ora raw ", o["raw"][t, f])
    print("  gpu diag", g["diag"][t, f]); print("  ora diag", o["diag"][t, f])
