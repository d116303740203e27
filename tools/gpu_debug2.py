import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "feature-extractor_b200"), os.path.join(ROOT, "tests")]
import fxb200, oracle_util as ou
np.set_printoptions(precision=7, suppress=False, linewidth=250)
N, H, sr = 2048, 1024, 48000.0
n = np.arange(40 * H)
hard = np.stack([np.sign(np.sin(2 * np.pi * 1000 * n / sr)), np.ones_like(n, dtype=np.float64), (n % 997 == 0).astype(np.float64)]).astype(np.float32)
with fxb200.Engine(n_tracks=3, window=N, hop=H, sample_rate=sr) as e:
    g = e.analyse_host(hard)
o = ou.port().analyse(hard, window=N, hop=H, sample_rate=sr)
ok = ou.close(g["raw"], o["raw"])
for t in range(3):
    print("track", t, "mismatch per feature", (~ok[t]).sum(axis=0), "lag mism", int((g["diag"][t,:,1] != o["diag"][t,:,1]).sum()))
    f = 20
    print("  gpu raw ", g["raw"][t, f]); print("  ora raw ", o["raw"][t, f])
    print("  gpu diag", g["diag"][t, f]); print("  ora diag", o["diag"][t, f])
