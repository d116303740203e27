"""Generate the golden fixtures in this directory from the REFERENCE ITSELF (oracle/_ref/libfxref.so = the headers
under /root/reference/Source compiled headless against oracle/juce_shim).  Run in the build container only:

    python tests/golden/make_golden.py

The fixtures are small .npz files: seeded input (regenerated from oracle_util.make_signal, stored as a checksum
plus the first samples) and the reference's raw / smoothed feature rows and observable diagnostics.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_util as ou

CASES = {
    # name: (window, hop, sample_rate, n_tracks, seconds, mode, extra oracle config)
    "c1_n1024_h512_sr44100_modeA": (1024, 512, 44100.0, 1, 10.0, 0, {}),                 # BASELINE configs[0], verbatim run() bodies
    "c2_n2048_h512_sr48000": (2048, 512, 48000.0, 8, 2.0, 1, {}),                         # configs[1] shape, 8-track slice
    "c3_n4096_h1024_sr48000": (4096, 1024, 48000.0, 8, 2.0, 1, {}),                       # configs[2] shape, 8-track slice
    "c5_n2048_h1024_sr48000_modeA": (2048, 1024, 48000.0, 8, 3.0, 0, {}),                 # configs[3]/[4] shape (reference defaults)
    "params_n2048_h1024": (2048, 1024, 48000.0, 4, 3.0, 1, dict(gain=0.7, onset_type=2, onset_hist=7, onset_multiplier=1.2, rms_pushes=1)),
}


def main():
    ref = ou.reference()
    if ref is None:
        raise SystemExit("oracle/_ref/libfxref.so missing: run `make -C oracle` where /root/reference exists")
    for name, (N, H, sr, T, sec, mode, extra) in CASES.items():
        S = (int(sr * sec) // H) * H
        audio = ou.make_tracks(T, S, sr)
        r = ref.analyse(audio, window=N, hop=H, sample_rate=sr, mode=mode, **extra)
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            window=N, hop=H, sample_rate=sr, n_tracks=T, n_samples=S, mode=mode,
            extra=np.array(sorted(extra.items()), dtype=object) if extra else np.array([], dtype=object),
            audio_head=audio[:, :64], audio_sum=audio.astype(np.float64).sum(axis=1),
            raw=r["raw"], smooth=r["smooth"], lag=r["diag"][..., ou.D["lag"]], true_oer=r["diag"][..., ou.D["true_oer"]],
        )
        print(name, r["raw"].shape)


def legacy():
    """the legacy offline analyser (AudioAnalysis.h AudioAnalyser) driven headless: tests/test_legacy.py"""
    if not os.path.exists(ou.REF_SO):
        raise SystemExit("oracle/_ref/libfxref.so missing")
    N, sr, T, F, first = 2048, 48000.0, 6, 150, 2
    S = int(sr * 4.0)
    audio = ou.make_tracks(T, S, sr, first_track=first)
    out, la = ou.legacy_analyse(ou.REF_SO, audio, F, N, sr)
    np.savez_compressed(os.path.join(HERE, "legacy_n2048_sr48000.npz"), window=N, sample_rate=sr, n_tracks=T, n_samples=S, n_frames=F,
                        first_track=first, audio_head=audio[:, :64], out=out, log_attack=la)
    print("legacy", out.shape)


if __name__ == "__main__":
    main()
    legacy()
