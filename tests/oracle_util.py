"""Test-side helpers: load the CPU checkers under oracle/, make seeded signals, compare feature blocks.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs may import this module: it is the
checker, never the product path.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
PORT_SO = os.path.join(ORACLE_DIR, "libfxoracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libfxref.so")

NUM_FEATURES = 12
NUM_DIAG = 10
F = {n: i for i, n in enumerate(("onset", "rms", "f0", "centroid", "spread", "flatness", "ler", "flux", "slope", "her", "oer", "inharm"))}
D = {n: i for i, n in enumerate(("true_oer", "lag", "pitch_margin", "num_peaks", "peak_margin", "flat_count", "flat_margin",
                                 "gate_margin", "onset_margin", "flat_state"))}


class OracleConfig(ctypes.Structure):
    _fields_ = [("window", ctypes.c_int), ("hop", ctypes.c_int), ("sample_rate", ctypes.c_double), ("gain", ctypes.c_float),
                ("onset_type", ctypes.c_int), ("onset_hist", ctypes.c_int), ("onset_multiplier", ctypes.c_float),
                ("rms_pushes", ctypes.c_int), ("mode", ctypes.c_int)]


def build_oracle():
    """(Re)build the C port, and oracle/_ref where /root/reference exists."""
    subprocess.run(["make", "-C", ORACLE_DIR, "-s"], check=True, stdout=subprocess.DEVNULL)


class Oracle:
    def __init__(self, path: str):
        self.lib = ctypes.CDLL(path)
        L = self.lib
        L.fxo_default_config.argtypes = [ctypes.POINTER(OracleConfig)]
        L.fxo_kind.restype = ctypes.c_char_p
        L.fxo_analyse_track.restype = ctypes.c_long
        L.fxo_analyse_track.argtypes = [ctypes.POINTER(OracleConfig), ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p,
                                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long]
        L.fxo_analyse_tracks.restype = ctypes.c_long
        L.fxo_analyse_tracks.argtypes = [ctypes.POINTER(OracleConfig), ctypes.c_void_p, ctypes.c_long, ctypes.c_long, ctypes.c_long,
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_int]
        L.fxo_fft_forward.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        L.fxo_fft_inverse.argtypes = [ctypes.c_void_p, ctypes.c_int]
        self.kind = L.fxo_kind().decode()

    def config(self, window=2048, hop=1024, sample_rate=48000.0, **kw) -> OracleConfig:
        c = OracleConfig()
        self.lib.fxo_default_config(ctypes.byref(c))
        c.window, c.hop, c.sample_rate = window, hop, sample_rate
        for k, v in kw.items():
            if not hasattr(c, k):
                raise TypeError(k)
            setattr(c, k, v)
        return c

    def analyse(self, audio: np.ndarray, threads: int = 0, **cfg):
        """audio [T, S] float32 -> dict(raw [T,F,12], smooth [T,F,12], diag [T,F,10])"""
        a = np.ascontiguousarray(np.atleast_2d(audio), dtype=np.float32)
        c = self.config(**cfg)
        T, S = a.shape
        Fr = S // c.hop
        raw = np.zeros((T, Fr, NUM_FEATURES), np.float32)
        smooth = np.zeros((T, Fr, NUM_FEATURES), np.float32)
        diag = np.zeros((T, Fr, NUM_DIAG), np.float32)
        nthreads = threads or min(T, os.cpu_count() or 1)
        n = self.lib.fxo_analyse_tracks(ctypes.byref(c), a.ctypes.data, T, S, S, raw.ctypes.data, smooth.ctypes.data,
                                        diag.ctypes.data, Fr, nthreads)
        assert n == Fr, (n, Fr)
        return {"raw": raw, "smooth": smooth, "diag": diag, "frames": Fr}

    def fft_forward(self, frame: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(frame, dtype=np.float32)
        out = np.zeros(2 * len(x), np.float32)
        self.lib.fxo_fft_forward(x.ctypes.data, len(x), out.ctypes.data)
        return out

    def fft_inverse(self, buf2n: np.ndarray) -> np.ndarray:
        b = np.ascontiguousarray(buf2n, dtype=np.float32).copy()
        self.lib.fxo_fft_inverse(b.ctypes.data, len(b) // 2)
        return b


# legacy offline analyser (AudioAnalysis.h AudioAnalyser): output slots of fxo_legacy_analyse
L = {n: i for i, n in enumerate(("centroid", "spread", "flatness", "flux", "slope", "f0", "her", "inharm", "zcr", "energy", "num_peaks", "margin"))}


def legacy_analyse(lib_path: str, audio: np.ndarray, n_frames: int, window: int = 2048, sample_rate: float = 48000.0):
    """fxo_legacy_analyse of one checker (PORT_SO or REF_SO) on every row of audio: ([T, n_frames, 12], log attack [T])."""
    lib = ctypes.CDLL(lib_path)
    lib.fxo_legacy_analyse.restype = ctypes.c_long
    lib.fxo_legacy_analyse.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_void_p, ctypes.c_long, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    a = np.ascontiguousarray(np.atleast_2d(audio), dtype=np.float32)
    out = np.zeros((a.shape[0], n_frames, len(L)), np.float32)
    la = np.zeros(a.shape[0], np.float32)
    for t in range(a.shape[0]):
        v = ctypes.c_float(0)
        n = lib.fxo_legacy_analyse(window, sample_rate, a[t].ctypes.data, a.shape[1], n_frames, out[t].ctypes.data, ctypes.byref(v))
        assert n == n_frames, n
        la[t] = v.value
    return out, la


def port() -> Oracle:
    if not os.path.exists(PORT_SO):
        build_oracle()
    return Oracle(PORT_SO)


# ---------------------------------------------------------------------------------------------------------
# file ingest (port only: the conversion lives in JUCE, not under /root/reference)
PCM_FORMATS = {"u8": 1, "s8": 2, "s16le": 3, "s16be": 4, "s24le": 5, "s24be": 6, "s32le": 7, "s32be": 8, "f32le": 9, "f32be": 10}
PCM_BYTES = {"u8": 1, "s8": 1, "s16le": 2, "s16be": 2, "s24le": 3, "s24be": 3, "s32le": 4, "s32be": 4, "f32le": 4, "f32be": 4}


def pcm_decode(pcm_row: np.ndarray, fmt: str, n_channels: int = 1, channel: int = 0) -> np.ndarray:
    """Decode one row of interleaved PCM bytes with the C port (oracle/fx_oracle.c: fxo_pcm_decode)."""
    lib = port().lib
    lib.fxo_pcm_decode.restype = ctypes.c_long
    lib.fxo_pcm_decode.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_long, ctypes.c_void_p]
    b = np.ascontiguousarray(pcm_row).view(np.uint8).ravel()
    n = len(b) // (PCM_BYTES[fmt] * n_channels)
    out = np.empty(n, np.float32)
    r = lib.fxo_pcm_decode(b.ctypes.data, PCM_FORMATS[fmt], n_channels, channel, n, out.ctypes.data)
    assert r == n
    return out


def pcm_encode(x: np.ndarray, fmt: str, seed: int = 0) -> np.ndarray:
    """Quantise float samples in [-1, 1) to `fmt` (test-vector generator, numpy only): returns uint8 bytes, sample-major."""
    x = np.asarray(x, np.float64)
    if fmt.startswith("f32"):
        v = x.astype("<f4" if fmt.endswith("le") else ">f4")
        return v.view(np.uint8).reshape(*x.shape[:-1], -1) if x.ndim > 1 else v.view(np.uint8)
    bits = {"u8": 8, "s8": 8}.get(fmt) or int(fmt[1:3])
    q = np.clip(np.round(x * (1 << (bits - 1))), -(1 << (bits - 1)), (1 << (bits - 1)) - 1).astype(np.int64)
    if fmt == "u8":
        return (q + 128).astype(np.uint8)
    if fmt == "s8":
        return q.astype(np.int8).view(np.uint8)
    nb = bits // 8
    u = (q & ((1 << bits) - 1)).astype(np.uint64)
    by = np.stack([((u >> (8 * i)) & 0xFF).astype(np.uint8) for i in range(nb)], axis=-1)      # little endian
    if fmt.endswith("be"):
        by = by[..., ::-1]
    return np.ascontiguousarray(by).reshape(*x.shape[:-1], -1)


def reference() -> Oracle | None:
    """The reference's own headers compiled headless (prebuilt in this container; travels to the GPU box)."""
    return Oracle(REF_SO) if os.path.exists(REF_SO) else None


class CheckedOracle:
    """The checker the parity tests use: VALUES from the reference's own classes (oracle/_ref), decision MARGINS from the
    plain-C port (bit-exact with _ref on every value, and the only one of the two that reports margins).  The exemptions of
    compare() therefore never rest on what the GPU says about its own decisions."""

    def __init__(self):
        self.ref = reference()
        self.port = port()
        self.kind = (self.ref.kind + " values + port margins") if self.ref is not None else (self.port.kind + " (oracle/_ref not built)")

    def analyse(self, audio: np.ndarray, threads: int = 0, **cfg):
        p = self.port.analyse(audio, threads=threads, **cfg)
        if self.ref is None:
            return p
        r = self.ref.analyse(audio, threads=threads, **cfg)
        # the port restates the reference bit for bit (tests/test_oracle.py): anything else is a broken checker, not a margin case
        assert np.array_equal(r["raw"], p["raw"], equal_nan=True), "port and oracle/_ref disagree"
        assert np.array_equal(r["diag"][..., D["lag"]], p["diag"][..., D["lag"]])
        return {"raw": r["raw"], "smooth": r["smooth"], "diag": p["diag"], "frames": r["frames"]}


def best_oracle() -> CheckedOracle:
    return CheckedOracle()


def fastest_oracle() -> Oracle:
    """values only (timing legs of bench.py: the reference build when present)"""
    return reference() or port()


# ---------------------------------------------------------------------------------------------------------
def make_signal(n_samples: int, sr: float, track: int, seed: int = 0x5EED, silence=True, bursts=True) -> np.ndarray:
    """Seeded sine + uniform noise in the spirit of SURVEY.md section 8d (numpy RNG; the same array is fed to the
    oracle and to the GPU, so no cross-platform bit-exactness of the generator is needed)."""
    rng = np.random.default_rng([seed, track])
    n = np.arange(n_samples, dtype=np.float64)
    reg = track % 8
    sigma = 0.5 if reg == 6 else (0.001 if reg == 7 else 0.05)
    freq = 110.0 * 2.0 ** ((track % 48) / 12.0)
    phi = 2 * np.pi * ((track * 0.6180339887498949) % 1.0)
    x = 0.5 * np.sin(2 * np.pi * freq * n / sr + phi) + sigma * rng.uniform(-1.0, 1.0, n_samples)
    isr = int(sr)
    if bursts:
        x[(np.arange(n_samples) % isr) < isr // 20] *= 4.0
    if silence and (track & 1):
        m = np.arange(n_samples) % (2 * isr)
        x[(m >= isr) & (m < isr + isr // 4)] = 0.0
    return x.astype(np.float32)


def make_tracks(n_tracks: int, n_samples: int, sr: float, first_track: int = 0, **kw) -> np.ndarray:
    return np.stack([make_signal(n_samples, sr, first_track + t, **kw) for t in range(n_tracks)])


# ---------------------------------------------------------------------------------------------------------
# The bench workload, bit-identical to the GPU's generator (feature-extractor_b200/csrc/fx_post.cu:
# k_synth): Philox-4x32-10 keyed by seed ^ track, counter = sample index / 4; a sine from IEEE double adds and multiplies
# only.  The reference arm of bench.py and the CPU baseline are fed from here, the GPU arm from k_synth.
def _philox4x32_10(counter: np.ndarray, key: int) -> np.ndarray:
    """counter: uint64 array of block indices -> uint32 array [..., 4]"""
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    mask = np.uint64(0xFFFFFFFF)
    c0 = counter & mask
    c1 = counter >> np.uint64(32)
    c2 = np.zeros_like(c0)
    c3 = np.zeros_like(c0)
    k0, k1 = key & 0xFFFFFFFF, (key >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + 0x9E3779B9) & 0xFFFFFFFF, (k1 + 0xBB67AE85) & 0xFFFFFFFF
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def _synth_sin_turns(r: np.ndarray) -> np.ndarray:
    q = np.floor(4.0 * r + 0.5)
    z = r - 0.25 * q
    a = z * 6.283185307179586
    a2 = a * a
    ps = np.full_like(a, -2.505210838544172e-08)
    for c in (2.755731922398589e-06, -1.984126984126984e-04, 8.333333333333333e-03, -1.666666666666667e-01, 1.0):
        ps = ps * a2 + c
    sn = a * ps
    pc = np.full_like(a, -2.755731922398589e-07)
    for c in (2.480158730158730e-05, -1.388888888888889e-03, 4.166666666666666e-02, -0.5, 1.0):
        pc = pc * a2 + c
    qi = q.astype(np.int64) & 3
    return np.where(qi == 0, sn, np.where(qi == 1, pc, np.where(qi == 2, -sn, -pc)))


def synth_tracks(n_tracks: int, n_samples: int, sr: float, first_track: int = 0, first_sample: int = 0, seed: int = 0x5EED) -> np.ndarray:
    """float32 [n_tracks, n_samples]: samples first_sample .. of tracks first_track .., as fx_synth_device_at writes them."""
    assert first_sample % 4 == 0
    out = np.empty((n_tracks, n_samples), np.float32)
    quads = (n_samples + 3) // 4
    n = first_sample + np.arange(quads * 4, dtype=np.int64)
    isr = int(sr)
    for i in range(n_tracks):
        track = first_track + i
        r = _philox4x32_10((np.uint64(first_sample // 4) + np.arange(quads, dtype=np.uint64)), seed ^ track).reshape(-1)
        u = (r >> np.uint32(8)).astype(np.float64) * (1.0 / 8388608.0) - 1.0
        reg = track & 7
        sigma = 0.5 if reg == 6 else (0.001 if reg == 7 else 0.05)
        freq = 110.0 * 2.0 ** ((track % 48) / 12.0)
        ph = float(track) * 0.61803398874989484820
        ph = ph - np.floor(ph)
        turns = (freq * n.astype(np.float64)) / sr + ph
        turns = turns - np.floor(turns)
        x = 0.5 * _synth_sin_turns(turns) + sigma * u
        x = np.where((n % isr) < isr // 20, x * 4.0, x)
        if track & 1:
            m = n % (2 * isr)
            x = np.where((m >= isr) & (m < isr + isr // 4), 0.0, x)
        out[i] = x[:n_samples].astype(np.float32)
    return out


# ---------------------------------------------------------------------------------------------------------
TOL = 1e-4          # BASELINE.json north_star: <= 1e-4 on normalised [0, 1] outputs
MARGIN_TOL = 1e-4   # decisions whose relative margin is below this are exempt (and counted)


def close(a: np.ndarray, b: np.ndarray, tol: float = TOL) -> np.ndarray:
    """Element-wise parity: inf/NaN by class, otherwise |a - b| <= tol * max (1, |b|)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    both_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    fin = np.isfinite(a) & np.isfinite(b)
    with np.errstate(invalid="ignore"):
        ok = fin & (np.abs(a - b) <= tol * np.maximum(1.0, np.abs(b)))
    return ok | both_nan | both_inf


CAUSES = ("gate", "pitch", "peak", "flat", "onset")


def compare(gpu: dict, ora: dict, tol: float = TOL, margin_tol: float = MARGIN_TOL, smooth_halo: int = 10) -> dict:
    """Compare GPU and oracle feature blocks frame by frame.

    A raw mismatch is EXEMPT only when a decision that feeds the value had a relative margin below margin_tol ON THE
    ORACLE'S SIDE (the port's margins: pitch lag -> f0/her/oer/inharm; peak set -> inharm; flatness gate; silence gates ->
    everything; onset comparisons).  The GPU's own margins are a diagnostic: 'gpu_only_low_margin_frames' counts frames
    the GPU flags and the oracle does not -- they exempt nothing.  An oracle without margins (-1: oracle/_ref alone, golden
    fixtures) exempts nothing either.  Smoothed values are exempt for smooth_halo frames after a mismatching raw frame.
    Returns counts; 'bad_raw' / 'bad_lag' / 'bad_smooth' must be zero for parity.
    """
    graw, oraw = gpu["raw"], ora["raw"]
    gd, od = gpu["diag"], ora["diag"]

    def m(name):   # the oracle's margin
        o = od[..., D[name]]
        return np.where(o < 0, np.inf, o)

    low = {"pitch": m("pitch_margin") < margin_tol, "peak": m("peak_margin") < margin_tol, "flat": m("flat_margin") < margin_tol,
           "gate": m("gate_margin") < margin_tol, "onset": m("onset_margin") < margin_tol}
    glow = np.zeros(low["gate"].shape, bool)
    for name in ("pitch_margin", "peak_margin", "flat_margin", "gate_margin", "onset_margin"):
        glow |= gd[..., D[name]] < margin_tol
    any_low = low["pitch"] | low["peak"] | low["flat"] | low["gate"] | low["onset"]

    ok = close(graw, oraw, tol)
    cause = {c: np.zeros_like(ok) for c in CAUSES}
    for name in ("f0", "her", "oer", "inharm"):
        cause["pitch"][..., F[name]] |= low["pitch"]
    cause["peak"][..., F["inharm"]] |= low["peak"]
    cause["flat"][..., F["flatness"]] |= low["flat"]
    # a flipped silence gate changes everything, including which spectrum the next flux is measured against
    cause["gate"] |= low["gate"][..., None]
    cause["onset"][..., F["onset"]] |= low["onset"]
    exempt = np.zeros_like(ok)
    for c in CAUSES:
        exempt |= cause[c]
    bad_raw = ~ok & ~exempt

    # lag must match exactly unless exempt
    lag_mismatch = (gd[..., D["lag"]] != od[..., D["lag"]])
    bad_lag = lag_mismatch & ~low["pitch"] & ~low["gate"]

    by_cause, left = {}, ~ok & exempt
    for c in CAUSES:                       # each exempted value is attributed to the first cause that covers it
        hit = left & cause[c]
        by_cause[c] = int(hit.sum())
        left &= ~hit

    # error statistics over the values that are compared numerically (finite on both sides, not exempt)
    fin = np.isfinite(graw) & np.isfinite(oraw) & ~exempt
    with np.errstate(invalid="ignore"):
        abs_err = np.where(fin, np.abs(graw.astype(np.float64) - oraw.astype(np.float64)), 0.0)
        rel_err = abs_err / np.maximum(1.0, np.abs(np.where(fin, oraw, 0.0).astype(np.float64)))
    names = list(F)
    res = {
        "frames": int(ok.shape[0] * ok.shape[1]),
        "raw_mismatch_total": int((~ok).sum()),
        "raw_mismatch_exempt": int((~ok & exempt).sum()),
        "bad_raw": int(bad_raw.sum()),
        "lag_mismatch": int(lag_mismatch.sum()),
        "bad_lag": int(bad_lag.sum()),
        "low_margin_frames": int(any_low.sum()),
        "gpu_only_low_margin_frames": int((glow & ~any_low).sum()),
        "exempt_by_cause": by_cause,
        "max_abs_err": {n: float(abs_err[..., F[n]].max()) if abs_err.size else 0.0 for n in names},
        "max_rel_err": {n: float(rel_err[..., F[n]].max()) if rel_err.size else 0.0 for n in names},
    }
    if gpu.get("smooth") is not None and ora.get("smooth") is not None:
        oks = close(gpu["smooth"], ora["smooth"], tol)
        # any raw mismatch (exempt or not) contaminates the next smooth_halo frames of that feature; onset feeds on RMS
        dirty = ~ok
        dirty[..., F["onset"]] |= low["onset"]
        halo = np.zeros_like(dirty)
        for k in range(min(smooth_halo + 6, dirty.shape[1])):
            halo[:, k:, :] |= dirty[:, : dirty.shape[1] - k, :] if k else dirty
        bad_s = ~oks & ~halo
        res["smooth_mismatch_total"] = int((~oks).sum())
        res["bad_smooth"] = int(bad_s.sum())
    return res


def summary(res: dict) -> dict:
    """the counts of a compare() result without the per-feature tables (one line in a log)"""
    return {k: v for k, v in res.items() if k not in ("max_abs_err", "max_rel_err")}
