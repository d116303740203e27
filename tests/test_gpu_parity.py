"""GPU parity tests (run on the B200 box with `-m gpu`): the CUDA path, called through the C ABI, against the
oracle on the same seeded inputs, against the committed golden fixtures, and -- at BASELINE.json's full sizes --
through size-independent properties.  Tolerances (BASELINE.json north_star): continuous features within 1e-4
(relative to max (1, |value|); inf / NaN compared by class); pitch lag and onset flags exact, except frames whose
decision margin is below 1e-4, which are counted and reported."""
import glob
import os

import numpy as np
import pytest

import oracle_util as ou

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = sorted(p for p in glob.glob(os.path.join(HERE, "golden", "*.npz")) if not os.path.basename(p).startswith("legacy_"))


@pytest.fixture(scope="module")
def fx():
    import fxb200

    fxb200.load_library()          # fails loudly if the CUDA library has not been built
    return fxb200


@pytest.fixture(scope="module")
def oracle():
    return ou.best_oracle()


def assert_parity(res, what="", max_exempt_frac=0.01):
    print(what, ou.summary(res))
    assert res["bad_raw"] == 0, (what, res)
    assert res["bad_lag"] == 0, (what, res)
    assert res.get("bad_smooth", 0) == 0, (what, res)
    # exemptions (low-margin decisions, counted and reported) must stay exceptional on ordinary signals
    assert res["raw_mismatch_exempt"] <= max(2, int(res["frames"] * max_exempt_frac)), (what, res)


CASES = [
    # (window, hop, sr, tracks, seconds)      BASELINE.json configs
    (1024, 512, 44100.0, 1, 10.0),           # configs[0] in full: 1 track, 10 s, 861 frames
    (2048, 512, 48000.0, 64, 3.0),           # configs[1] shape: 64 tracks, hop N/4 (3 s slice; the oracle needs seconds, not minutes)
    (4096, 1024, 48000.0, 32, 4.0),          # configs[2] shape
    (2048, 1024, 48000.0, 48, 4.0),          # configs[3]/[4] shape (reference defaults)
    (1024, 256, 48000.0, 8, 2.0),
    (4096, 2048, 96000.0, 8, 2.0),
    (2048, 64, 48000.0, 2, 0.5),             # many hops per window
    (1024, 1024, 44100.0, 4, 2.0),           # no overlap at all
    (1024, 16, 44100.0, 2, 1.0),             # the smallest hop: windows start off the ring's 32-sample groups, 4 copying threads per hop
    (4096, 32, 48000.0, 1, 0.1),             # 128 hops per window
]


@pytest.mark.parametrize("N,H,sr,T,sec", CASES)
def test_parity_against_oracle(fx, oracle, N, H, sr, T, sec):
    S = (int(sr * sec) // H) * H
    audio = ou.make_tracks(T, S, sr)
    o = oracle.analyse(audio, window=N, hop=H, sample_rate=sr)
    with fx.Engine(n_tracks=T, window=N, hop=H, sample_rate=sr) as e:
        g = e.analyse_host(audio)
        assert e.kernel_launches >= 4
    assert g["frames"] == o["frames"] == S // H
    assert_parity(ou.compare(g, o), f"N={N} H={H} T={T} oracle={oracle.kind}")


def test_config2_full_size(fx, oracle):
    """BASELINE.json configs[1] in full: 64 tracks x 60 s at 48 kHz, 2048-point frames, hop 512 -> 5625 frames per
    track, 360 000 frames, every one of the 12 features of every frame against the oracle."""
    N, H, sr, T = 2048, 512, 48000.0, 64
    S = 60 * 48000
    audio = ou.make_tracks(T, S, sr)
    with fx.Engine(n_tracks=T, window=N, hop=H, sample_rate=sr) as e:
        g = e.analyse_host(audio)
    assert g["frames"] == 5625
    o = oracle.analyse(audio, window=N, hop=H, sample_rate=sr)
    res = ou.compare(g, o)
    assert_parity(res, f"configs[1] full size, oracle={oracle.kind}", max_exempt_frac=0.002)
    # onset flags: exact except exempt frames
    on_g, on_o = g["raw"][..., 0], o["raw"][..., 0]
    assert on_o.sum() > 1000 and abs(float(on_g.sum() - on_o.sum())) <= 5


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_parity_against_golden_fixtures(fx, path):
    z = np.load(path, allow_pickle=True)
    N, H, sr, T, S = int(z["window"]), int(z["hop"]), float(z["sample_rate"]), int(z["n_tracks"]), int(z["n_samples"])
    extra = dict(z["extra"].tolist()) if z["extra"].size else {}
    audio = ou.make_tracks(T, S, sr)
    assert np.array_equal(audio[:, :64], z["audio_head"])
    kw = dict(n_tracks=T, window=N, hop=H, sample_rate=sr)
    if extra:
        kw.update(gain=float(extra["gain"]), onset_type=int(extra["onset_type"]), onset_hist=int(extra["onset_hist"]),
                  onset_multiplier=float(extra["onset_multiplier"]), rms_pushes_per_frame=int(extra["rms_pushes"]))
    with fx.Engine(**kw) as e:
        g = e.analyse_host(audio)
    # values from the fixture (generated from oracle/_ref), decision margins from the port on the same input
    okw = dict(window=N, hop=H, sample_rate=sr, mode=int(z["mode"]))
    if extra:
        okw.update(gain=float(extra["gain"]), onset_type=int(extra["onset_type"]), onset_hist=int(extra["onset_hist"]),
                   onset_multiplier=float(extra["onset_multiplier"]), rms_pushes=int(extra["rms_pushes"]))
    pd = ou.port().analyse(audio, **okw)
    assert np.array_equal(pd["raw"], z["raw"], equal_nan=True)
    diag = pd["diag"].copy()
    diag[..., ou.D["lag"]] = z["lag"]
    o = {"raw": z["raw"], "smooth": z["smooth"], "diag": diag}
    assert_parity(ou.compare(g, o), os.path.basename(path))
    if int(z["mode"]) == 1:
        assert ou.close(g["diag"][..., ou.D["true_oer"]], z["true_oer"]).mean() > 0.99


def test_chunking_streaming_and_split_calls_are_bit_identical(fx):
    """The same stream analysed (a) in one call with many chunks per track, (b) in several calls with carried state,
    (c) through the real-time push/process path, gives bit-identical raw and smoothed features."""
    N, H, sr, T = 2048, 512, 48000.0, 3
    S = 400 * H
    audio = ou.make_tracks(T, S, sr)
    with fx.Engine(n_tracks=T, window=N, hop=H, sample_rate=sr) as e:
        one = e.analyse_host(audio)                      # 3 tracks x 400 frames -> many chunks per track
        e.reset()
        parts = []
        cuts = [0, 7 * H, 8 * H, 150 * H, 151 * H, 333 * H, S]      # 7, 1, 142, 1, 182, 67 frames
        for a, b in zip(cuts[:-1], cuts[1:]):
            parts.append(e.analyse_host(audio[:, a:b]))
        split_raw = np.concatenate([p["raw"] for p in parts], axis=1)
        split_smooth = np.concatenate([p["smooth"] for p in parts], axis=1)
    assert np.array_equal(one["raw"], split_raw, equal_nan=True)
    assert np.array_equal(one["smooth"], split_smooth, equal_nan=True)

    # real-time path: 256-sample blocks (BASELINE configs[3]); every processed hop must reproduce the batch row
    with fx.Engine(n_tracks=T, window=N, hop=H, sample_rate=sr, ring_hops=16, tracks_per_group=2) as e:
        done = 0
        blk = 256
        latest = {}
        for b0 in range(0, 120 * H, blk):
            e.push_block(audio[:, b0:b0 + blk])
            new = e.process()
            if new:
                done += new
                for t in range(T):
                    vec, idx = e.poll(t)
                    assert idx == done
                    latest[t] = vec
                    assert np.array_equal(vec, one["smooth"][t, done - 1], equal_nan=True), (t, done)
        assert done == 120


def test_device_api_matches_host_api(fx):
    import torch

    N, H, sr, T = 4096, 1024, 48000.0, 5
    S = 60 * H
    audio = ou.make_tracks(T, S, sr)
    with fx.Engine(n_tracks=T, window=N, hop=H, sample_rate=sr) as e:
        host = e.analyse_host(audio)
        e.reset()
        d_audio = torch.from_numpy(audio).cuda()
        d_raw = torch.empty((T, 60, 12), device="cuda")
        d_smooth = torch.empty((T, 60, 12), device="cuda")
        d_diag = torch.empty((T, 60, 10), device="cuda")
        s = torch.cuda.current_stream()
        nf = e.analyse_device(d_audio.data_ptr(), S, S, d_raw.data_ptr(), d_smooth.data_ptr(), d_diag.data_ptr(), stream=s.cuda_stream)
        torch.cuda.synchronize()
        assert nf == 60
        assert np.array_equal(host["raw"], d_raw.cpu().numpy(), equal_nan=True)
        assert np.array_equal(host["smooth"], d_smooth.cpu().numpy(), equal_nan=True)
        assert np.array_equal(host["diag"], d_diag.cpu().numpy(), equal_nan=True)
        # unaligned rows (stride not a multiple of 4 floats) take the plain-load path instead of cp.async.bulk
        e.reset()
        pad = torch.zeros((T, S + 3), device="cuda")
        pad[:, :S] = d_audio
        e.analyse_device(pad.data_ptr(), S + 3, S, d_raw.data_ptr(), d_smooth.data_ptr(), None, stream=s.cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(host["raw"], d_raw.cpu().numpy(), equal_nan=True)


def test_edge_cases(fx, oracle):
    N, H, sr = 2048, 1024, 48000.0
    with fx.Engine(n_tracks=2, window=N, hop=H, sample_rate=sr) as e:
        # empty and shorter-than-a-hop inputs: zero frames, no error
        r = e.analyse_host(np.zeros((2, 0), np.float32))
        assert r["frames"] == 0
        r = e.analyse_host(np.zeros((2, H - 1), np.float32))
        assert r["frames"] == 0
        # features before the first analysed hop are NaN, as AudioFeatures::getValue is (RealTimeAnalyser.h:87)
        vec, idx = e.poll(0)
        assert idx == 0 and np.isnan(vec).all()
        # ragged length: only complete hops are analysed
        x = ou.make_tracks(2, 5 * H + 100, sr)
        r = e.analyse_host(x)
        assert r["frames"] == 5
        o = oracle.analyse(x, window=N, hop=H, sample_rate=sr)
        assert_parity(ou.compare(r, o), "ragged")
    # all-silence: zeros everywhere except pitch = sr / 2 / 5000 (lag 2)
    with fx.Engine(n_tracks=1, window=N, hop=H, sample_rate=sr) as e:
        r = e.analyse_host(np.zeros((1, 12 * H), np.float32))
        expect = np.zeros(12, np.float32)
        expect[ou.F["f0"]] = np.float32(sr / 2 / 5000.0)
        assert np.array_equal(r["raw"][0, 6], expect)
        assert (r["diag"][0, :, ou.D["lag"]] == 2).all()
    # degenerate signals at maximum level: full-scale square wave, DC, an impulse train.  Their spectra and
    # autocorrelations are exact zeros plus rounding noise, so peak and lag decisions between noise-level values are
    # decided by each FFT's rounding: those frames must all be flagged low-margin (exempt), none may be a silent miss,
    # and everything that does not hang on such a decision must still agree.
    n = np.arange(40 * H)
    hard = np.stack([np.sign(np.sin(2 * np.pi * 1000 * n / sr)), np.ones_like(n, dtype=np.float64), (n % 997 == 0).astype(np.float64)]).astype(np.float32)
    with fx.Engine(n_tracks=3, window=N, hop=H, sample_rate=sr) as e:
        g = e.analyse_host(hard)
    o = ou.port().analyse(hard, window=N, hop=H, sample_rate=sr)       # the port reports margins on its side too
    assert_parity(ou.compare(g, o), "square/dc/impulses", max_exempt_frac=1.0)
    robust = [ou.F[k] for k in ("rms", "centroid", "spread", "flatness", "ler", "flux", "slope")]
    assert ou.close(g["raw"][..., robust], o["raw"][..., robust]).all()
    assert ou.close(g["raw"][1], o["raw"][1]).all()                     # DC: every feature agrees


@pytest.mark.parametrize("N,H,sr", [(4096, 1024, 48000.0), (2048, 1024, 48000.0), (1024, 512, 44100.0)])
def test_inharmonicity_at_exact_integer_ratios(fx, oracle, N, H, sr):
    """Pitch lags that share a large factor with the window (sr / 2^k and neighbours like 3 * 2^k): many peak bins then have an
    edge whose ratio to f0 is an exact integer, where the reference's own fp64 rounding decides whether a multiple of f0 lies in
    the bin (HarmonicCharacteristics.h:223-236).  The GPU takes those (lag, bin) values from a table evaluated on the host in
    the reference's arithmetic; everything else is integer arithmetic.  Rich harmonic signals put peaks on those bins."""
    # a sine of period 4 L is analysed with lag L (the first threshold crossing is a quarter period); its own bin N / (4 L)
    # lies below f0's with bin * lag = N / 4, and the weak partials at 8, 12, 16, 20 cycles per 4 L sit on multiples of f0
    lags = [l for l in (16, 32, 64, 128, 256, 48, 96, 192, 80, 160, 320, 112, 224) if 4 * l < N]
    S = 16 * H
    n = np.arange(S, dtype=np.float64)
    rng = np.random.default_rng(7)
    rows = []
    for l in lags:
        period = 4.0 * l
        x = 0.5 * np.sin(2 * np.pi * n / period + 0.3)
        for k, a in ((8, 0.02), (12, 0.02), (16, 0.015), (20, 0.01)):
            if k / period < 0.45:
                x += a * np.sin(2 * np.pi * k * n / period + rng.uniform(0, 2 * np.pi))
        rows.append(x + 0.0005 * rng.uniform(-1, 1, S))
    audio = np.stack(rows).astype(np.float32)
    o = oracle.analyse(audio, window=N, hop=H, sample_rate=sr)
    with fx.Engine(n_tracks=len(lags), window=N, hop=H, sample_rate=sr) as e:
        g = e.analyse_host(audio)
    res = ou.compare(g, o)
    assert_parity(res, f"exact-ratio lags N={N}", max_exempt_frac=0.25)
    # the case must actually be exercised: several tracks are analysed at lags that share a factor >= 16 with the window
    glag = g["diag"][..., ou.D["lag"]].astype(np.int64)
    special = ((glag > 0) & ((glag & -glag) >= 16)).any(axis=1)
    assert special.sum() >= 4, glag[:, -1]
    assert (g["raw"][..., ou.F["inharm"]] > 0).any()


def test_nan_sample_takes_the_no_crossing_branch(fx, oracle):
    """A NaN sample makes every cnd value of its frames NaN: no lag crosses the threshold and no value is a strict minimum, so the
    reference reports lag -1 (PitchAnalyser.h:165,188) and f0 = -sr; the frames before and after are untouched.  This is the only
    way into the kernel's no-crossing branch on finite-length frames (a running sum cannot keep growing 1 % per lag)."""
    N, H, sr = 2048, 1024, 48000.0
    x = ou.make_tracks(3, 12 * H, sr)
    x[1, 5 * H + 10] = np.nan
    o = oracle.analyse(x, window=N, hop=H, sample_rate=sr)
    with fx.Engine(n_tracks=3, window=N, hop=H, sample_rate=sr) as e:
        g = e.analyse_host(x)
    assert_parity(ou.compare(g, o), "NaN sample")
    assert (g["diag"][1, 5:7, ou.D["lag"]] == -1).all() and (o["diag"][1, 5:7, ou.D["lag"]] == -1).all()
    assert np.array_equal(np.isnan(g["raw"]), np.isnan(o["raw"]))
    assert np.array_equal(g["raw"][[0, 2]], fx_clean(fx, x[[0, 2]], N, H, sr))          # the other tracks do not notice


def fx_clean(fx, x, N, H, sr):
    with fx.Engine(n_tracks=x.shape[0], window=N, hop=H, sample_rate=sr) as e:
        return e.analyse_host(np.ascontiguousarray(x))["raw"]


def test_runtime_parameters(fx, oracle):
    """gain, onset type / window / sensitivity, single RMS push: same surface as the reference's setters."""
    N, H, sr, T = 2048, 1024, 48000.0, 6
    audio = ou.make_tracks(T, 150 * H, sr)
    kw = dict(gain=0.7, onset_type=2, onset_hist=7, onset_multiplier=1.2)
    o = oracle.analyse(audio, window=N, hop=H, sample_rate=sr, rms_pushes=1, **kw)
    with fx.Engine(n_tracks=T, window=N, hop=H, sample_rate=sr, rms_pushes_per_frame=1) as e:
        e.set_gain(0.7)
        e.set_onset(type=2, hist_len=7, multiplier=1.2)
        g = e.analyse_host(audio)
    assert_parity(ou.compare(g, o), "params")
    for typ in (0, 1):
        o = oracle.analyse(audio, window=N, hop=H, sample_rate=sr, onset_type=typ, onset_multiplier=1.3)
        with fx.Engine(n_tracks=T, window=N, hop=H, sample_rate=sr, onset_type=typ, onset_multiplier=1.3) as e:
            g = e.analyse_host(audio)
        assert_parity(ou.compare(g, o), f"onset type {typ}")
        assert g["raw"][..., 0].sum() == o["raw"][..., 0].sum() > 0


def test_full_size_properties(fx, oracle):
    """BASELINE configs[2] at full width (4096 tracks, N = 4096, H = 1024; 2 s per track to bound the run):
    determinism, track-permutation equivariance, chunking invariance, and a sampled subset against the oracle."""
    import torch

    N, H, sr, T = 4096, 1024, 48000.0, 4096
    S = 2 * 48000 // H * H
    F = S // H
    with fx.Engine(n_tracks=T, window=N, hop=H, sample_rate=sr) as e:
        d_audio = torch.empty((T, S), device="cuda")
        e.synth_device(d_audio.data_ptr(), S, S, first_track=0)
        outs = []
        for _ in range(2):
            e.reset()
            d_raw = torch.empty((T, F, 12), device="cuda")
            d_smooth = torch.empty((T, F, 12), device="cuda")
            d_diag = torch.empty((T, F, 10), device="cuda")
            e.analyse_device(d_audio.data_ptr(), S, S, d_raw.data_ptr(), d_smooth.data_ptr(), d_diag.data_ptr())
            e.flush()
            torch.cuda.synchronize()
            outs.append((d_raw.cpu().numpy(), d_smooth.cpu().numpy(), d_diag.cpu().numpy()))
        assert np.array_equal(outs[0][0], outs[1][0], equal_nan=True)          # idempotent / deterministic
        assert np.array_equal(outs[0][1], outs[1][1], equal_nan=True)
        raw, smooth, diag = outs[0]
        # all three flatness regimes occur at scale
        st = diag[..., ou.D["flat_state"]]
        assert (st == 0).any() and (st == 1).any() and (st == 2).any() and (st == 3).any()
        assert raw[..., 0].sum() > T          # bursts produce onsets
        # permutation equivariance: reversed track order gives reversed rows, bit for bit
        e.reset()
        d_rev = torch.flip(d_audio, dims=[0]).contiguous()
        d_raw = torch.empty((T, F, 12), device="cuda")
        e.analyse_device(d_rev.data_ptr(), S, S, d_raw.data_ptr(), None, None)
        e.flush()
        torch.cuda.synchronize()
        assert np.array_equal(d_raw.cpu().numpy()[::-1], raw, equal_nan=True)
        sample_ids = np.array([0, 1, 6, 7, 14, 15, 1023, 2049, 3000, 4093, 4094, 4095])
        sample_audio = d_audio[torch.from_numpy(sample_ids).cuda()].cpu().numpy()
    # chunking invariance: the same 12 tracks alone (many chunks per track) reproduce their rows from the 4096-track run
    with fx.Engine(n_tracks=len(sample_ids), window=N, hop=H, sample_rate=sr) as e:
        g = e.analyse_host(sample_audio)
    assert np.array_equal(g["raw"], raw[sample_ids], equal_nan=True)
    assert np.array_equal(g["smooth"], smooth[sample_ids], equal_nan=True)
    o = oracle.analyse(sample_audio, window=N, hop=H, sample_rate=sr)
    assert_parity(ou.compare(g, o), "full-size sample")


def test_cpp_facade_matches_reference_wiring(fx, oracle, tmp_path):
    """The C++ host facade (reference class names over the C ABI), driven the way MainComponent drives the
    reference -- one AnalyserTrackController per channel, 256-sample device blocks -- reproduces the smoothed
    features of the reference's own wiring (oracle mode A: collector -> overlapper -> run() bodies)."""
    import socket
    import subprocess

    root = os.path.dirname(HERE)
    exe = tmp_path / "facade_driver"
    subprocess.run(["g++", "-std=c++17", "-O2", "-o", str(exe), os.path.join(HERE, "cpp", "facade_driver.cpp"),
                    "-L" + os.path.join(root, "feature-extractor_b200", "lib"), "-lfxb200",
                    "-Wl,-rpath," + os.path.join(root, "feature-extractor_b200", "lib")], check=True)
    T, N, H, sr, block = 4, 2048, 1024, 48000.0, 256
    S = 80 * H
    audio = ou.make_tracks(T, S, sr)
    (tmp_path / "audio.f32").write_bytes(audio.tobytes())
    rx = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
    rx.bind(("127.0.0.1", 0))
    rx.settimeout(5.0)
    port = rx.getsockname()[1]
    out = subprocess.run([str(exe), str(tmp_path / "audio.f32"), str(T), str(S), str(block), str(sr), str(tmp_path / "out.f32"), str(port)],
                         check=True, capture_output=True, text=True).stdout
    rows = np.frombuffer((tmp_path / "out.f32").read_bytes(), np.float32).reshape(-1, T, 13)
    assert rows.shape[0] == 80 and "hops 80" in out
    assert "push_errors 0" in out and "recreated_hops 3 others_hops 83" in out      # a re-created controller starts afresh, its neighbours carry on
    assert np.array_equal(rows[:, 0, 0], np.arange(1, 81, dtype=np.float32))
    o = oracle.analyse(audio, window=N, hop=H, sample_rate=sr, mode=0)
    g = {"raw": None, "smooth": np.ascontiguousarray(rows[:, :, 1:].transpose(1, 0, 2)), "diag": None}
    ok = ou.close(g["smooth"], o["smooth"])
    assert ok.mean() > 0.999, ok.mean()
    n_on = int(o["smooth"][..., 0].sum())
    assert f"onset_callbacks {n_on}" in out and n_on > 0
    # OSC 1.0 datagram of the first reported hop of track 0: "/Audio/A0", ",ffffffffffff", 12 big-endian floats in
    # the order of OSCFeatureAnalysisOutput.h:107
    pkt = rx.recv(4096)
    assert pkt.startswith(b"/Audio/A0\x00") and len(pkt) == 12 + 16 + 48
    assert pkt[12:28] == b",ffffffffffff\x00\x00\x00"
    vals = np.frombuffer(pkt[28:], ">f4")
    order = [fxb200_name for fxb200_name in fx.OSC_ORDER_CODE]
    expect = np.array([rows[0, 0, 1 + fx.FEATURES.index(n)] for n in order], np.float32)
    assert np.array_equal(vals.astype(np.float32), expect, equal_nan=True)
    rx.close()


def test_ten_minute_stream_in_sixty_carried_calls(fx, oracle):
    """BASELINE configs[4] stream length: 10 minutes per track at 48 kHz, N = 2048 / H = 1024 -> 28 125 frames
    per track, analysed in 60 consecutive calls with the per-track state carried (overlap, previous spectrum, smoothing and onset
    histories), against ONE run of the reference over the whole stream.  The input is the bench workload (Philox generator, bursts
    and silences included), so the 60 call boundaries fall on every kind of frame."""
    N, H, sr, T = 2048, 1024, 48000.0, 2
    S = 10 * 60 * 48000 // H * H
    audio = ou.synth_tracks(T, S, sr, first_track=5)
    F = S // H
    cuts = [round(i * F / 60) * H for i in range(61)]
    parts = []
    with fx.Engine(n_tracks=T, window=N, hop=H, sample_rate=sr, ring_hops=0) as e:
        for a, b in zip(cuts[:-1], cuts[1:]):
            parts.append(e.analyse_host(audio[:, a:b]))
    g = {k: np.concatenate([p[k] for p in parts], axis=1) for k in ("raw", "smooth", "diag")}
    assert g["raw"].shape[1] == 28125
    o = oracle.analyse(audio, window=N, hop=H, sample_rate=sr)
    assert_parity(ou.compare(g, o), "10 min x 60 carried calls", max_exempt_frac=0.002)


@pytest.mark.parametrize("N,H,sr", [(4096, 1024, 48000.0), (2048, 1024, 48000.0), (2048, 512, 48000.0), (1024, 512, 44100.0)])
def test_features_do_not_depend_on_the_diagnostics(fx, N, H, sr):
    """A call that does not ask for the diagnostics runs the kernel instantiation without the decision margins (they feed
    nothing).  Raw and smoothed features must be the same bits either way, including NaN input (the no-crossing branch of
    the lag search) and silence."""
    T = 12
    audio = ou.make_tracks(T, 40 * H, sr)
    audio[3, 7 * H + 5] = np.nan
    audio[4] = 0.0
    with fx.Engine(n_tracks=T, window=N, hop=H, sample_rate=sr, ring_hops=0) as e:
        a = e.analyse_host(audio, want_diag=True)
        e.reset()
        b = e.analyse_host(audio, want_diag=False)
    assert b["diag"] is None
    assert np.array_equal(a["raw"], b["raw"], equal_nan=True)
    assert np.array_equal(a["smooth"], b["smooth"], equal_nan=True)
