"""GPU tests of the real-time path and the runtime parameter surface (b, f2), through the
C ABI: worker threads (fx_rt_start) instead of an analysis call on the pushing thread, track groups that advance
independently, tracks created / destroyed while the engine runs, clearBuffer, gain and sample-rate changes mid-stream,
per-track parameters, all-or-nothing overruns.  Reference behaviour cited per test (paths relative to /root/reference/Source)."""
import time

import numpy as np
import pytest

import oracle_util as ou

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fx():
    import fxb200

    fxb200.load_library()
    return fxb200


@pytest.fixture(scope="module")
def oracle():
    return ou.best_oracle()


def wait_for(engine, tracks, hop, timeout=10.0):
    t0 = time.time()
    while time.time() - t0 < timeout:
        _, idx = engine.poll_block()
        if all(int(idx[t]) >= hop for t in tracks):
            return True
        time.sleep(0.0005)
    return False


def test_synth_generator_is_bit_identical_on_cpu_and_gpu(fx):
    """one generator, bit-identical on CPU (tests/oracle_util.py::synth_tracks, numpy) and GPU (k_synth), also
    when a long stream is produced slab by slab."""
    import torch

    T, S, sr = 24, 48000 + 4 * 37, 48000.0
    with fx.Engine(n_tracks=T, window=2048, hop=1024, sample_rate=sr, ring_hops=0) as e:
        d = torch.empty((T, S), device="cuda")
        e.synth_device(d.data_ptr(), S, S, first_track=100)
        torch.cuda.synchronize()
        whole = d.cpu().numpy()
        e.synth_device(d.data_ptr(), S, 20000, first_track=100, first_sample=36000)
        torch.cuda.synchronize()
        slab = d[:, :20000].cpu().numpy()
    cpu = ou.synth_tracks(T, S + 8000, sr, first_track=100)
    assert np.array_equal(whole, cpu[:, :S])
    assert np.array_equal(slab, cpu[:, 36000:56000])
    assert np.array_equal(ou.synth_tracks(T, 20000, sr, first_track=100, first_sample=36000), slab)


def test_workers_publish_every_hop_of_the_batch_result(fx):
    """fx_rt_start: the analysis runs on the engine's group workers (the reference's analyser threads,
    RealTimeAnalyser.h:97-127); the pushing thread only copies (AudioDataCollector.h:36-70).  Every published hop equals the
    row of the one-call analysis, bit for bit; the callback reports each group's hops exactly once, in order."""
    N, H, sr, T = 2048, 1024, 48000.0, 5
    hops = 40
    audio = ou.make_tracks(T, hops * H, sr)
    with fx.Engine(n_tracks=T, window=N, hop=H, sample_rate=sr, ring_hops=0) as e:
        one = e.analyse_host(audio)
    seen = {}

    def cb(t0, n, idx, new):
        seen.setdefault(t0, []).append((idx, new))

    with fx.Engine(n_tracks=T, window=N, hop=H, sample_rate=sr, ring_hops=8, tracks_per_group=2) as e:
        e.rt_start(cb)
        blk = 256
        for b0 in range(0, hops * H, blk):
            e.push_block(audio[:, b0:b0 + blk])
            due = (b0 + blk) // H
            if due != b0 // H:
                assert wait_for(e, range(T), due), f"hop {due} never published"
                vec, idx = e.poll_block()
                assert (idx == due).all()
                assert np.array_equal(vec, one["smooth"][:, due - 1], equal_nan=True), due
        st = e.rt_stats()
        e.rt_stop()
        assert st["overruns"] == 0 and st["hops"] == 3 * hops and st["batches"] == 3 * hops      # three groups, one hop per batch
        with pytest.raises(fx.FxError):
            e.rt_start(cb) or e.process()           # fx_process is refused while the workers run
        e.rt_stop()
    assert sorted(seen) == [0, 2, 4]
    for t0, calls in seen.items():
        assert [c[0] for c in calls] == list(range(1, hops + 1)) and all(c[1] == 1 for c in calls)


def test_groups_advance_independently_and_inactive_tracks_do_not_gate(fx):
    """one track that never pushes must not freeze the others.  Groups own their read position; inside a
    group only ACTIVE tracks gate progress (a destroyed AnalyserTrackController stops feeding its collector)."""
    N, H, sr, T = 1024, 512, 44100.0, 4
    audio = ou.make_tracks(T, 12 * H, sr)
    with fx.Engine(n_tracks=T, window=N, hop=H, sample_rate=sr, ring_hops=8, tracks_per_group=2) as e:
        # only group 0 (tracks 0, 1) receives audio
        e.push_block(audio[:2, : 3 * H], first_track=0)
        assert e.process() == 3
        _, idx = e.poll_block()
        assert idx.tolist() == [3, 3, 0, 0]
        # group 1: track 3 never pushes -> the group waits for it ...
        e.push_block(audio[2:3, : 2 * H], first_track=2)
        assert e.process() == 0
        # ... until its controller is gone
        e.set_track_active(3, False)
        assert e.process() == 2
        vec, idx = e.poll_block()
        assert idx.tolist() == [3, 3, 2, 2]
    with fx.Engine(n_tracks=1, window=N, hop=H, sample_rate=sr, ring_hops=0) as e1:
        ref2 = e1.analyse_host(audio[2:3, : 2 * H])
        e1.reset()
        ref_silence = e1.analyse_host(np.zeros((1, 2 * H), np.float32))
    assert np.array_equal(vec[2], ref2["smooth"][0, 1], equal_nan=True)
    assert np.array_equal(vec[3], ref_silence["smooth"][0, 1], equal_nan=True)       # the inactive track is fed silence


def test_track_created_while_the_engine_runs_starts_from_fresh_state(fx):
    """MainComponent creates and destroys AnalyserTrackControllers as channels toggle (MainComponent.cpp:137-171): a new
    controller has a zero overlap buffer (RealTimeAudioAnalysis.h:202), a zero previous spectrum (SpectralCharacteristics.h:34-38)
    and empty histories (RealTimeAnalyser.h:70-74) whatever the engine analysed on that channel before."""
    N, H, sr, T = 2048, 1024, 48000.0, 3
    audio = ou.make_tracks(T, 30 * H, sr)
    late = ou.make_tracks(1, 18 * H, sr, first_track=40)
    with fx.Engine(n_tracks=T, window=N, hop=H, sample_rate=sr, ring_hops=8) as e:
        e.push_block(audio[:, : 6 * H]); assert e.process() == 6
        e.set_track_active(1, False)
        e.push_block(audio[[0, 2], 6 * H: 12 * H][:1], first_track=0)
        e.push_block(audio[2:3, 6 * H: 12 * H], first_track=2)
        assert e.process() == 6
        e.set_track_active(1, True, reset_state=True)
        vec, idx = e.poll(1)
        assert idx == 0 and np.isnan(vec).all()                       # AudioFeatures::getValue before the first push
        got = []
        for k in range(18):
            blk = np.stack([audio[0, (12 + k) * H: (13 + k) * H], late[0, k * H: (k + 1) * H], audio[2, (12 + k) * H: (13 + k) * H]])
            e.push_block(blk)
            assert e.process() == 1
            vec, idx = e.poll(1)
            assert idx == k + 1                                         # hops of THIS track's stream
            got.append(vec)
        others, oidx = e.poll_block()
        assert oidx.tolist() == [30, 18, 30]
    with fx.Engine(n_tracks=1, window=N, hop=H, sample_rate=sr, ring_hops=0) as e1:
        fresh = e1.analyse_host(late)
    with fx.Engine(n_tracks=T, window=N, hop=H, sample_rate=sr, ring_hops=0) as e3:
        whole = e3.analyse_host(audio)
    assert np.array_equal(np.stack(got), fresh["smooth"][0], equal_nan=True)
    assert np.array_equal(others[[0, 2]], whole["smooth"][[0, 2], 29], equal_nan=True)          # the neighbours never noticed


def test_clear_buffer_zeroes_what_was_collected_but_not_analysed(fx):
    """AudioDataCollector::clearBuffer (AudioDataCollector.h:122; fired by play / pause / stop, AnalyserTrackController.h:131-133,
    :167-171): the ring becomes zeros, the indices stay, the overlapper keeps its half window."""
    N, H, sr, T = 2048, 1024, 48000.0, 2
    audio = ou.make_tracks(T, 10 * H, sr)
    with fx.Engine(n_tracks=T, window=N, hop=H, sample_rate=sr, ring_hops=8) as e:
        e.push_block(audio[:, : 3 * H + 300])
        assert e.process() == 3                     # 300 samples stay in the ring
        e.push_block(audio[:, 3 * H + 300: 4 * H + 500])
        e.clear_buffer(1)                           # track 1 only: samples [3 H, 4 H + 500) become zeros
        e.push_block(audio[:, 4 * H + 500:])
        assert e.process() == 7
        e.flush()
        vec, idx = e.poll_block()
    expect = audio.copy()
    expect[1, 3 * H: 4 * H + 500] = 0.0
    with fx.Engine(n_tracks=T, window=N, hop=H, sample_rate=sr, ring_hops=0) as e2:
        ref = e2.analyse_host(expect)
    assert idx.tolist() == [10, 10]
    assert np.array_equal(vec, ref["smooth"][:, 9], equal_nan=True)


def test_overrun_copies_nothing_and_is_counted(fx):
    """a multi-track push that does not fit must not publish some tracks and refuse others."""
    N, H, sr, T = 1024, 512, 44100.0, 3
    audio = ou.make_tracks(T, 40 * H, sr)
    with fx.Engine(n_tracks=T, window=N, hop=H, sample_rate=sr, ring_hops=4) as e:
        e.push_block(audio[:, : 3 * H])
        with pytest.raises(fx.FxError):
            e.push_block(audio[:, 3 * H: 5 * H])     # 5 hops do not fit a 4-hop ring
        assert e.rt_stats()["overruns"] == 1
        assert e.process() == 3
        e.push_block(audio[:, 3 * H: 5 * H])        # the same block goes in once the ring has been drained
        assert e.process() == 2
        vec, idx = e.poll_block()
    with fx.Engine(n_tracks=T, window=N, hop=H, sample_rate=sr, ring_hops=0) as e2:
        ref = e2.analyse_host(audio[:, : 5 * H])
    assert idx.tolist() == [5, 5, 5]
    assert np.array_equal(vec, ref["smooth"][:, 4], equal_nan=True)


def test_per_track_parameters_across_groups(fx, oracle):
    """distinct gain / onset type / window / sensitivity per track (AnalyserTrack's per-track controls,
    AnalyserTrack.h:162-166 -> RealTimeAnalyser.h:244-258, AudioDataCollector.h:124), tracks spread over several groups."""
    N, H, sr, T = 2048, 1024, 48000.0, 7
    audio = ou.make_tracks(T, 120 * H, sr)
    params = [dict(gain=0.3 + 0.25 * t, onset_type=t % 3, onset_hist=3 + t, onset_multiplier=1.1 + 0.1 * t) for t in range(T)]
    with fx.Engine(n_tracks=T, window=N, hop=H, sample_rate=sr, ring_hops=8, tracks_per_group=3) as e:
        for t, p in enumerate(params):
            e.set_gain(p["gain"], track=t)
            e.set_onset(type=p["onset_type"], hist_len=p["onset_hist"], multiplier=p["onset_multiplier"], track=t)
        g = e.analyse_host(audio)
    n_on = 0
    for t, p in enumerate(params):
        o = oracle.analyse(audio[t:t + 1], window=N, hop=H, sample_rate=sr, **p)
        gt = {k: (v[t:t + 1] if isinstance(v, np.ndarray) else v) for k, v in g.items()}
        res = ou.compare(gt, o)
        print(t, p, res)
        assert res["bad_raw"] == 0 and res["bad_lag"] == 0 and res["bad_smooth"] == 0, (t, res)
        n_on += int(o["raw"][..., 0].sum())
    assert n_on > 0


def test_gain_change_between_hops_keeps_the_collected_samples_at_their_gain(fx, oracle):
    """AudioDataCollector::getAnalysisBuffer multiplies by the gain on the way out of the ring
    (AudioDataCollector.h:88), so after setGain the older half of the overlapped window keeps the old gain.  The reference's
    behaviour for a change at a hop boundary is therefore the analysis of the pre-multiplied stream at gain 1."""
    N, H, sr, T = 2048, 512, 48000.0, 3
    audio = ou.make_tracks(T, 60 * H, sr)
    cuts = [(0, 20, 0.8), (20, 21, 1.9), (21, 45, 0.25), (45, 60, 1.0)]
    pre = audio.copy()
    for a, b, gain in cuts:
        pre[:, a * H: b * H] *= np.float32(gain)
    o = oracle.analyse(pre, window=N, hop=H, sample_rate=sr)
    parts = []
    with fx.Engine(n_tracks=T, window=N, hop=H, sample_rate=sr, ring_hops=0) as e:
        for a, b, gain in cuts:
            e.set_gain(gain)
            parts.append(e.analyse_host(audio[:, a * H: b * H]))
    g = {k: np.concatenate([p[k] for p in parts], axis=1) for k in ("raw", "smooth", "diag")}
    res = ou.compare(g, o)
    print(res)
    assert res["bad_raw"] == 0 and res["bad_lag"] == 0 and res["bad_smooth"] == 0
    # without the rescaled overlap the frames after each change would miss by far more than the tolerance
    assert not ou.close(g["raw"][:, 20:24, ou.F["rms"]], oracle.analyse(audio * np.float32(1.9), window=N, hop=H, sample_rate=sr)["raw"][:, 20:24, ou.F["rms"]]).all()


def test_sample_rate_change_mid_stream(fx, oracle):
    """RealTimeAnalyser::sampleRateChanged -> FFTAnalyser::setNyquistValue (RealTimeAnalyser.h:111-114, called on every
    prepareToPlay, AnalyserTrackController.h:178-179): only the nyquist value changes; overlap, previous spectrum and
    histories carry on.  Raw features of a frame depend on the rate in force when it is analysed."""
    N, H, T = 2048, 1024, 4
    sr1, sr2, k = 48000.0, 44100.0, 30
    audio = ou.make_tracks(T, 70 * H, sr1)
    o1 = oracle.analyse(audio, window=N, hop=H, sample_rate=sr1)
    o2 = oracle.analyse(audio, window=N, hop=H, sample_rate=sr2)
    with fx.Engine(n_tracks=T, window=N, hop=H, sample_rate=sr1, ring_hops=0) as e:
        a = e.analyse_host(audio[:, : k * H])
        e.set_sample_rate(sr2)
        b = e.analyse_host(audio[:, k * H:])
    ra = ou.compare(a, {key: v[:, :k] for key, v in o1.items() if key != "frames"})
    assert ra["bad_raw"] == 0 and ra["bad_lag"] == 0 and ra["bad_smooth"] == 0, ra
    # after the change: raw rows as at sr2 from the first frame on; smoothed rows once the 10-deep histories hold only new rows
    gb = {"raw": b["raw"], "diag": b["diag"], "smooth": None}
    rb = ou.compare(gb, {"raw": o2["raw"][:, k:], "diag": o2["diag"][:, k:], "smooth": None})
    assert rb["bad_raw"] == 0 and rb["bad_lag"] == 0, rb
    sm_ok = ou.close(b["smooth"][:, 16:], o2["smooth"][:, k + 16:])
    assert sm_ok.mean() > 0.999, sm_ok.mean()
    assert not ou.close(b["raw"][..., ou.F["f0"]], o1["raw"][:, k:, ou.F["f0"]]).all()          # the rate did change something


def test_osc_batch_encoder_matches_the_published_block(fx):
    """fx_osc_encode_tracks: one pass over the published block -> one OSC 1.0 datagram per track (address pattern, ",f" x 12,
    big-endian floats in the order of OSCFeatureAnalysisOutput.h:107; 10 floats in README.md:55-57 order)."""
    N, H, sr, T = 1024, 512, 44100.0, 6
    audio = ou.make_tracks(T, 14 * H, sr)
    with fx.Engine(n_tracks=T, window=N, hop=H, sample_rate=sr, ring_hops=16, tracks_per_group=4) as e:
        e.push_block(audio)
        assert e.process() == 14
        vec, _ = e.poll_block()
        tracks = [5, 0, 3]
        addrs = [f"/Audio/A{t}" for t in tracks]
        for n_floats, order in ((12, fx.OSC_ORDER_CODE), (10, fx.OSC_ORDER_README)):
            grams = e.osc_encode(tracks, addrs, n_floats=n_floats)
            for t, a, g in zip(tracks, addrs, grams):
                alen = (len(a) + 4) & ~3
                tlen = (1 + n_floats + 4) & ~3
                assert len(g) == alen + tlen + 4 * n_floats
                assert g[:alen] == a.encode() + b"\0" * (alen - len(a))
                assert g[alen:alen + tlen] == b"," + b"f" * n_floats + b"\0" * (tlen - 1 - n_floats)
                vals = np.frombuffer(g[alen + tlen:], ">f4").astype(np.float32)
                expect = np.array([vec[t, fx.FEATURES.index(n)] for n in order], np.float32)
                assert np.array_equal(vals, expect, equal_nan=True)
        assert e.osc_encode([1], ["/x"], stride=16) == [b""]              # does not fit the stride: size 0


def test_sixty_hz_osc_output_of_512_tracks_two_senders_each(fx, tmp_path):
    """SURVEY.md section 8 row f1 as written: OSC at 60 Hz per track from the feature block.  512 AnalyserTrackControllers, two
    senders each (AnalyserTrackController.h:22-23), timers started by connectToAddress (OSCFeatureAnalysisOutput.h:133); the
    shared timer thread encodes all 1024 datagrams of a tick in one pass (fx_osc_encode_tracks) and ships them with sendmmsg.
    Datagrams are counted on both ports and the last one of every track is compared byte for byte with the final features."""
    import os
    import socket
    import subprocess
    import threading

    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    exe = tmp_path / "osc_driver"
    lib = os.path.join(root, "feature-extractor_b200", "lib")
    subprocess.run(["g++", "-std=c++17", "-O2", "-o", str(exe), os.path.join(here, "cpp", "osc_driver.cpp"), "-L" + lib, "-lfxb200", "-lpthread",
                    "-Wl,-rpath," + lib], check=True)
    T, seconds = 512, 2.0
    socks, got, stop = [], [[], []], threading.Event()
    for _ in range(2):
        s = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
        s.setsockopt(socket.SOL_SOCKET, socket.SO_RCVBUF, 64 << 20)
        s.bind(("127.0.0.1", 0))
        s.settimeout(0.2)
        socks.append(s)

    def rx(i):
        while not stop.is_set():
            try:
                got[i].append(socks[i].recv(4096))
            except socket.timeout:
                pass

    threads = [threading.Thread(target=rx, args=(i,)) for i in range(2)]
    for th in threads:
        th.start()
    out = subprocess.run([str(exe), str(T), str(seconds), str(socks[0].getsockname()[1]), str(socks[1].getsockname()[1]), str(tmp_path / "final.f32")],
                         capture_output=True, text=True, timeout=120)
    time.sleep(0.5)
    stop.set()
    for th in threads:
        th.join()
    print(out.stdout, out.stderr[-500:])
    assert out.returncode == 0, out.stderr[-500:]
    words = out.stdout.split()
    ticks, sent = int(words[words.index("ticks") + 1]), int(words[words.index("datagrams") + 1])
    assert "push_errors 0" in out.stdout
    assert ticks >= int(60 * seconds * 0.9)                              # the timer thread held 60 Hz
    # every tick carries both senders of every track (the first ticks run while the 1024 senders are still being constructed)
    assert sent >= (ticks - 10) * 2 * T * 0.98
    final = np.frombuffer((tmp_path / "final.f32").read_bytes(), np.float32).reshape(T, 12)
    for i in range(2):
        assert len(got[i]) >= 0.8 * ticks * T, (len(got[i]), ticks)      # (the loopback socket may drop under this burst rate)
        last = {}
        for g in got[i]:
            assert g.startswith(b"/Audio/A") and len(g) in (76, 80)
            alen = (g.index(b"\0") + 4) & ~3
            assert g[alen:alen + 16] == b",ffffffffffff\0\0\0"
            last[int(g[8:g.index(b"\0")])] = g[alen + 16:]
        assert len(last) == T
        for t in range(T):
            vals = np.frombuffer(last[t], ">f4").astype(np.float32)
            expect = np.array([final[t, fx.FEATURES.index(n)] for n in fx.OSC_ORDER_CODE], np.float32)
            assert np.array_equal(vals, expect, equal_nan=True), t
