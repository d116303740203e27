"""File ingest: PCM decode.  CPU tests pin the C port against an independent numpy statement of the
conversion; the GPU tests compare kernel k_pcm_decode with the port bit for bit and run the whole analysis from PCM."""
import numpy as np
import pytest

import oracle_util as ou

FORMATS = list(ou.PCM_FORMATS)


def numpy_decode(b: np.ndarray, fmt: str, n_channels: int, channel: int) -> np.ndarray:
    """Independent restatement: left-justify into int32, (float) int32 * 2^-31 (JUCE's 1.0f / 0x7fffffff in fp32)."""
    bps = ou.PCM_BYTES[fmt]
    fr = b.reshape(-1, n_channels, bps)[:, channel, :].astype(np.uint32)
    if fmt.endswith("be"):
        fr = fr[:, ::-1]
    if fmt.startswith("f32"):
        u = fr[:, 0] | (fr[:, 1] << 8) | (fr[:, 2] << 16) | (fr[:, 3] << 24)
        return u.astype(np.uint32).view(np.float32)
    u = np.zeros(len(fr), np.uint32)
    for i in range(bps):
        u |= fr[:, i] << np.uint32(8 * (4 - bps + i))
    if fmt == "u8":
        u ^= np.uint32(0x80000000)
    return (u.view(np.int32).astype(np.float32) * np.float32(2.0 ** -31)).astype(np.float32)


@pytest.mark.parametrize("fmt", FORMATS)
@pytest.mark.parametrize("n_channels,channel", [(1, 0), (2, 1), (3, 0)])
def test_port_matches_numpy(fmt, n_channels, channel):
    rng = np.random.default_rng(7)
    n = 1000
    raw = rng.integers(0, 256, n * n_channels * ou.PCM_BYTES[fmt], dtype=np.uint8)
    if fmt.startswith("f32"):
        raw = rng.uniform(-1, 1, n * n_channels).astype("<f4" if fmt.endswith("le") else ">f4").view(np.uint8)
    got = ou.pcm_decode(raw, fmt, n_channels, channel)
    assert np.array_equal(got.view(np.uint32), numpy_decode(raw, fmt, n_channels, channel).view(np.uint32))


def test_port_known_answers():
    """Extremes and zero of every integer width: full-scale negative is exactly -1, the largest positive is 1 - 2^-(bits-1),
    8-bit WAV is offset binary (0x80 is silence)."""
    assert ou.pcm_decode(np.array([0x00, 0x80, 0xFF], np.uint8), "u8").tolist() == [-1.0, 0.0, 127 / 128]
    assert ou.pcm_decode(np.array([0x80, 0x00, 0x7F], np.uint8), "s8").tolist() == [-1.0, 0.0, 127 / 128]
    assert ou.pcm_decode(np.array([0x00, 0x80, 0x00, 0x00, 0xFF, 0x7F], np.uint8), "s16le").tolist() == [-1.0, 0.0, 32767 / 32768]
    assert ou.pcm_decode(np.array([0x80, 0x00, 0x7F, 0xFF], np.uint8), "s16be").tolist() == [-1.0, 32767 / 32768]
    assert ou.pcm_decode(np.array([0x00, 0x00, 0x80, 0xFF, 0xFF, 0x7F], np.uint8), "s24le").tolist() == [-1.0, 8388607 / 8388608]
    assert ou.pcm_decode(np.array([0x80, 0x00, 0x00], np.uint8), "s24be").tolist() == [-1.0]
    # 32-bit: (float) int32 rounds to 24 bits first; INT32_MAX rounds up to 2^31 -> exactly 1.0
    assert ou.pcm_decode(np.array([0xFF, 0xFF, 0xFF, 0x7F, 0x00, 0x00, 0x00, 0x80], np.uint8), "s32le").tolist() == [1.0, -1.0]


def test_sixteen_bit_round_trip_is_exact():
    x = ou.make_tracks(1, 4096, 48000.0)[0] * 0.2
    q = np.round(x.astype(np.float64) * 32768).astype(np.int16)
    dec = ou.pcm_decode(q.astype("<i2").view(np.uint8), "s16le")
    assert np.array_equal(dec, (q.astype(np.float64) / 32768).astype(np.float32))


# ---------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def fx():
    import fxb200
    fxb200.load_library()
    return fxb200


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", FORMATS)
def test_gpu_decode_bit_exact(fx, fmt):
    import torch
    rng = np.random.default_rng(11)
    T, n, nch, ch = 3, 4099, 2, 1
    bps = ou.PCM_BYTES[fmt]
    if fmt.startswith("f32"):
        rows = rng.uniform(-1, 1, (T, n * nch)).astype("<f4" if fmt.endswith("le") else ">f4").view(np.uint8).reshape(T, -1)
    else:
        rows = rng.integers(0, 256, (T, n * nch * bps), dtype=np.uint8)
    with fx.Engine(n_tracks=T, window=1024, hop=512, sample_rate=44100.0) as e:
        d_pcm = torch.from_numpy(rows).cuda()
        d_out = torch.zeros((T, n + 5), dtype=torch.float32, device="cuda")
        e.decode_pcm_device(d_pcm.data_ptr(), fmt, nch, ch, rows.shape[1], n, T, d_out.data_ptr(), n + 5)
        torch.cuda.synchronize()
        got = d_out.cpu().numpy()
    for t in range(T):
        want = ou.pcm_decode(rows[t], fmt, nch, ch)
        assert np.array_equal(got[t, :n].view(np.uint32), want.view(np.uint32)), (fmt, t)
        assert (got[t, n:] == 0).all()


@pytest.mark.gpu
def test_gpu_decode_mono_s16_vector_path(fx):
    """16-byte aligned mono 16-bit rows take the vector kernel; a ragged tail falls to the scalar one."""
    import torch
    rng = np.random.default_rng(12)
    T, n = 5, 8 * 1000 + 3
    rows = np.zeros((T, 2 * n + 12), np.uint8)          # row pitch 16018 bytes: not 16-byte aligned -> scalar path
    rows[:, : 2 * n] = rng.integers(0, 256, (T, 2 * n), dtype=np.uint8)
    rows16 = np.zeros((T, 16016), np.uint8)             # 16-byte aligned pitch -> vector path + scalar tail
    rows16[:, : 2 * n] = rows[:, : 2 * n]
    with fx.Engine(n_tracks=T, window=1024, hop=512, sample_rate=44100.0) as e:
        outs = []
        for r in (rows, rows16):
            d_pcm = torch.from_numpy(r).cuda()
            d_out = torch.zeros((T, n + 1), dtype=torch.float32, device="cuda")      # stride n + 1 = 8004: multiple of 4
            l0 = e.kernel_launches
            e.decode_pcm_device(d_pcm.data_ptr(), "s16le", 1, 0, r.shape[1], n, T, d_out.data_ptr(), n + 1)
            torch.cuda.synchronize()
            outs.append((d_out.cpu().numpy(), e.kernel_launches - l0))
    assert outs[0][1] == 1 and outs[1][1] == 2
    for t in range(T):
        want = ou.pcm_decode(rows[t, : 2 * n], "s16le")
        for got, _ in outs:
            assert np.array_equal(got[t, :n], want)


@pytest.mark.gpu
@pytest.mark.parametrize("fmt,nch,ch", [("s16le", 1, 0), ("s24le", 2, 1), ("s16be", 2, 0), ("f32le", 1, 0), ("u8", 1, 0)])
def test_gpu_analysis_from_pcm_matches_float_path(fx, fmt, nch, ch):
    """fx_analyse_host_pcm == fx_analyse_host on the decoded samples (bit for bit), and == the oracle on them."""
    N, H, sr, T = 2048, 1024, 48000.0, 4
    S = 40 * H + 17                                           # ragged: only complete hops are analysed
    x = ou.make_tracks(T, S, sr) * 0.2
    other = ou.make_tracks(T, S, sr, first_track=100) * 0.2
    chans = [other] * nch
    chans[ch] = x
    inter = np.stack(chans, axis=-1).reshape(T, S * nch)     # sample frames of nch interleaved samples
    rows = ou.pcm_encode(inter, fmt)
    rows = np.ascontiguousarray(rows.reshape(T, -1))
    dec = np.stack([ou.pcm_decode(rows[t], fmt, nch, ch) for t in range(T)])
    with fx.Engine(n_tracks=T, window=N, hop=H, sample_rate=sr) as e:
        g_pcm = e.analyse_host_pcm(rows, fmt, nch, ch)
    with fx.Engine(n_tracks=T, window=N, hop=H, sample_rate=sr) as e:
        g_f32 = e.analyse_host(dec)
    assert g_pcm["frames"] == g_f32["frames"] == 40
    for k in ("raw", "smooth", "diag"):
        assert np.array_equal(g_pcm[k], g_f32[k], equal_nan=True), k
    o = ou.best_oracle().analyse(dec, window=N, hop=H, sample_rate=sr)
    res = ou.compare(g_pcm, o)
    assert res["bad_raw"] == 0 and res["bad_lag"] == 0 and res.get("bad_smooth", 0) == 0, res


@pytest.mark.gpu
def test_gpu_pcm_argument_errors(fx):
    with fx.Engine(n_tracks=1, window=1024, hop=512, sample_rate=44100.0) as e:
        buf = np.zeros((1, 4096), np.uint8)
        import ctypes
        nf = ctypes.c_long(0)
        L = e.lib
        assert L.fx_analyse_host_pcm(e._h, buf.ctypes.data, 99, 1, 0, 4096, 1024, None, None, None, ctypes.byref(nf)) == -3   # FX_ERR_UNSUPPORTED
        assert L.fx_analyse_host_pcm(e._h, buf.ctypes.data, 3, 2, 2, 4096, 1024, None, None, None, ctypes.byref(nf)) == -1    # channel out of range
        assert L.fx_analyse_host_pcm(e._h, buf.ctypes.data, 3, 1, 0, 100, 1024, None, None, None, ctypes.byref(nf)) == -1     # row shorter than the samples
        assert L.fx_pcm_bytes_per_sample(5) == 3 and L.fx_pcm_bytes_per_sample(0) == 0
