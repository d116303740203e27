"""File route of the C++ facade: AudioFormatReader parses WAV / AIFF containers on the host (CPU tests),
AudioFilePlayer::analyseLoadedFile sends the file's PCM to the GPU as it lies in the file (GPU test)."""
import os
import struct
import subprocess
import wave

import numpy as np
import pytest

import oracle_util as ou

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIBDIR = os.path.join(ROOT, "feature-extractor_b200", "lib")


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    if not os.path.exists(os.path.join(LIBDIR, "libfxb200.so")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "feature-extractor_b200"), "-s"], check=True)
    exe = tmp_path_factory.mktemp("bin") / "file_driver"
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-o", str(exe), os.path.join(HERE, "cpp", "file_driver.cpp"),
                    "-L" + LIBDIR, "-lfxb200", "-Wl,-rpath," + LIBDIR], check=True)
    return str(exe)


def write_wav(path, x, sr, bits):
    """x: float [frames, channels] in [-1, 1).  stdlib wave for 8/16/24/32-bit PCM."""
    q = np.clip(np.round(x * (1 << (bits - 1))), -(1 << (bits - 1)), (1 << (bits - 1)) - 1).astype(np.int64)
    with wave.open(str(path), "wb") as w:
        w.setnchannels(x.shape[1]); w.setsampwidth(bits // 8); w.setframerate(int(sr))
        if bits == 8:
            w.writeframes((q + 128).astype(np.uint8).tobytes())
        else:
            u = (q & ((1 << bits) - 1)).astype(np.uint64)
            by = np.stack([((u >> (8 * i)) & 0xFF).astype(np.uint8) for i in range(bits // 8)], axis=-1)
            w.writeframes(by.tobytes())
    return q


def ext80(v: float) -> bytes:
    """80-bit IEEE extended, as AIFF stores the sample rate."""
    import math
    m, e = math.frexp(v)                       # v = m * 2^e, 0.5 <= m < 1
    mant = int(m * (1 << 64))
    return struct.pack(">HQ", e - 1 + 16383, mant)


def write_aiff(path, x, sr, bits, aifc=None):
    q = np.clip(np.round(x * (1 << (bits - 1))), -(1 << (bits - 1)), (1 << (bits - 1)) - 1).astype(np.int64)
    u = (q & ((1 << bits) - 1)).astype(np.uint64)
    by = np.stack([((u >> (8 * i)) & 0xFF).astype(np.uint8) for i in range(bits // 8)], axis=-1)
    if aifc != b"sowt":
        by = by[..., ::-1]
    data = np.ascontiguousarray(by).tobytes()
    comm = struct.pack(">hIh", x.shape[1], x.shape[0], bits) + ext80(sr)
    if aifc:
        comm += aifc + b"\x00\x00"
    ssnd = struct.pack(">II", 0, 0) + data
    chunks = b"COMM" + struct.pack(">I", len(comm)) + comm + b"SSND" + struct.pack(">I", len(ssnd)) + ssnd + (b"\x00" if len(ssnd) & 1 else b"")
    if aifc:
        chunks = b"FVER" + struct.pack(">II", 4, 0xA2805140) + chunks
    form = (b"AIFC" if aifc else b"AIFF") + chunks
    open(path, "wb").write(b"FORM" + struct.pack(">I", len(form)) + form)
    return q


def header(driver, path):
    return subprocess.run([driver, "header", str(path)], check=True, capture_output=True, text=True).stdout.strip()


def test_wav_headers(driver, tmp_path):
    x = np.stack([ou.make_signal(3000, 44100.0, t) * 0.2 for t in range(2)], axis=1)
    for bits, fmt in ((8, 1), (16, 3), (24, 5), (32, 7)):
        p = tmp_path / f"a{bits}.wav"
        write_wav(p, x, 44100, bits)
        assert header(driver, p) == f"WAV file rate 44100.000 channels 2 bits {bits} float 0 frames 3000 format {fmt} bytes {3000 * 2 * bits // 8}"
    # IEEE float WAV with an extra LIST chunk of odd length before the data (word alignment) and an 18-byte fmt chunk
    f32 = x.astype("<f4").tobytes()
    fmtc = struct.pack("<HHIIHHH", 3, 2, 48000, 48000 * 8, 8, 32, 0)
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmtc)) + fmtc + b"LIST" + struct.pack("<I", 5) + b"INFOx" + b"\x00" + b"data" + struct.pack("<I", len(f32)) + f32
    p = tmp_path / "f32.wav"
    p.write_bytes(b"RIFF" + struct.pack("<I", len(body)) + body)
    assert header(driver, p) == "WAV file rate 48000.000 channels 2 bits 32 float 1 frames 3000 format 9 bytes 24000"
    # WAVE_FORMAT_EXTENSIBLE, 24-bit PCM sub-format
    pcm = ou.pcm_encode(x.reshape(-1), "s24le").tobytes()
    ext = struct.pack("<HHIIHHHHI", 0xFFFE, 2, 96000, 96000 * 6, 6, 24, 22, 24, 3) + struct.pack("<H", 1) + b"\x00\x00\x00\x00\x10\x00\x80\x00\x00\xaa\x00\x38\x9b\x71"
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(ext)) + ext + b"data" + struct.pack("<I", len(pcm)) + pcm
    p = tmp_path / "ext.wav"
    p.write_bytes(b"RIFF" + struct.pack("<I", len(body)) + body)
    assert header(driver, p) == "WAV file rate 96000.000 channels 2 bits 24 float 0 frames 3000 format 5 bytes 18000"


def test_aiff_headers(driver, tmp_path):
    x = np.stack([ou.make_signal(2001, 48000.0, t) * 0.2 for t in range(3)], axis=1)
    for bits, fmt in ((8, 2), (16, 4), (24, 6), (32, 8)):
        p = tmp_path / f"a{bits}.aiff"
        write_aiff(p, x, 48000.0, bits)
        assert header(driver, p) == f"AIFF file rate 48000.000 channels 3 bits {bits} float 0 frames 2001 format {fmt} bytes {2001 * 3 * bits // 8}"
    p = tmp_path / "sowt.aifc"
    write_aiff(p, x, 44100.0, 16, aifc=b"sowt")
    assert header(driver, p) == "AIFF file rate 44100.000 channels 3 bits 16 float 0 frames 2001 format 3 bytes 12006"


def test_unreadable_and_truncated(driver, tmp_path):
    p = tmp_path / "junk.wav"
    p.write_bytes(b"RIFF\x00\x00\x00\x00WAVEjunkjunkjunk")
    assert header(driver, p) == "unreadable"
    assert header(driver, tmp_path / "missing.wav") == "unreadable"
    x = np.stack([ou.make_signal(1000, 44100.0, 0) * 0.2], axis=1)
    q = tmp_path / "t.wav"
    write_wav(q, x, 44100, 16)
    b = q.read_bytes()
    q.write_bytes(b[: len(b) - 500])                       # the data chunk claims more than the file holds
    assert "frames 750 " in header(driver, q)
    # compressed WAV (ADPCM tag 2) is refused, as JUCE's basic formats refuse it
    fmtc = struct.pack("<HHIIHH", 2, 1, 44100, 44100, 1, 4)
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmtc)) + fmtc + b"data" + struct.pack("<I", 8) + b"\x00" * 8
    p.write_bytes(b"RIFF" + struct.pack("<I", len(body)) + body)
    assert header(driver, p) == "unreadable"


@pytest.mark.gpu
@pytest.mark.parametrize("kind,bits", [("wav", 16), ("wav", 24), ("aiff", 16)])
def test_gpu_file_player_matches_oracle(driver, tmp_path, kind, bits):
    """A 2-channel file analysed by 4 tracks: track t takes channel t % 2; the features equal the oracle's on the samples
    JUCE's reader would have produced (quantised value * 2^-(bits-1))."""
    sr, H, nf = 48000.0, 1024, 30
    S = nf * H + 100
    x = np.stack([ou.make_signal(S, sr, t) * 0.2 for t in range(2)], axis=1)
    p = tmp_path / f"in.{kind}"
    q = write_wav(p, x, sr, bits) if kind == "wav" else write_aiff(p, x, sr, bits)
    out = tmp_path / "out.f32"
    txt = subprocess.run([driver, "analyse", str(p), "4", str(out)], check=True, capture_output=True, text=True).stdout
    assert f"frames {nf}" in txt
    g = np.frombuffer(out.read_bytes(), np.float32).reshape(4, nf, 12)
    dec = (q.astype(np.float64) / (1 << (bits - 1))).astype(np.float32).T            # [channels, frames]
    tracks = np.ascontiguousarray(dec[[0, 1, 0, 1]])
    o = ou.best_oracle().analyse(tracks, window=2048, hop=H, sample_rate=sr)
    ok = ou.close(g, o["smooth"])
    assert ok.mean() > 0.999, ok.mean()
    assert np.array_equal(g[0], g[2], equal_nan=True) and np.array_equal(g[1], g[3], equal_nan=True)
