"""Where does the 16-bit end-to-end step spend its time?  device-resident analysis on the original and on the quantised
workload, the decode kernel alone, and the host pipeline."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "feature-extractor_b200")]
import torch, fxb200
T, N, H, SR = 4096, 4096, 1024, 48000.0
S = (int(SR * 10) // H) * H
F = S // H
dev = torch.device("cuda", 0)
eng = fxb200.Engine(n_tracks=T, window=N, hop=H, sample_rate=SR, device=0)
stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream); sp = stream.cuda_stream
audio = torch.empty((T, S), dtype=torch.float32, device=dev)
smooth = torch.empty((T, F, 12), dtype=torch.float32, device=dev)
eng.synth_device(audio.data_ptr(), S, S, stream=sp)
q = (audio.clamp(-1.0, 32767.0 / 32768.0) * 32768.0).round().to(torch.int16)
deq = q.to(torch.float32) / 32768.0
def timeit(fn, n=3):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(n): fn()
    b.record(stream); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
print("analyse original  ms", timeit(lambda: eng.analyse_device(audio.data_ptr(), S, S, None, smooth.data_ptr(), None, stream=sp)))
print("analyse quantised ms", timeit(lambda: eng.analyse_device(deq.data_ptr(), S, S, None, smooth.data_ptr(), None, stream=sp)))
out = torch.empty_like(audio)
print("decode s16 ms", timeit(lambda: eng.decode_pcm_device(q.data_ptr(), "s16le", 1, 0, S * 2, S, T, out.data_ptr(), S, stream=sp)))
assert torch.equal(out, deq)
h_pcm = torch.empty((T, S), dtype=torch.int16, pin_memory=True); h_pcm.copy_(q)
h_f = torch.empty((T, S), dtype=torch.float32, pin_memory=True); h_f.copy_(deq)
h_s = torch.empty((T, F, 12), dtype=torch.float32, pin_memory=True)
def wall(fn, n=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
print("host pcm16 ms", wall(lambda: eng.analyse_host_pcm_ptr(h_pcm.data_ptr(), "s16le", 1, 0, S * 2, S, None, h_s.data_ptr(), None)))
print("host f32 (quantised) ms", wall(lambda: eng.analyse_host_ptr(h_f.data_ptr(), S, S, None, h_s.data_ptr(), None)))
a = torch.empty((T, S), dtype=torch.int16, device=dev)
print("h2d pcm only ms", wall(lambda: a.copy_(h_pcm, non_blocking=True)))
# does a concurrent host->device copy slow the analysis kernel down?
s2 = torch.cuda.Stream(dev)
def with_copy():
    with torch.cuda.stream(s2):
        a.copy_(h_pcm, non_blocking=True)
    eng.analyse_device(audio.data_ptr(), S, S, None, smooth.data_ptr(), None, stream=sp)
print("analyse with concurrent 3.9 GB h2d ms (event, analysis stream)", timeit(with_copy))
print("same, wall ms", wall(with_copy))
import subprocess
print(subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,pcie.link.gen.current,pcie.link.width.current", "--format=csv,noheader"], capture_output=True, text=True).stdout)
