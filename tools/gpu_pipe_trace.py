import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "feature-extractor_b200")]
import torch, fxb200
T, N, H, SR = 4096, 4096, 1024, 48000.0
S = (int(SR * 10) // H) * H
F = S // H
eng = fxb200.Engine(n_tracks=T, window=N, hop=H, sample_rate=SR, device=0)
audio = torch.empty((T, S), dtype=torch.float32, device="cuda")
eng.synth_device(audio.data_ptr(), S, S)
torch.cuda.synchronize()
q = (audio.clamp(-1.0, 32767.0 / 32768.0) * 32768.0).round().to(torch.int16)
h_pcm = torch.empty((T, S), dtype=torch.int16, pin_memory=True); h_pcm.copy_(q)
h_s = torch.empty((T, F, 12), dtype=torch.float32, pin_memory=True)
os.environ.pop("FXB200_PIPE_TRACE", None)
eng.analyse_host_pcm_ptr(h_pcm.data_ptr(), "s16le", 1, 0, S * 2, S, None, h_s.data_ptr(), None)
os.environ["FXB200_PIPE_TRACE"] = "1"
eng.analyse_host_pcm_ptr(h_pcm.data_ptr(), "s16le", 1, 0, S * 2, S, None, h_s.data_ptr(), None)
