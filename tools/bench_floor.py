"""Microbenchmark of the top_db floor pass behind the fused log-mel kernel (BASELINE config c2):
time of (mel + floor) minus time of mel alone, for a floor that bites nowhere / on the benchmark
data / everywhere."""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mlx_audio_primitives_b200 as ap

dev = torch.device("cuda")
B, L, sr = 64, 480000, 16000
g = torch.Generator(device=dev); g.manual_seed(42)
t = torch.arange(L, device=dev, dtype=torch.float64) / sr
base = torch.sin(2 * np.pi * (100 + 1000 * t) * t).to(torch.float32)
ys = [base[None, :] + 0.1 * torch.randn((B, L), generator=g, device=dev, dtype=torch.float32) for _ in range(5)]


def timeit(fn, n=200):
    for i in range(10): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

raw = ap.LogMelPlan(B, L, sr=sr, n_fft=400, hop_length=160, n_mels=80, to_db=False)
o0 = raw.empty_output()
print(f"raw mel (no dB, no minima): {timeit(lambda i: raw.mel(ys[i % 5], o0)):7.1f} us")
full = ap.LogMelPlan(B, L, sr=sr, n_fft=400, hop_length=160, n_mels=80, ref="max")
def both_full(i):
    full.mel(ys[i % 5], o0); full.db(o0)
print(f"raw mel + full to_db pass (ref=max path): {timeit(both_full):7.1f} us")
for top_db in (500.0, 80.0, 60.0, 20.0, 1.0):
    plan = ap.LogMelPlan(B, L, sr=sr, n_fft=400, hop_length=160, n_mels=80, top_db=top_db)
    out = plan.empty_output()
    t_mel = timeit(lambda i: plan.mel(ys[i % 5], out))
    plan.block_min.fill_(float("inf")); plan.peaks.zero_()
    def both(i):
        plan.mel(ys[i % 5], out); plan.db(out)
    t_both = timeit(both)
    frac = float((out == out.min()).float().mean())
    print(f"top_db {top_db:6.1f}: mel {t_mel:7.1f} us, mel+floor {t_both:7.1f} us, floor {t_both - t_mel:6.1f} us, clamped fraction {frac:.4f}")
