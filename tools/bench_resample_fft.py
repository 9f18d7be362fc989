import sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
import mlx_audio_primitives_b200 as mb
from oracle import features as of
rng = np.random.default_rng(0)
for B, n, a, b in ((64, 480000, 16000, 22050), (64, 220500, 22050, 16000), (8, 1323000, 44100, 16000)):
    y = rng.standard_normal((B, n)).astype(np.float32)
    yd = torch.from_numpy(y).cuda()
    out = mb.resample(yd, a, b); torch.cuda.synchronize()
    want = of.resample_fft(y[:2], a, b)
    err = np.abs(out[:2].cpu().numpy() - want).max() / np.abs(want).max()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(5): mb.resample(yd, a, b)
    ev1.record(); torch.cuda.synchronize()
    print(f"B={B} n={n} {a}->{b}: {ev0.elapsed_time(ev1)/5:.3f} ms  max rel err {err:.2e}")
