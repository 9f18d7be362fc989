#!/usr/bin/env python
"""Small invocations of every hot kernel family, for `compute-sanitizer --tool memcheck python tools/sanitize_smoke.py`
(one tool per GPU call; sizes kept tiny because memcheck slows kernels ~20x)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mlx_audio_primitives_b200 as ap

rng = np.random.default_rng(0)
for n_fft in [32, 64, 256, 400, 480, 600, 800, 1000, 1024, 1200, 1600, 2000, 2048, 3072, 4096, 8192]:
    hop = max(1, n_fft // 4)
    for B, frames, center in [(2, 37, True), (3, 5, False)]:
        L = (frames - 1) * hop + (0 if center else n_fft) + 3
        y = torch.from_numpy(rng.standard_normal((B, L)).astype(np.float32)).cuda()
        S = ap.stft(y, n_fft, hop, center=center)
        r = ap.istft(S, hop, center=center)
        if n_fft >= 128:
            M = ap.melspectrogram(y, sr=16000, n_fft=n_fft, hop_length=hop, n_mels=40, center=center)
            D = ap.power_to_db(M)
            C = ap.mfcc(y, sr=16000, n_mfcc=13, n_fft=n_fft, hop_length=hop, n_mels=40, center=center)
    if n_fft >= 64:
        mag = ap.magnitude(ap.stft(y, n_fft))
        g = ap.griffinlim(mag, n_iter=2, n_fft=n_fft, random_state=1)
    torch.cuda.synchronize()
    print("ok", n_fft, flush=True)
# the benchmark shape in small: many tiles per CTA, edge tiles, partial last tile, the planned log-mel and host paths
y = torch.from_numpy(rng.standard_normal((150, 16000 * 3 + 77)).astype(np.float32)).cuda()
plan = ap.LogMelPlan(y.shape[0], y.shape[1], sr=16000, n_fft=400, hop_length=160, n_mels=80)
D = plan(y)
yh = y.cpu().pin_memory()
oh = torch.empty(tuple(D.shape)).pin_memory()
plan.run_host(yh, oh)
assert torch.equal(oh, D.cpu())
torch.cuda.synchronize()
print("ok plan", flush=True)
