#!/usr/bin/env python
"""Throughput of every compiled plan through the fused kernels (stft and melspectrogram, hop = n_fft / 4,
64 clips x 20 s at 22.05 kHz): one JSON line per n_fft with audio-seconds/second and its ratio to the
neighbouring powers of two.    python tools/bench_plans.py"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mlx_audio_primitives_b200 as ap
from bench_configs import event_ms

SIZES = [32, 64, 128, 256, 400, 480, 512, 600, 800, 1000, 1024, 1200, 1600, 2000, 2048, 3072, 4096, 8192, 1536]
B, L, sr = 64, 441000, 22050
y = torch.randn((B, L), device="cuda")
rows = {}
for n in SIZES:
    hop = n // 4
    ms_s = event_ms(lambda: ap.stft(y, n, hop), 5)
    ms_m = event_ms(lambda: ap.melspectrogram(y, sr=sr, n_fft=n, hop_length=hop, n_mels=64), 5) if n >= 128 else None
    rows[n] = dict(n_fft=n, fast_plan=bool(ap._extension._ext.mlxa_has_fast_plan(n)), stft_ms=ms_s, mel_ms=ms_m,
                   stft_audio_s_per_s=B * L / sr / (ms_s * 1e-3), mel_audio_s_per_s=(B * L / sr / (ms_m * 1e-3)) if ms_m else None)
pow2 = sorted(n for n in rows if n & (n - 1) == 0)
for n, r in rows.items():
    if n & (n - 1):
        lo = max(p for p in pow2 if p < n); hi = min(p for p in pow2 if p > n)
        r["stft_vs_pow2_neighbours"] = [r["stft_audio_s_per_s"] / rows[lo]["stft_audio_s_per_s"], r["stft_audio_s_per_s"] / rows[hi]["stft_audio_s_per_s"]]
        if r["mel_ms"] and rows[lo]["mel_ms"]:
            r["mel_vs_pow2_neighbours"] = [r["mel_audio_s_per_s"] / rows[lo]["mel_audio_s_per_s"], r["mel_audio_s_per_s"] / rows[hi]["mel_audio_s_per_s"]]
    print(json.dumps(r), flush=True)
