#!/usr/bin/env python
"""Time all five BASELINE.json configs through the public API on one GPU (secondary to bench.py,
which is the contract benchmark for configs[1]).  Prints one JSON line per config with
audio-seconds/second and the fraction of the SURVEY 8(d) roofline (max of algorithmic bytes at the
measured HBM peak and algorithmic flops at the FP32 peak).

    python tools/bench_configs.py [--configs c1,c2,c3,c4,c5] [--iters 20] [--clips-scale 1.0]
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import mlx_audio_primitives_b200 as ap

HBM_GBS = 6449.7
FP32_TFLOPS = 74.4
try:
    with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
        HBM_GBS = json.load(f).get("hbm_gbs", HBM_GBS)
except Exception:
    pass


def fft_flops(N):
    return 2.5 * N * math.log2(N)


def clips(B, L, sr, seed=0):
    g = torch.Generator(device="cuda"); g.manual_seed(seed + 1000 * RANK)
    t = torch.arange(L, device="cuda", dtype=torch.float64) / sr
    base = torch.sin(2 * np.pi * (100 + 1000 * t) * t).to(torch.float32)
    return base[None] + 0.1 * torch.randn((B, L), generator=g, device="cuda")


WORLD = int(os.environ.get("WORLD_SIZE", "1"))
RANK = int(os.environ.get("RANK", "0"))


def timed(fn, iters, warm=3):
    """ms per call on the device; under torchrun: barrier on both sides, max over ranks."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    if WORLD > 1:
        torch.distributed.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    if WORLD > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())
    return ms


_ROWS = None  # collect() gathers the rows here instead of printing them


def report(name, label, audio_s, ms, bytes_, flops, extra=None):
    # under torchrun every rank holds the same share (weak scaling): whole-job units over the slowest rank's time
    audio_s, bytes_, flops = audio_s * WORLD, bytes_ * WORLD, flops * WORLD
    if RANK != 0:
        return
    t_roof = max(bytes_ / (HBM_GBS * 1e9), flops / (FP32_TFLOPS * 1e12)) * 1e3 / WORLD
    line = {"config": name, "workload": label, "ms": ms, "audio_s_per_s": audio_s / (ms * 1e-3),
            "alg_bytes": bytes_, "alg_flops": flops, "t_roof_ms": t_roof, "frac_of_roofline": t_roof / ms,
            "n_gpus": WORLD, "bound": "hbm" if bytes_ / (HBM_GBS * 1e9) >= flops / (FP32_TFLOPS * 1e12) else "fp32",
            "achieved_GBs": bytes_ / (ms * 1e-3) / 1e9, "achieved_TFLOPs": flops / (ms * 1e-3) / 1e12}
    if extra:
        line.update(extra)
    if _ROWS is not None:
        _ROWS.append(line)
    else:
        print(json.dumps(line), flush=True)


def event_ms(fn, iters):
    """mean device time of fn() over iters calls (CUDA events on the current stream, after warm-up)."""
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def collect(configs, iters=20, clips_scale=1.0):
    """The rows of `configs` as dicts (bench.py puts them into its JSON line).  The process group, if any, is the
    caller's: under torchrun every rank must call this."""
    global _ROWS
    _ROWS = []
    try:
        run(configs, iters, clips_scale)
        return _ROWS
    finally:
        _ROWS = None


def main():
    a = argparse.ArgumentParser()
    a.add_argument("--configs", default="c1,c2,c3,c4,c5,f1,f2,f3")
    a.add_argument("--iters", type=int, default=20)
    a.add_argument("--clips-scale", type=float, default=1.0, help="scale the batch (e.g. 0.125 = one of 8 GPUs' share)")
    args = a.parse_args()
    if WORLD > 1:  # torchrun: one rank per GPU, clips sharded, the dB peak exchanged
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
        ap.distributed.enable()
    run(args.configs.split(","), args.iters, args.clips_scale)


def run(want, iters, sc):
    class _A:
        pass
    args = _A()
    args.iters = iters

    if "c1" in want:  # stft + istft round trip, 1 x 10 s @ 22.05 kHz, 2048/512
        L, N, hop = 220500, 2048, 512
        y = clips(1, L, 22050)
        T, F = 1 + L // hop, N // 2 + 1
        S = ap.stft(y, N, hop)
        ms_f = timed(lambda: ap.stft(y, N, hop), args.iters * 5)
        ms_i = timed(lambda: ap.istft(S, hop, length=L), args.iters * 5)
        r = ap.istft(S, hop, length=L)
        err = float((r[:, 1:] - y[:, 1:]).abs().max())
        by = (4 * L + 8 * F * T) * 2
        fl = T * (2 * fft_flops(N) + N + 4 * N)
        report("c1", "stft+istft 2048/512 hann 1x10 s @22.05k", 10.0, ms_f + ms_i, by, fl,
               {"ms_stft": ms_f, "ms_istft": ms_i, "round_trip_max_err": err, "clips": 1,
                "kernel": "fwd_kernel<EP_STFT> + inv_kernel n_fft=2048 (one launch each; launch/latency bound: 432 frames)",
                "kernel_ms": ms_f + ms_i})
    if "c2" in want:
        B, L = max(1, int(64 * sc)), 480000
        y = clips(B, L, 16000)
        fn = lambda: ap.power_to_db(ap.melspectrogram(y, sr=16000, n_fft=400, hop_length=160, n_mels=80))
        ms = timed(fn, args.iters * 5)
        T, F = 1 + L // 160, 201
        report("c2", f"log-mel 16k 400/160 80 mels {B}x30 s", B * 30.0, ms, B * (4 * L + 4 * 80 * T),
               B * T * (fft_flops(400) + 400 + 7 * F + 3 * 80))
    if "c3" in want:
        # BASELINE configs[2]: 1024 x 30 s sharded over 8 GPUs = 128 clips per GPU; under torchrun every rank owns 128
        # clips and the batch-global peak of ref=max crosses the GPUs (peer-memory exchange / NCCL)
        B, L = max(1, int(1024 * sc * 0.125)), 661500
        y = clips(B, L, 22050)
        fn = lambda: ap.power_to_db(ap.melspectrogram(y, sr=22050, n_fft=2048, hop_length=512, n_mels=128), ref=torch.max)
        ms = timed(fn, args.iters)
        plan = ap.LogMelPlan(B, L, sr=22050, n_fft=2048, hop_length=512, n_mels=128, ref="max", top_db=80.0)
        out = plan.empty_output()
        def both():
            plan.mel(y, out); plan.db(out)
        both(); both()
        k_ms = event_ms(lambda: plan.mel(y, out), 5) if WORLD == 1 else None  # (alone only without the exchange's pairing)
        if WORLD > 1:
            both()
        T, F = 1 + L // 512, 1025
        report("c3", f"mel+power_to_db(ref=max) 22.05k 2048/512 128 mels {B}x30 s per GPU", B * 30.0, ms,
               B * (4 * L + 4 * 128 * T), B * T * (fft_flops(2048) + 2048 + 7 * F + 3 * 128),
               {"clips": B * WORLD, "kernel": "fwd_kernel<EP_MEL> n_fft=2048", "kernel_ms": k_ms})
        del plan, out
    if "c4" in want:
        B, L = max(1, int(256 * sc)), 2646000  # BASELINE configs[3]: 256 x 60 s (2.7 GB of input per GPU)
        y = clips(B, L, 44100)
        fn = lambda: ap.mfcc(y, sr=44100, n_mfcc=40, n_fft=4096, hop_length=1024)
        ms = timed(fn, max(3, args.iters // 2))
        k_ms = event_ms(lambda: ap.melspectrogram(y, sr=44100, n_fft=4096, hop_length=1024), 3)
        T, F = 1 + L // 1024, 2049
        report("c4", f"MFCC-40 44.1k 4096/1024 128 mels {B}x60 s per GPU", B * 60.0, ms, B * (4 * L + 4 * 40 * T),
               B * T * (fft_flops(4096) + 4096 + 7 * F + 3 * 128 + 2 * 128 * 40),
               {"clips": B * WORLD, "kernel": "fwd_kernel<EP_MEL> n_fft=4096 (melspectrogram alone; mfcc adds mfcc_tail_kernel)",
                "kernel_ms": k_ms})
        del y
        torch.cuda.empty_cache()
    if "c5" in want:
        B, L, N, hop, iters = max(1, int(128 * sc)), 220500, 1024, 256, 32
        y = clips(B, L, 22050)
        Sc = ap.stft(y, N, hop)
        S = ap.magnitude(Sc)
        fn = lambda: ap.griffinlim(S, n_iter=iters, hop_length=hop, random_state=0)
        ms = timed(fn, max(2, args.iters // 5), warm=1)
        T, F = S.shape[-1], N // 2 + 1
        # dominant kernel: the inverse transform with the signal-domain momentum step (one launch of 33 per call)
        from mlx_audio_primitives_b200.stft import _istft_physical, _spectrum_physical
        from mlx_audio_primitives_b200.windows import padded_window
        win = padded_window("hann", N, N)
        P = _spectrum_physical(Sc)
        u0, u1, yy = (torch.empty((B, L), device="cuda") for _ in range(3))
        k_ms = event_ms(lambda: _istft_physical(P, N, hop, win, True, None, out=yy, u_prev=u0, momentum=0.99, u_out=u1), 5)
        # the other kernel of an iteration: forward transform + projection onto the target magnitudes
        from mlx_audio_primitives_b200._extension import _ext, check
        from mlx_audio_primitives_b200._tensor import ptr, stream_ptr
        magp = S.transpose(1, 2).contiguous()
        cur = torch.empty((B, T, F, 2), dtype=torch.float32, device="cuda")
        f_ms = event_ms(lambda: check(_ext.mlxa_griffinlim_project_f32(ptr(yy), B, L, yy.stride(0), ptr(win), N, hop, 1, 0, T, T,
                                                                     ptr(magp), ptr(cur), stream_ptr(yy)), "griffinlim"), 5)
        del magp, cur
        # per iteration: inverse reads 8FT + 4L (previous inverse), writes 8L; projection reads 4L + 4FT, writes 8FT
        per_it = 8 * F * T + 4 * L + 8 * L + 4 * L + 4 * F * T + 8 * F * T
        by = B * (iters * per_it + 8 * F * T + 12 * L)
        fl = B * T * (iters * (2 * fft_flops(N) + N + 4 * N + 40 * F) + fft_flops(N) + 4 * N)
        report("c5", f"Griffin-Lim 32 it 1024/256 {B}x10 s @22.05k (device RNG init included)", B * 10.0, ms, by, fl,
               {"clips": B, "kernel": "inv_kernel n_fft=1024 (momentum form; 33 launches per call, 32 x fwd_kernel<EP_GL> beside them)",
                "kernel_ms": k_ms, "kernel_alg_bytes": B * (8 * F * T + 12 * L), "project_kernel_ms": f_ms,
                "project_kernel_alg_bytes": B * (4 * L + 12 * F * T)})
        del y, Sc, S, P, u0, u1, yy
    if "f1" in want:  # section 8(f) rank 1: the four spectral features of a music batch (C3's shape), from audio
        B, L, N, hop = max(1, int(1024 * sc * 0.125)), 661500, 2048, 512
        y = clips(B, L, 22050)
        T, F = 1 + L // hop, N // 2 + 1
        kw = dict(sr=22050, n_fft=N, hop_length=hop)

        def feats():
            ap.spectral_centroid(y, **kw); ap.spectral_bandwidth(y, **kw); ap.spectral_rolloff(y, **kw)
            ap.spectral_flatness(y, n_fft=N, hop_length=hop)
        ms = timed(feats, args.iters)
        ms_contrast = timed(lambda: ap.spectral_contrast(y, **kw), max(2, args.iters // 4))
        # algorithmic: every feature reads the clip once and writes T floats; the current two-launch form also
        # writes and re-reads the (B, T, F) complex spectrum (8FT + passes x 8FT bytes)
        report("f1", f"spectral centroid+bandwidth+rolloff+flatness 22.05k 2048/512 {B}x30 s (4 calls from audio)", B * 30.0, ms,
               4 * B * (4 * L + 4 * T), 4 * B * T * (fft_flops(N) + N + 8 * F),
               {"ms_spectral_contrast_extra": ms_contrast})
    if "f2" in want:  # section 8(f) rank 2: rms + zero-crossing rate + pre-emphasis of the same batch
        B, L = max(1, int(1024 * sc * 0.125)), 661500
        y = clips(B, L, 22050)
        T = 1 + L // 512

        def tdom():
            ap.rms(y, 2048, 512); ap.zero_crossing_rate(y, 2048, 512); ap.preemphasis(y)
        ms = timed(tdom, args.iters)
        report("f2", f"rms+zcr (2048/512)+preemphasis 22.05k {B}x30 s (3 calls)", B * 30.0, ms,
               B * (2 * (4 * L + 4 * T) + 8 * L), B * (2 * 2 * 2048 * T + 2 * L))
    if "f3" in want:  # section 8(f) rank 3 / 4: autocorrelation pitch detector and delta features
        B, L = max(1, int(1024 * sc * 0.125)), 661500
        y = clips(B, L, 22050)
        T = 1 + L // 512
        ms = timed(lambda: ap.pitch_detect_acf(y, sr=22050), args.iters)
        report("f3", f"pitch_detect_acf 2048/512 (two 4096-point transforms per frame) 22.05k {B}x30 s", B * 30.0, ms,
               B * (4 * L + 5 * T), B * T * (2 * fft_flops(4096) + 3 * 2049 + 2 * 2048))
        M = torch.randn((B, 40, 2584), device="cuda")
        ms_d = timed(lambda: ap.delta(M), args.iters)
        report("f3b", f"delta width 9 on ({B}, 40, 2584) MFCCs", B * 60.0, ms_d, 2 * M.numel() * 4, 2 * 9 * M.numel())
    if "f4" in want:  # section 8(f) ranks 3 / 4, second halves: whole-signal transforms and time-domain filters
        B, L = max(1, int(64 * sc)), 480000
        y = clips(B, L, 16000)
        M = lambda n: 1 << max(0, (2 * n - 2)).bit_length()
        n_out = int(round(L * 22050 / 16000))
        ms = timed(lambda: ap.resample(y, 16000, 22050), max(2, args.iters // 2))
        passes = lambda m: (m.bit_length() - 1 + 4) // 5
        by = B * 8 * 2 * (2 * passes(M(L)) * M(L) + 2 * passes(M(n_out)) * M(n_out))  # pass traffic, not compulsory bytes
        report("f4a", f"resample fft 16k -> 22.05k {B}x30 s (two chirp-z transforms; bytes = pass traffic)", B * 30.0, ms, by,
               B * 5 * (2 * M(L) * math.log2(M(L)) + 2 * M(n_out) * math.log2(M(n_out))))
        ms = timed(lambda: ap.resample_poly(y, 147, 160), max(2, args.iters // 2))
        n_poly = -(-L * 147 // 160)
        report("f4b", f"resample_poly 147/160 {B}x30 s", B * 30.0, ms, B * 4 * (L + n_poly), B * n_poly * 2 * (20 * 160 + 1) / 147)
        yp = ap.preemphasis(y)
        ms = timed(lambda: ap.deemphasis(yp), args.iters)
        report("f4c", f"deemphasis {B}x30 s (scan of the recurrence, float64 inside)", B * 30.0, ms, B * 8 * L, B * L * 6)
        ms = timed(lambda: ap.autocorrelation(y, max_lag=1000), args.iters)
        report("f4d", f"autocorrelation 1000 lags {B}x30 s (direct)", B * 30.0, ms, B * 4 * (L + 1000), B * 2.0 * L * 1000)
        ms = timed(lambda: ap.autocorrelation(y), args.iters)
        report("f4e", f"autocorrelation all lags {B}x30 s (two transforms; bytes = pass traffic)", B * 30.0, ms,
               B * 8 * 2 * 2 * passes(M(L)) * M(L), B * 5 * 2 * M(L) * math.log2(M(L)))
        ms = timed(lambda: ap.periodicity(y, sr=16000), args.iters)
        T = 1 + L // 512
        report("f4f", f"periodicity 2048/512 {B}x30 s", B * 30.0, ms, B * (4 * L + 4 * T), B * T * (2 * fft_flops(4096) + 3 * 2049 + 2 * 2048))


if __name__ == "__main__":
    main()
    if WORLD > 1:
        torch.distributed.destroy_process_group()
