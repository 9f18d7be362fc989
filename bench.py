#!/usr/bin/env python
"""Benchmark of the spectral hot path (contract: one JSON line on stdout from rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3]

Workload c2 (default, BASELINE.json configs[1]): Whisper-style log-mel, 16 kHz, n_fft 400, hop 160,
80 mels, 64 clips x 30 s per GPU, power_to_db(ref=1.0, amin=1e-10, top_db=80) -- one "step" is one
pass over one such batch.  Weak scaling: every rank owns its own 64 clips; the only exchange is the
one-float all-reduce(MAX) that top_db's batch-global max needs.
Metric: audio-seconds per second.  `value` times the step with inputs resident in HBM;
`e2e` times the same step through the C-ABI host-buffer entry point (pinned host in, host out).
`--impl reference` times the CPU restatement of the reference path (oracle/) on the host cores.
The same line also carries `configs` (one row per other BASELINE.json config: c1 stft+istft, c3 music mel +
dB(ref=max) at 128 clips per GPU -- the 1024-clip batch at 8 GPUs --, c4 MFCC at 256 clips, c5 Griffin-Lim),
`sustained` (the headline step looped for >= 2 s with its own clock record) and, under torchrun,
`sharded_parity` (every rank's shard bit-equal to the unsharded public API fed the batch-global peak).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (sr, n_fft, hop, n_mels, clips per GPU, seconds per clip)
    "c2": dict(sr=16000, n_fft=400, hop=160, n_mels=80, clips=64, seconds=30.0,
               label="Whisper-style log-mel: 16 kHz, n_fft=400 hop=160 n_mels=80, batch 64x30 s"),
    "c3": dict(sr=22050, n_fft=2048, hop=512, n_mels=128, clips=128, seconds=30.0,
               label="Music mel+power_to_db: 22.05 kHz, n_fft=2048 hop=512 n_mels=128, 128x30 s per GPU"),
}
FP32_NOMINAL_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12  # 74.4


def algorithmic_cost(w, T, L):
    """bytes = 4L + 4*n_mels*T per clip; flops per frame = 2.5 N log2 N (rFFT) + N (window)
    + 3F (|X|^2) + 4F (band-sparse mel) + 3*n_mels (dB)."""
    import math
    N, F = w["n_fft"], w["n_fft"] // 2 + 1
    bytes_clip = 4 * L + 4 * w["n_mels"] * T
    flops_frame = 2.5 * N * math.log2(N) + N + 3 * F + 4 * F + 3 * w["n_mels"]
    return bytes_clip, flops_frame * T


# ------------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline (the only place bench.py executes oracle/)
# ------------------------------------------------------------------------------------------------
def synth_clips_np(n, L, sr, seed0=42):
    import numpy as np
    t = np.arange(L) / sr
    base = np.sin(2 * np.pi * (100 + 1000 * t) * t)
    out = np.empty((n, L), np.float32)
    for i in range(n):  # reference benchmarks/utils.py:92-115, one seed per clip
        out[i] = base + 0.1 * np.random.default_rng(seed0 + i).standard_normal(L)
    return out


def cpu_reference_pass(y, w, threads):
    """One pass of the reference's CPU path (float32 restatement, oracle/spectral.py) over clips y,
    clips spread over `threads` host threads; the dB clamp uses the batch-global max like the reference."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    from oracle import spectral as o
    try:
        from threadpoolctl import threadpool_limits
    except Exception:  # pragma: no cover
        threadpool_limits = None

    def mel_one(i):
        return o.melspectrogram(y[i], sr=w["sr"], n_fft=w["n_fft"], hop_length=w["hop"], n_mels=w["n_mels"])

    def run():
        with ThreadPoolExecutor(threads) as ex:
            mels = list(ex.map(mel_one, range(y.shape[0])))
            peak = max(float(m.max()) for m in mels)
            floor = float(o.power_to_db(np.array([peak], np.float32), top_db=None)[0]) - 80.0
            return list(ex.map(lambda m: np.maximum(o.power_to_db(m, top_db=None), floor), mels))

    t0 = time.perf_counter()
    if threadpool_limits is not None and threads > 1:
        with threadpool_limits(limits=1):
            out = run()
    else:
        out = run()
    return time.perf_counter() - t0, out


def cpu_baseline(w, threads, clips):
    L = int(w["sr"] * w["seconds"])
    y = synth_clips_np(clips, L, w["sr"])
    cpu_reference_pass(y[: max(1, min(2, clips))], w, threads)  # warm caches / thread pool
    dt, _ = cpu_reference_pass(y, w, threads)
    return {"value": clips * w["seconds"] / dt, "unit": "audio-s/s", "cores": threads, "kind": "port",
            "sample": f"{clips} clips x {w['seconds']:.0f} s of the same workload, 1 pass, float32 NumPy/pocketfft "
                      f"restatement of the reference CPU path (MLX not installable), {threads} host threads"}


def run_reference_arm(args, w):
    """The reference's CPU path (float32 restatement; MLX is not installable here) on this host's cores: the full
    batch per step, exactly --steps timed steps after --warmup untimed ones.  Under torchrun rank 0 alone runs it:
    the number describes ONE host, whatever --gpus says."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    L = int(w["sr"] * w["seconds"])
    clips = w["clips"]
    y = synth_clips_np(clips, L, w["sr"])
    warm, steps = max(0, args.warmup), max(1, args.steps)
    for _ in range(warm):
        cpu_reference_pass(y, w, threads)
    t = [cpu_reference_pass(y, w, threads)[0] for _ in range(steps)]
    dt = sum(t) / len(t)
    val = clips * w["seconds"] / dt
    line = {"impl": "reference", "metric": "audio-seconds per second, log-mel", "value": val, "unit": "audio-s/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["label"], "clips_per_step": clips,
                       "scope": "one CPU process on rank 0 using every host core; the same number whatever --gpus is"},
            "cpu_baseline": {"value": val, "unit": "audio-s/s", "cores": threads, "kind": "port",
                             "sample": f"{clips} clips x {w['seconds']:.0f} s per step (the full per-GPU batch), float32 "
                                       f"NumPy/pocketfft restatement of the reference CPU path (MLX not installable)"},
            "e2e": {"value": val, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.thr = [], None, None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        def pump():
            for line in self.proc.stdout:
                self.rows.append((time.perf_counter(), line.strip()))
        self.thr = threading.Thread(target=pump, daemon=True)
        self.thr.start()

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0, t1):
        rows = [r for (t, r) in self.rows if t0 <= t <= t1]
        note = None
        if len(rows) < 3:
            rows = [r for (_, r) in self.rows]
            note = "timed region shorter than the sampling period; summary covers warm-up + timed + e2e steps"
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            p = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except Exception:
                continue
            for n, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        out = {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
               "reasons": sorted(reasons), "samples": len(sm)}
        if note:
            out["note"] = note
        return out


# ------------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The ONE JSON line of the contract, on the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # anything a library prints to fd 1 (NCCL's version banner under torchrun) must not pollute the JSON line
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap_ = argparse.ArgumentParser()
    ap_.add_argument("--gpus", type=int, default=1)
    ap_.add_argument("--steps", type=int, default=2000)
    ap_.add_argument("--warmup", type=int, default=20)
    ap_.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap_.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap_.add_argument("--e2e-steps", type=int, default=10)
    ap_.add_argument("--no-cpu-baseline", action="store_true")
    ap_.add_argument("--no-configs", action="store_true", help="skip the rows of the other BASELINE configs")
    ap_.add_argument("--sustained-seconds", type=float, default=2.0, help="0 skips the sustained loop")
    args = ap_.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference_arm(args, w)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import mlx_audio_primitives_b200 as ap
    from mlx_audio_primitives_b200._extension import _ext, check

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        ap.distributed.enable()
    K, W = max(1, args.steps), max(3, args.warmup)

    B, L, sr = w["clips"], int(w["sr"] * w["seconds"]), w["sr"]
    plan = ap.LogMelPlan(B, L, sr=sr, n_fft=w["n_fft"], hop_length=w["hop"], n_mels=w["n_mels"],
                         ref=1.0, amin=1e-10, top_db=80.0)
    T = plan.T
    # synthetic clips (chirp + 0.1*noise, reference benchmarks/utils.py:92-115), generated on the device;
    # NBUF distinct batches are rotated so every step reads inputs that are not L2-resident.
    in_bytes = B * L * 4
    NBUF = max(2, int(np.ceil(3 * 126e6 / in_bytes)) + 1)
    g = torch.Generator(device=dev); g.manual_seed(42 + rank)
    t = torch.arange(L, device=dev, dtype=torch.float64) / sr
    base = torch.sin(2 * np.pi * (100 + 1000 * t) * t).to(torch.float32)
    ys = [base[None, :] + 0.1 * torch.randn((B, L), generator=g, device=dev, dtype=torch.float32) for _ in range(NBUF)]
    outs = [plan.empty_output() for _ in range(2)]
    del t

    def step(i, ev=None):
        y, out = ys[i % NBUF], outs[i % 2]
        if ev is not None:
            ev[0].record()
        plan.mel(y, out)
        if ev is not None:
            ev[1].record()
        plan.db(out)
        return out

    # ---- sharded parity on the record (world > 1): the ranks hold clips of DIFFERENT levels, so a rank that used
    # its local peak for top_db would get other bits.  Expected shard = the unsharded public API
    # (melspectrogram -> power_to_db, reference convert.py:42-58) over this rank's clips plus one extra row that
    # carries the batch-global peak of the raw mel values (gathered over NCCL). -------------------------------
    sharded_parity = None
    if world > 1:
        yp = ys[0][:8] * (10.0 ** (-3.0 + 3.0 * rank / (world - 1)))
        plan_p = ap.LogMelPlan(yp.shape[0], L, sr=sr, n_fft=w["n_fft"], hop_length=w["hop"], n_mels=w["n_mels"],
                               ref=1.0, amin=1e-10, top_db=80.0)
        got = [plan_p(yp).clone() for _ in range(2)][-1]
        ap.distributed.disable()
        mel = ap.melspectrogram(yp, sr=sr, n_fft=w["n_fft"], hop_length=w["hop"], n_mels=w["n_mels"])
        peaks = [torch.zeros(1, device=dev) for _ in range(world)]
        dist.all_gather(peaks, mel.max().reshape(1))
        gpeak = torch.stack(peaks).max()
        want = ap.power_to_db(torch.cat([mel, gpeak.expand(1, mel.shape[1], mel.shape[2])]), ref=1.0, amin=1e-10,
                              top_db=80.0)[:-1]
        local_only = ap.power_to_db(mel, ref=1.0, amin=1e-10, top_db=80.0)
        ap.distributed.enable()
        flag = torch.tensor([1.0 if torch.equal(got, want) else 0.0,
                             1.0 if torch.equal(local_only, want) else 0.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        sharded_parity = {"bit_equal_on_every_rank": bool(flag[0].item() == 1.0),
                          "test_is_sensitive": bool(flag[1].item() == 0.0),  # some rank's local-peak result differs
                          "clips_per_rank": int(yp.shape[0]), "levels": "1e-3 .. 1 across ranks",
                          "exchange": "peer memory" if plan_p.xchg is not None else "NCCL all-reduce"}
        del plan_p, got, want, mel, local_only
        if not sharded_parity["bit_equal_on_every_rank"]:
            if rank == 0:
                emit({"error": "sharded parity failed", "sharded_parity": sharded_parity})
            dist.destroy_process_group()
            sys.exit(1)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    for i in range(W):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start = time.perf_counter()
    e0.record()
    for i in range(K):
        step(i, evs[i])
    e1.record()
    torch.cuda.synchronize()
    t_end = time.perf_counter()
    if world > 1:
        dist.barrier()
    ms_total = e0.elapsed_time(e1)
    kern_ms = [a.elapsed_time(b) for a, b in evs]
    tt = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_total = float(tt.item())
    ms_step = ms_total / K
    audio_s = B * w["seconds"] * world
    value = audio_s / (ms_step * 1e-3)

    # ---- end to end through the C-ABI host-buffer entry point (pinned host in / out) -----------
    KE = max(1, min(args.e2e_steps, K))
    yh = [torch.empty((B, L), dtype=torch.float32).pin_memory() for _ in range(2)]
    for h in yh:
        h.copy_(ys[0])
    oh = torch.empty((B, plan.n_mels, T), dtype=torch.float32).pin_memory()
    plan.run_host(yh[0], oh)  # warm-up (workspace allocation)
    plan.run_host(yh[1], oh)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    te0 = time.perf_counter()
    for i in range(KE):
        plan.run_host(yh[i % 2], oh)
    te = (time.perf_counter() - te0) / KE
    tt = torch.tensor([te], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    te = float(tt.item())
    e2e_value = audio_s / te
    # sanity: the e2e result equals the resident-path result for the same input
    plan.run_host(yh[0], oh)
    ref_out = plan(ys[0])
    torch.cuda.synchronize()
    # (under torchrun both paths use the batch-global peak: the host-buffer entry takes the same peer-memory exchange)
    same = torch.tensor([1.0 if torch.equal(ref_out.cpu(), oh) else 0.0], device=dev)
    if world > 1:
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
    e2e_match = bool(same.item() == 1.0)
    # the platform's copy ceiling for the same bytes: one plain pinned H2D of the clips and one D2H of the result
    # on two streams, nothing else -- what a perfect overlap of copies and kernels could reach on this box
    d_in, d_out = torch.empty((B, L), device=dev), plan.empty_output()
    s_h2d, s_d2h = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    def copies():
        with torch.cuda.stream(s_h2d):
            d_in.copy_(yh[0], non_blocking=True)
        with torch.cuda.stream(s_d2h):
            oh.copy_(d_out, non_blocking=True)
    copies()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    tc0 = time.perf_counter()
    for _ in range(KE):
        copies()
    torch.cuda.synchronize()
    tcopy = torch.tensor([(time.perf_counter() - tc0) / KE], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tcopy, op=dist.ReduceOp.MAX)
    copy_ceiling_ms = float(tcopy.item()) * 1e3
    del d_in, d_out

    # ---- sustained: the same step looped for >= 2 s (a burst of 20 steps runs at boost clocks; this is what a
    # serving loop gets), with its own clock record ----------------------------------------------------------
    sustained = None
    if args.sustained_seconds > 0:
        n_sus = max(K, int(args.sustained_seconds / (ms_step * 1e-3)) + 1)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts0 = time.perf_counter()
        s0.record()
        for i in range(n_sus):
            step(i)
        s1.record()
        torch.cuda.synchronize()
        ts1 = time.perf_counter()
        tt = torch.tensor([s0.elapsed_time(s1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        sus_ms = float(tt.item()) / n_sus
        sustained = {"steps": n_sus, "seconds": float(tt.item()) * 1e-3, "ms_per_step": sus_ms,
                     "value": audio_s / (sus_ms * 1e-3), "unit": "audio-s/s",
                     "clocks": sampler.summary(ts0, ts1) if sampler else None}

    # ---- the other BASELINE configs, through the public API (tools/bench_configs.py) ------------------------
    config_rows = None
    if not args.no_configs:
        del ys, outs, yh, oh
        torch.cuda.empty_cache()
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_configs
        config_rows = bench_configs.collect(["c1", "c3", "c4", "c5"], iters=10)

    if sampler:
        sampler.stop()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- FP32 peak probe (MEASURED_PEAKS.json has HBM and bf16 only) -----------------------------
    scratch = torch.zeros(4, device=dev)
    s_ptr = torch.cuda.current_stream().cuda_stream
    blocks, threads, iters = 148 * 8, 256, 20000
    check(_ext.mlxa_ffma_probe(scratch.data_ptr(), blocks, threads, iters, s_ptr), "probe")
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(5):
        p0.record(); check(_ext.mlxa_ffma_probe(scratch.data_ptr(), blocks, threads, iters, s_ptr), "probe"); p1.record()
        torch.cuda.synchronize()
        best = min(best, p0.elapsed_time(p1))
    fp32_measured = blocks * threads * iters * 16 / (best * 1e-3) / 1e12

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks \
        else (6650.0, "fallback (B200_PROFILING.md)")
    bytes_clip, flops_clip = algorithmic_cost(w, T, L)
    k_ms = statistics.mean(kern_ms)
    achieved = bytes_clip * B / (k_ms * 1e-3) / 1e9
    fl_ach = flops_clip * B / (k_ms * 1e-3) / 1e12
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": None, "kernel": ("mel_rows_kernel" if w["n_fft"] == 400 else "fwd_kernel<EP_MEL>") + f" n_fft={w['n_fft']}", "kernel_ms": k_ms,
                "kernel_ms_min": min(kern_ms), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bytes_clip * B, "algorithmic_flops_per_launch": flops_clip * B,
                "fp32": {"achieved_tflops": fl_ach, "peak_measured_tflops": fp32_measured,
                         "peak_nominal_tflops": FP32_NOMINAL_TFLOPS, "frac_of_measured": fl_ach / fp32_measured,
                         "frac_of_nominal": fl_ach / FP32_NOMINAL_TFLOPS},
                "t_roof_ms": max(bytes_clip * B / (hbm_peak * 1e9), flops_clip * B / (fp32_measured * 1e12)) * 1e3,
                "frac_of_roofline_time": max(bytes_clip * B / (hbm_peak * 1e9), flops_clip * B / (fp32_measured * 1e12)) * 1e3 / k_ms}
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        roofline["traffic"] = tj.get(args.workload)
        roofline["traffic_source"] = tj.get("_source")  # an ncu --set full capture of this kernel (a profiler cannot run inside the bench)
    except Exception:
        pass

    line = {"metric": "audio-seconds per second, log-mel", "value": value, "unit": "audio-s/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["label"], "clips_per_gpu": B, "samples_per_clip": L, "frames_per_clip": T,
                       "power_to_db": "ref=1.0 amin=1e-10 top_db=80 (batch-global max)",
                       "parallelism": (f"clips sharded x{world}, one-float MAX exchange " +
                                       ("over peer memory inside the mel / floor kernels" if plan.xchg is not None
                                        else "by NCCL all-reduce")) if world > 1 else "single GPU",
                       "l2": f"{NBUF} distinct input batches rotated ({NBUF * in_bytes / 1e6:.0f} MB > 126 MB L2)"},
            "roofline": roofline,
            "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": in_bytes,
                    "d2h_bytes_per_step": B * plan.n_mels * T * 4, "ms_per_step": te * 1e3, "steps": KE,
                    "api": "LogMelPlan.run_host -> mlxa_logmel_host_f32 (pinned host in/out, chunked copy/compute overlap)",
                    "matches_resident_path": e2e_match,
                    "copy_ceiling_ms": copy_ceiling_ms, "frac_of_copy_ceiling": copy_ceiling_ms / (te * 1e3),
                    "copy_ceiling": "plain pinned cudaMemcpyAsync of the same bytes (H2D clips + D2H result on two streams, "
                                    "no kernels), max over ranks"},
            "gpu_launches": K * plan.kernel_launches_per_call,
            "clocks": sampler.summary(t_start, t_end) if sampler else None,
            "sustained": sustained, "configs": config_rows}
    if sharded_parity is not None:
        line["sharded_parity"] = sharded_parity
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(w, os.cpu_count() or 1, min(B, 64))
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
