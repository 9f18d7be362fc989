#!/usr/bin/env python
"""What the box gives N processes that only copy: the ceiling of bench.py's `e2e` figure.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_h2d_ceiling.py

Every rank owns one GPU and moves the bytes of one e2e step of the Whisper log-mel workload -- 122.9 MB of clips
host -> device, 61.5 MB of log-mel device -> host -- with plain pinned `cudaMemcpyAsync` copies (one per direction
per step, nothing batched, no kernels), first each direction alone, then both at once on two streams.  Done twice:
with the process scheduled wherever the OS puts it, and with the rank pinned to its own slice of the host cores
BEFORE its pinned buffers are allocated and first touched (NUMA-local pages where the box has more than one node).
Barrier on both sides, max over ranks.  Rank 0 prints one JSON line per mode: per-rank GB/s, aggregate GB/s and the
audio-seconds/second an e2e path that hid every kernel behind the copies would reach."""
from __future__ import annotations

import json
import os
import sys
import time

import torch
import torch.distributed as dist

H2D_BYTES, D2H_BYTES, AUDIO_S = 64 * 480000 * 4, 64 * 80 * 3001 * 4, 64 * 30.0

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def max_over_ranks(x: float) -> float:
    t = torch.tensor([x], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def measure(mode: str, iters: int = 20):
    if mode == "pinned_cpu_slice":
        cpus = sorted(os.sched_getaffinity(0))
        k = max(1, len(cpus) // world)
        os.sched_setaffinity(0, set(cpus[rank * k:(rank + 1) * k]) or set(cpus))
    h_in = torch.empty(H2D_BYTES // 4, dtype=torch.float32).pin_memory()
    h_in.fill_(1.0)  # first touch on the (possibly pinned) CPU slice
    h_out = torch.empty(D2H_BYTES // 4, dtype=torch.float32).pin_memory()
    h_out.fill_(0.0)
    d_in = torch.empty(H2D_BYTES // 4, dtype=torch.float32, device=dev)
    d_out = torch.zeros(D2H_BYTES // 4, dtype=torch.float32, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def h2d():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)

    out = {}
    for name, fns in (("h2d", (h2d,)), ("d2h", (d2h,)), ("both", (h2d, d2h))):
        for f in fns:
            f()
        barrier()
        t0 = time.perf_counter()
        for _ in range(iters):
            for f in fns:
                f()
        torch.cuda.synchronize()
        out[name + "_ms"] = max_over_ranks((time.perf_counter() - t0) / iters * 1e3)
        barrier()
    if mode == "pinned_cpu_slice":
        os.sched_setaffinity(0, set(range(os.cpu_count() or 1)))
    if rank == 0:
        line = {"mode": mode, "n_procs": world, **out,
                "h2d_GBs_per_rank": H2D_BYTES / out["h2d_ms"] / 1e6, "d2h_GBs_per_rank": D2H_BYTES / out["d2h_ms"] / 1e6,
                "both_aggregate_GBs": world * (H2D_BYTES + D2H_BYTES) / out["both_ms"] / 1e6,
                "e2e_ceiling_audio_s_per_s": world * AUDIO_S / (out["both_ms"] * 1e-3),
                "host_cpus": os.cpu_count()}
        print(json.dumps(line), flush=True)


for m in ("default", "pinned_cpu_slice"):
    measure(m)
if world > 1:
    dist.destroy_process_group()
sys.exit(0)
