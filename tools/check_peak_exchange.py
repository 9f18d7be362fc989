"""Multi-GPU check of the peer-memory peak exchange (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_peak_exchange.py

Every rank computes the log-mel of ITS shard through LogMelPlan with sharding enabled and compares it, bit
for bit, with its slice of the unsharded computation (melspectrogram -> power_to_db over the whole batch on
one GPU).  Loud clips live on one rank and quiet ones on another, so a rank that used its local peak would
fail.  Prints which path carried the peak (peer memory or the NCCL fallback) and the step time of both."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mlx_audio_primitives_b200 as ap

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
ap.distributed.enable()

B, L = 6 * world, 160 * 64 * 7 + 321
rng = np.random.default_rng(0)
y = rng.standard_normal((B, L)).astype(np.float32)
for b in range(B):  # a different level per shard: the peak lives on the last rank
    y[b] *= 10.0 ** (-3 + 3.0 * (b * world // B) / max(world - 1, 1))
y[1, : L // 3] = 0.0
yt = torch.from_numpy(y).to(dev)
lo, hi = ap.distributed.shard_bounds(B, rank, world)
ok = True
for kw, ref, top_db in [(dict(sr=16000, n_fft=400, hop_length=160, n_mels=80), 1.0, 80.0),
                        (dict(sr=16000, n_fft=400, hop_length=160, n_mels=80), "max", 80.0),
                        (dict(sr=22050, n_fft=1024, hop_length=256, n_mels=64), 1.0, 30.0),
                        (dict(sr=22050, n_fft=2048, hop_length=512, n_mels=128), "max", None)]:
    ap.distributed.disable()
    want = ap.power_to_db(ap.melspectrogram(yt, **kw), ref=(torch.max if ref == "max" else ref), top_db=top_db)[lo:hi]
    ap.distributed.enable()
    plan = ap.LogMelPlan(hi - lo, L, ref=ref, top_db=top_db, **kw)
    for it in range(3):
        got = plan(yt[lo:hi].contiguous())
        same = torch.equal(got, want)
        ok &= same
    if rank == 0:
        print(f"{kw['n_fft']}/{kw['hop_length']} ref={ref} top_db={top_db}: "
              f"{'peer-memory exchange' if plan.xchg is not None else 'NCCL all-reduce'}, bits equal: {same}", flush=True)

# step time with and without the exchange on the benchmark shape
Bb, Lb = 64, 480000
yb = torch.randn((Bb, Lb), device=dev)
for use in (True, False):
    plan = ap.LogMelPlan(Bb, Lb, sr=16000, n_fft=400, hop_length=160, n_mels=80)
    if not use:
        plan.xchg = None
    out = plan.empty_output()
    for _ in range(20):
        plan(yb, out)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(500):
        plan(yb, out)
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 500 * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"step, {'peer-memory exchange' if (use and plan.xchg is not None) else 'NCCL all-reduce'}: {float(t):.1f} us (max over {world} ranks)", flush=True)
flag = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if float(flag) == 1.0 else 1)
