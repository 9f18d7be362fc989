"""End-to-end log-mel (BASELINE config c2) through the host-buffer C-ABI entry point: pinned host clips in,
pinned host result out, H2D / kernels / D2H inside the timed region.  MLXA_HOST_CHUNKS sets the overlap
granularity.  Also prints the plain pinned H2D / D2H copy times of the same buffers (the PCIe floor)."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mlx_audio_primitives_b200 as ap

B, L, sr = 64, 480000, 16000
TOP = None if os.environ.get("E2E_NO_TOPDB") else 80.0
plan = ap.LogMelPlan(B, L, sr=sr, n_fft=400, hop_length=160, n_mels=80, top_db=TOP)
t = np.arange(L) / sr
base = np.sin(2 * np.pi * (100 + 1000 * t) * t).astype(np.float32)
rng = np.random.default_rng(42)
yh = [torch.from_numpy(base[None] + 0.1 * rng.standard_normal((B, L)).astype(np.float32)).pin_memory() for _ in range(2)]
oh = torch.empty((B, 80, plan.T), dtype=torch.float32).pin_memory()
for i in range(3):
    plan.run_host(yh[i % 2], oh)
torch.cuda.synchronize()
n = 20
t0 = time.perf_counter()
for i in range(n):
    plan.run_host(yh[i % 2], oh)
dt = (time.perf_counter() - t0) / n
d = torch.empty((B, L), device="cuda"); o = torch.empty((B, 80, plan.T), device="cuda")
def timed(fn):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / 10
h2d = timed(lambda: d.copy_(yh[0], non_blocking=True)); d2h = timed(lambda: oh.copy_(o, non_blocking=True))
print(f"top_db {TOP} chunks {os.environ.get('MLXA_HOST_CHUNKS', 'default')}: e2e {dt * 1e3:.3f} ms = {B * 30 / dt / 1e6:.3f} M audio-s/s; "
      f"plain H2D {h2d * 1e3:.3f} ms ({yh[0].numel() * 4 / h2d / 1e9:.1f} GB/s), plain D2H {d2h * 1e3:.3f} ms ({oh.numel() * 4 / d2h / 1e9:.1f} GB/s)")
