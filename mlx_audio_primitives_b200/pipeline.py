"""Pre-planned log-mel pipeline: ``power_to_db(melspectrogram(y))`` for a fixed batch shape with
every buffer and constant resident, so one call is two kernel launches (fused mel kernel +
dB pass) plus, when sharding is enabled, the one-float all-reduce between them.

This is the serving-shaped entry point the benchmark drives; results are identical to calling
``melspectrogram`` then ``power_to_db`` (same kernels, same arguments)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import distributed
from ._extension import _ext, check
from ._tensor import ptr, require_cuda
from .mel import (_resolve_stft_args, check_band_args, frames_or_raise, mel_filterbank_host, pad_mode_code,
                  sparse_bank_device)
from .windows import padded_window, window_host


class LogMelPlan:
    def __init__(self, batch: int, length: int, sr: int = 22050, n_fft: int = 2048, hop_length: int | None = None,
                 win_length: int | None = None, window: str = "hann", center: bool = True,
                 pad_mode: str = "constant", power: float = 2.0, n_mels: int = 128, fmin: float = 0.0,
                 fmax: float | None = None, htk: bool = False, norm: str | None = "slaney", ref=1.0,
                 amin: float = 1e-10, top_db: float | None = 80.0, to_db: bool = True, device=None):
        require_cuda()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.B, self.L = int(batch), int(length)
        self.n_fft = int(n_fft)
        self.hop, self.win_length = _resolve_stft_args(n_fft, hop_length, win_length)
        self.center, self.pad_mode, self.mode = bool(center), pad_mode, pad_mode_code(pad_mode)
        self.power, self.n_mels = float(power), int(n_mels)
        fmax = check_band_args(n_mels, "n_mels", fmin, fmax, sr)
        if top_db is not None and top_db <= 0:
            raise ValueError(f"top_db must be positive, got {top_db}")
        self.ref_is_max = ref in ("max", torch.max, torch.amax, np.max, max)
        self.ref = 1.0 if self.ref_is_max else float(ref)
        self.amin, self.top_db, self.to_db = float(amin), top_db, bool(to_db)
        self.T = frames_or_raise(self.L, self.n_fft, self.hop, self.center, pad_mode)
        with torch.cuda.device(self.device):
            key = ("mel", sr, n_fft, n_mels, float(fmin), float(fmax), bool(htk), norm)
            self.bank = sparse_bank_device(key, lambda: mel_filterbank_host(sr, n_fft, n_mels, float(fmin),
                                                                            float(fmax), bool(htk), norm))
            self.win = padded_window(window, self.win_length, self.n_fft)
            self.peaks = torch.zeros(2, dtype=torch.float32, device=self.device)  # two slots, alternated per call
            # per (clip, 64-frame block) minima for the block-wise top_db floor; the floor kernel re-arms them
            self.block_min = torch.full((self.B, -(-self.T // 64)), float("inf"), dtype=torch.float32, device=self.device)
            self._slot = 0
        self._win_host = np.zeros(self.n_fft, np.float32)
        left = (self.n_fft - self.win_length) // 2
        self._win_host[left:left + self.win_length] = window_host(window, self.win_length, True)
        self.need_peak = self.to_db and (self.ref_is_max or self.top_db is not None)
        # sharded batch: the peak crosses GPUs through peer memory inside the two kernels (NCCL if unavailable)
        self.xchg = distributed.PeakExchange.create(self.device) if (self.to_db and (self.ref_is_max or self.top_db is not None)) else None
        # ref a constant: the mel kernel's epilogue writes dB itself and top_db is a read-mostly floor pass
        self.fused_db = self.to_db and not self.ref_is_max
        self.kernel_launches_per_call = 1 + (1 if self.need_peak else 0)
        self._awaiting_db = False  # with a peer exchange every mel() must be followed by its db() (epochs pair up)

    def empty_output(self) -> torch.Tensor:
        return torch.empty((self.B, self.n_mels, self.T), dtype=torch.float32, device=self.device)

    # -- the two launches, separately callable so a benchmark can time the dominant kernel ----
    def mel(self, y: torch.Tensor, out: torch.Tensor) -> None:
        s = torch.cuda.current_stream(self.device).cuda_stream
        fuse = self.fused_db  # the peak slot was zeroed by the previous call's dB / floor kernel
        if self.xchg is not None:
            if self._awaiting_db:
                raise RuntimeError("LogMelPlan.mel() was called twice without db(): with a peer-memory peak exchange the "
                                   "two launches pair up by epoch on every rank")
            self.xchg.next_epoch()
            self._awaiting_db = True
        check(_ext.mlxa_melspec_f32(ptr(y), self.B, self.L, y.stride(0), ptr(self.win), self.n_fft, self.hop,
                                    int(self.center), self.mode, self.power, ptr(self.bank.packed),
                                    self.n_mels, self.bank.n_w4, ptr(out), self._peak_ptr() if self.need_peak else None,
                                    int(fuse), 10.0, self.amin, self.ref,
                                    ptr(self.block_min) if (fuse and self.need_peak) else None,
                                    self.xchg.ref if self.xchg is not None else None, s), "melspectrogram")

    def _peak_ptr(self, other: bool = False) -> int:
        return self.peaks.data_ptr() + 4 * (self._slot ^ int(other))

    def db(self, out: torch.Tensor) -> None:
        if not self.need_peak:
            return
        xr = self.xchg.ref if self.xchg is not None else None
        self._awaiting_db = False
        if xr is None:
            distributed.all_reduce_max_(self.peaks[self._slot:self._slot + 1])
        s = torch.cuda.current_stream(self.device).cuda_stream
        if self.fused_db:
            check(_ext.mlxa_db_floor_blocks_f32(ptr(out), self.B, self.n_mels, self.T, 10.0, self.amin, self.ref,
                                                float(self.top_db), self._peak_ptr(), ptr(self.block_min),
                                                self._peak_ptr(other=True), None, xr, s), "db_floor_blocks")
            self._slot ^= 1
            return
        check(_ext.mlxa_to_db_f32(ptr(out), out.numel(), 10.0, self.amin, self.ref,
                                  self._peak_ptr() if self.ref_is_max else None, int(self.top_db is not None),
                                  float(self.top_db or 0.0), self._peak_ptr(), ptr(out), self._peak_ptr(other=True), xr, s),
              "to_db")
        self._slot ^= 1

    def __call__(self, y: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        if y.shape != (self.B, self.L) or y.dtype != torch.float32 or not y.is_cuda or y.stride(1) != 1:
            raise ValueError(f"expected a float32 CUDA tensor of shape ({self.B}, {self.L})")
        if out is None:
            out = self.empty_output()
        self.mel(y, out)
        self.db(out)
        return out

    # -- host buffers in, host buffers out (H2D / kernels / D2H overlapped inside the library) --
    def run_host(self, y_host: torch.Tensor, out_host: torch.Tensor) -> torch.Tensor:
        """y_host (B, L) and out_host (B, n_mels, T): contiguous float32 CPU tensors (pinned for full
        copy bandwidth).  Synchronous.  With sharding enabled the batch-global max crosses the GPUs like in the
        resident path (peer-memory exchange), so every rank gets the bits of the unsharded computation."""
        if y_host.is_cuda or out_host.is_cuda or not y_host.is_contiguous() or not out_host.is_contiguous():
            raise ValueError("run_host takes contiguous CPU tensors")
        if tuple(y_host.shape) != (self.B, self.L) or tuple(out_host.shape) != (self.B, self.n_mels, self.T):
            raise ValueError("shape mismatch with the plan")
        vp = lambda a: a.ctypes.data_as(C.c_void_p).value
        use_x = self.xchg is not None and self.need_peak
        if not use_x and self.need_peak and distributed.is_enabled():
            raise RuntimeError("run_host needs the peer-memory peak exchange when the batch is sharded "
                               "(symmetric memory unavailable: use the resident path, which falls back to NCCL)")
        if use_x:
            if self._awaiting_db:
                raise RuntimeError("LogMelPlan.mel() is awaiting its db(): finish the pair before run_host()")
            self.xchg.next_epoch()
        with torch.cuda.device(self.device):
            check(_ext.mlxa_logmel_host_f32(y_host.data_ptr(), self.B, self.L, vp(self._win_host), self.n_fft, self.hop,
                                            int(self.center), self.mode, self.power, vp(self.bank.host),
                                            self.n_mels, self.bank.n_w4, int(self.to_db), int(self.ref_is_max),
                                            self.ref, self.amin, int(self.top_db is not None),
                                            float(self.top_db or 0.0), self.xchg.ref if use_x else None,
                                            out_host.data_ptr()), "logmel_host")
        return out_host
