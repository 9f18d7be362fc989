"""Framing, RMS and pre-emphasis (reference ``framing.py:16-295`` + ``_frame_impl.py:18-82``)."""
from __future__ import annotations

import torch

from ._extension import _ext, check
from ._tensor import f32c, ptr, stream_ptr, to_tensor
from ._validation import validate_positive


def frame_signal_batched(y: torch.Tensor, frame_length: int, hop_length: int) -> torch.Tensor:
    """(B, L) -> (B, T, frame_length), frames[b, t, s] = y[b, t*hop + s]."""
    B, L = y.shape
    if frame_length <= 0:
        raise ValueError(f"frame_length must be positive, got {frame_length}")
    if hop_length <= 0:
        raise ValueError(f"hop_length must be positive, got {hop_length}")
    if L < frame_length:
        raise ValueError(
            f"Signal length ({L}) must be >= frame_length ({frame_length}). Consider padding the signal."
        )
    T = 1 + (L - frame_length) // hop_length
    out = torch.empty((B, T, frame_length), dtype=torch.float32, device=y.device)
    check(_ext.mlxa_frame_signal_f32(ptr(y), B, L, frame_length, hop_length, ptr(out), stream_ptr(y)), "frame_signal")
    return out


def frame(y, frame_length: int, hop_length: int, axis: int = -1) -> torch.Tensor:
    """Overlapping frames (T, frame_length) / (B, T, frame_length) -- transposed relative to
    librosa, like the reference (framing.py:44-46)."""
    validate_positive(frame_length, "frame_length")
    validate_positive(hop_length, "hop_length")
    if axis != -1:
        raise ValueError(f"axis must be -1, got {axis}")
    y = f32c(y)
    one_d = y.ndim == 1
    if one_d:
        y = y[None, :]
    frames = frame_signal_batched(y, frame_length, hop_length)
    return frames[0] if one_d else frames


def rms(y, frame_length: int = 2048, hop_length: int = 512, center: bool = True, pad_mode: str = "constant") -> torch.Tensor:
    """sqrt(mean(frame^2)) per frame -> (1, T) / (B, 1, T) (reference framing.py:81-151); one kernel, the
    frames are formed by index arithmetic on the centre-padded clip."""
    from .features import _frame_stat
    return _frame_stat(y, frame_length, hop_length, center, pad_mode, 0, ("constant", "edge"))


def preemphasis(y, coef: float = 0.97, zi=None, return_zf: bool = False, use_mlx: bool = True):
    """y[n] - coef*y[n-1]; the first sample is y[0] + zi with zi = 2 y[0] - y[1] by default; zf = y[-1]
    (reference framing.py:194-295).  ``use_mlx`` is accepted for signature compatibility."""
    if not 0.0 <= coef <= 1.0:
        raise ValueError(f"coef must be in [0, 1], got {coef}")
    y = f32c(y)
    one_d = y.ndim == 1
    if one_d:
        y = y[None, :]
    if y.ndim != 2:
        raise ValueError(f"y must be 1D or 2D, got {y.ndim}D")
    B, L = y.shape
    z = None
    if zi is not None:
        z = to_tensor(zi, torch.float32, y.device).reshape(-1)
        z = (z if z.numel() == B else z[:1].expand(B)).contiguous()
    out = torch.empty_like(y)
    zf = torch.empty((B, 1), dtype=torch.float32, device=y.device) if return_zf else None
    if B * L:
        check(_ext.mlxa_preemphasis_f32(ptr(y), B, L, y.stride(0), float(coef), ptr(z), ptr(out), ptr(zf), stream_ptr(y)),
              "preemphasis")
    if one_d:
        out = out[0]
        zf = zf[0] if zf is not None else None
    return (out, zf) if return_zf else out


def deemphasis(y, coef: float = 0.97, zi=None, return_zf: bool = False):
    """out[n] = y[n] + coef * out[n-1], the inverse of ``preemphasis``; with zi=None the reference's correction
    -corr * coef^n, corr = ((2 - coef) y[0] - y[1]) / (3 - coef), undoes preemphasis's default initial state; zf is the
    final state scipy.signal.lfilter reports (reference framing.py:298-392, lfilter on the host there)."""
    if not 0.0 <= coef <= 1.0:
        raise ValueError(f"coef must be in [0, 1], got {coef}")
    y = f32c(y)
    one_d = y.ndim == 1
    if one_d:
        y = y[None, :]
    if y.ndim != 2:
        raise ValueError(f"y must be 1D or 2D, got {y.ndim}D")
    B, L = y.shape
    z = None
    if zi is not None:
        z = to_tensor(zi, torch.float32, y.device).reshape(-1)
        z = (z if z.numel() == B else z[:1].expand(B)).contiguous()
    elif L < 2:
        raise ValueError("deemphasis with the default initial state needs at least two samples")
    out = torch.empty_like(y)
    zf = torch.empty((B, 1), dtype=torch.float32, device=y.device)
    if B * L:
        check(_ext.mlxa_deemphasis_f32(ptr(y), B, L, y.stride(0), float(coef), ptr(z), int(zi is None), ptr(out), out.stride(0),
                                       ptr(zf), stream_ptr(y)), "deemphasis")
    if one_d:
        out, zf = out[0], zf[0]
    return (out, zf) if return_zf else out
