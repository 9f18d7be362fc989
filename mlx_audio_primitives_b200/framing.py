"""Framing (reference ``framing.py:16-78`` + ``_frame_impl.py:18-82``)."""
from __future__ import annotations

import torch

from ._extension import _ext, check
from ._tensor import f32c, ptr, stream_ptr
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
