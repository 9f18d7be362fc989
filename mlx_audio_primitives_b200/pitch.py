"""Pitch detection by autocorrelation (reference ``pitch.py:118-260``).

The reference loops over frames in NumPy on the host (rfft -> |Y|^2 -> irfft -> peak picking per frame); here
one kernel does it for every frame of the batch with two passes through the shared-memory FFT engine and
writes only (f0, voiced).  Frames up to 2048 samples (transform sizes up to 4096) are served.  The whole-signal
``autocorrelation`` (ONE zero-padded FFT of the entire signal in the reference -- not a shared-memory transform) is the
direct, deterministic sum over the requested lags instead: O(n * max_lag), exact to float64 accumulation."""
from __future__ import annotations

import torch

from ._extension import _ext, check
from ._tensor import f32c, ptr, stream_ptr
from ._validation import validate_positive


_DIRECT_MAX_LAGS = 1024        # beyond this many lags two global-memory transforms beat the direct sum
_FFT_WORK_BYTES = 2 << 30


def autocorrelation(y, max_lag: int | None = None, normalize: bool = True, center: bool = True) -> torch.Tensor:
    """r[k] = sum_n y[n] y[n + k] for k < max_lag (default: the signal length), mean removed first when ``center``,
    divided by max(r[0], 1e-10) when ``normalize``; (max_lag,) or (B, max_lag) (reference pitch.py:16-116).
    Up to 1024 lags: the direct sum (float64 across chunks); more: two power-of-two transforms in global memory."""
    y = f32c(y)
    one_d = y.ndim == 1
    if one_d:
        y = y[None, :]
    if y.ndim != 2:
        raise ValueError("signal must be 1-dimensional (samples,) or 2-dimensional (batch, samples)")
    B, n = y.shape
    lag = n if (max_lag is None or max_lag <= 0) else min(int(max_lag), n)
    out = torch.empty((B, lag), dtype=torch.float32, device=y.device)
    if B and lag:
        if lag <= _DIRECT_MAX_LAGS or n > (1 << 23):
            scratch = torch.empty(2 * B, dtype=torch.float32, device=y.device)
            for b0 in range(0, B, 65535):
                nb = min(65535, B - b0)
                check(_ext.mlxa_autocorrelation_f32(ptr(y[b0:]), nb, n, y.stride(0), lag, int(bool(normalize)), int(bool(center)),
                                                    ptr(out[b0:]), ptr(scratch), stream_ptr(y)), "autocorrelation")
        else:
            per_row = _ext.mlxa_autocorrelation_fft_work_bytes(1, n)
            step = int(max(1, min(B, 65535, _FFT_WORK_BYTES // per_row)))
            work = torch.empty(per_row * step, dtype=torch.uint8, device=y.device)
            scratch = torch.empty(2 * step, dtype=torch.float32, device=y.device)
            for b0 in range(0, B, step):
                nb = min(step, B - b0)
                check(_ext.mlxa_autocorrelation_fft_f32(ptr(y[b0:]), nb, n, y.stride(0), lag, int(bool(normalize)), int(bool(center)),
                                                        ptr(out[b0:]), ptr(scratch), ptr(work), work.numel(), stream_ptr(y)),
                      "autocorrelation")
    return out[0] if one_d else out


def pitch_detect_acf(y, sr: int = 22050, fmin: float = 50.0, fmax: float = 2000.0, frame_length: int = 2048,
                     hop_length: int = 512, threshold: float = 0.1, center: bool = True):
    """(f0, voiced_flag): fundamental frequency (Hz, 0 where unvoiced) and a bool mask per frame, (T,) or (B, T).
    f0 = sr / lag of the first local maximum of the normalised autocorrelation above ``threshold`` inside
    [int(sr / fmax), int(sr / fmin)], else of the range's global maximum if that is above the threshold."""
    validate_positive(frame_length, "frame_length")
    validate_positive(hop_length, "hop_length")
    if fmin >= fmax:
        raise ValueError(f"fmin ({fmin}) must be less than fmax ({fmax})")
    if not 33 <= frame_length <= 2048:
        raise ValueError(f"frame_length must be within 33..2048 on the device path, got {frame_length}")
    y = f32c(y)
    one_d = y.ndim == 1
    if one_d:
        y = y[None, :]
    if y.ndim != 2:
        raise ValueError(f"y must be 1D or 2D, got {y.ndim}D")
    B, L = y.shape
    Lp = L + (2 * (frame_length // 2) if center else 0)
    if Lp < frame_length:
        raise ValueError(f"Signal length ({Lp}) must be >= frame_length ({frame_length}). Consider padding the signal.")
    if int(sr / fmin) + 1 > frame_length:
        raise ValueError(f"sr / fmin = {int(sr / fmin)} lags exceed the frame ({frame_length} samples)")
    T = 1 + (Lp - frame_length) // hop_length
    f0 = torch.empty((B, T), dtype=torch.float32, device=y.device)
    voiced = torch.empty((B, T), dtype=torch.uint8, device=y.device)
    check(_ext.mlxa_pitch_acf_f32(ptr(y), B, L, y.stride(0), int(frame_length), int(hop_length), int(center), float(sr),
                                  float(fmin), float(fmax), float(threshold), ptr(f0), ptr(voiced), stream_ptr(y)), "pitch_acf")
    voiced = voiced.to(torch.bool)
    return (f0[0], voiced[0]) if one_d else (f0, voiced)


def periodicity(y, sr: int = 22050, fmin: float = 50.0, fmax: float = 2000.0, frame_length: int = 2048, hop_length: int = 512,
                center: bool = True) -> torch.Tensor:
    """Autocorrelation strength per frame: the maximum of r / r[0] over lags [int(sr / fmax), int(sr / fmin)], 0 for silent
    frames; (1, T) or (B, 1, T) (reference pitch.py:267-383, a per-frame NumPy loop on the host there)."""
    validate_positive(frame_length, "frame_length")
    validate_positive(hop_length, "hop_length")
    if not 33 <= frame_length <= 2048:
        raise ValueError(f"frame_length must be within 33..2048 on the device path, got {frame_length}")
    y = f32c(y)
    one_d = y.ndim == 1
    if one_d:
        y = y[None, :]
    if y.ndim != 2:
        raise ValueError(f"y must be 1D or 2D, got {y.ndim}D")
    B, L = y.shape
    Lp = L + (2 * (frame_length // 2) if center else 0)
    if Lp < frame_length:
        raise ValueError(f"Signal length ({Lp}) must be >= frame_length ({frame_length}). Consider padding the signal.")
    if int(sr / fmin) + 1 > frame_length:
        raise ValueError(f"sr / fmin = {int(sr / fmin)} lags exceed the frame ({frame_length} samples)")
    T = 1 + (Lp - frame_length) // hop_length
    out = torch.empty((B, 1, T), dtype=torch.float32, device=y.device)
    check(_ext.mlxa_periodicity_f32(ptr(y), B, L, y.stride(0), int(frame_length), int(hop_length), int(center), float(sr),
                                    float(fmin), float(fmax), ptr(out), stream_ptr(y)), "periodicity")
    return out[0] if one_d else out
