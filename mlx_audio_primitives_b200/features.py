"""Spectral features and zero-crossing rate (reference ``features.py``).

From audio, a feature is ONE kernel for every n_fft with a compiled plan: the fused STFT with a reduction
epilogue -- |X| stays in the registers of the lane group that produced it and only the (B, T) statistic is
written.  Other sizes take two launches (STFT + a per-frame reduction kernel that reads the PHYSICAL (B, T, F)
complex spectrum directly: |X| is formed on load, no magnitude pass, no transposed copy).  From a pre-computed spectrogram ``S`` (logical (B, F, T), as the reference takes it) the same kernel
runs on its (B, T, F) layout -- zero-copy when ``S`` came from ``magnitude(stft(...))``.  Results have the
reference's shapes: (1, T) for 1-D input, (B, 1, T) for batches ((n_bands + 1, T) / (B, n_bands + 1, T) for
``spectral_contrast``).
"""
from __future__ import annotations

import threading

import numpy as np
import torch

from ._extension import _ext, check
from ._tensor import f32c, ptr, publish, require_cuda, stream_ptr, to_tensor
from .mel import _resolve_stft_args, frames_or_raise, pad_mode_code
from .stft import _stft_physical
from .windows import padded_window

_CENTROID, _BANDWIDTH, _ROLLOFF, _FLATNESS = 0, 1, 2, 3
_lock = threading.Lock()
_freq_cache: dict[tuple, torch.Tensor] = {}


def fft_frequencies_device(sr: int, n_fft: int) -> torch.Tensor:
    """linspace(0, sr/2, n_fft//2 + 1) as float32 on the current device (features.py:19-21), cached."""
    require_cuda()
    key = (int(sr), int(n_fft), torch.cuda.current_device())
    with _lock:
        f = _freq_cache.get(key)
        if f is None:
            f = torch.from_numpy(np.linspace(0, sr / 2.0, n_fft // 2 + 1).astype(np.float32)).cuda()
            _freq_cache[key] = publish(f)
        return f


def _frames_bins(y, S, n_fft, hop_length, win_length, window, center, pad_mode):
    """-> (data (B, T, F) contiguous, complex?, batched?).  features.py:24-54."""
    if S is not None:
        S = to_tensor(S)
        if S.is_complex():
            raise ValueError("S must be a real magnitude spectrogram")
        S = S.to(torch.float32)
        batched = S.ndim == 3
        if not batched:
            if S.ndim != 2:
                raise ValueError(f"S must be 2D or 3D, got {S.ndim}D")
            S = S[None]
        P = S.transpose(1, 2)
        if not P.is_contiguous():  # a plain (B, F, T) array: one transposing copy ((B, T, F) views, e.g. magnitude(stft(...)), pass as they are)
            S = S.contiguous()
            B, F, T = S.shape
            P = torch.empty((B, T, F), dtype=torch.float32, device=S.device)
            if S.numel():
                check(_ext.mlxa_transpose_f32(ptr(S), B, F, T, ptr(P), stream_ptr(S)), "transpose")
        return P, False, batched
    if y is None:
        raise ValueError("Either y (audio) or S (spectrogram) must be provided")
    hop, win_length = _resolve_stft_args(n_fft, hop_length, win_length)
    y = f32c(y)
    batched = y.ndim == 2
    if not batched:
        if y.ndim != 1:
            raise ValueError(f"y must be 1D or 2D, got {y.ndim}D")
        y = y[None, :]
    X = _stft_physical(y, n_fft, hop, padded_window(window, win_length, n_fft), center, pad_mode)
    return X, True, batched


def _fused_from_audio(y, S, n_fft, hop_length, win_length, window, center, pad_mode, sr, freq, kind, p1=0.0, p2=0.0,
                      norm=True, centroid=None):
    """From audio with a compiled plan for n_fft: ONE kernel (fused STFT + per-frame reduction in the lane group
    that produced the spectrum; nothing but the (B, T) result is written).  Returns None when the two-launch
    form has to be used (a spectrogram was supplied, or n_fft has no compiled plan)."""
    if S is not None or y is None or not _ext.mlxa_has_fused_feature(int(n_fft)):
        return None
    hop, win_length = _resolve_stft_args(n_fft, hop_length, win_length)
    y = f32c(y)
    batched = y.ndim == 2
    if not batched:
        if y.ndim != 1:
            raise ValueError(f"y must be 1D or 2D, got {y.ndim}D")
        y = y[None, :]
    B, L = y.shape
    T = frames_or_raise(L, n_fft, hop, center, pad_mode)
    F = n_fft // 2 + 1
    f = _freq(freq, sr, n_fft, y.device)
    if f.numel() != F:
        raise ValueError(f"freq has {f.numel()} entries, the spectrogram has {F} bins")
    c = None
    if centroid is not None:
        c = to_tensor(centroid, torch.float32, y.device).contiguous()
        if c.numel() != B * T:
            raise ValueError(f"centroid has {c.numel()} entries for {B * T} frames")
    win = padded_window(window, win_length, n_fft)
    out = torch.empty((B, 1, T), dtype=torch.float32, device=y.device)
    check(_ext.mlxa_spectral_feature_f32(ptr(y), B, L, y.stride(0), ptr(win), n_fft, hop, int(center), pad_mode_code(pad_mode),
                                         ptr(f), (float(sr) / 2.0 / (F - 1)) if freq is None else 0.0, kind, float(p1), float(p2),
                                         int(norm), ptr(c), ptr(out), stream_ptr(y)),
          "spectral_feature")
    return out if batched else out[0]


def _stat(data, is_complex, batched, freq, kind, p1=0.0, p2=0.0, norm=True, centroid=None):
    B, T, F = data.shape
    if freq.numel() != F:
        raise ValueError(f"freq has {freq.numel()} entries, the spectrogram has {F} bins")
    out = torch.empty((B, 1, T), dtype=torch.float32, device=data.device)
    if B * T:
        check(_ext.mlxa_spectral_stats_f32(ptr(data), int(is_complex), B, T, F, ptr(freq), kind, float(p1), float(p2),
                                           int(norm), ptr(centroid), ptr(out), stream_ptr(data)), "spectral_stats")
    return out if batched else out[0]


def _freq(freq, sr, n_fft, device):
    if freq is None:
        return fft_frequencies_device(sr, n_fft)
    return to_tensor(freq, torch.float32, device).contiguous().reshape(-1)


def spectral_centroid(y=None, sr: int = 22050, S=None, n_fft: int = 2048, hop_length: int = 512,
                      win_length: int | None = None, window="hann", center: bool = True, pad_mode: str = "constant",
                      freq=None) -> torch.Tensor:
    """sum_k f_k S[k] / (sum_k S[k] + 1e-10) per frame (reference features.py:57-134)."""
    r = _fused_from_audio(y, S, n_fft, hop_length, win_length, window, center, pad_mode, sr, freq, _CENTROID)
    if r is not None:
        return r
    data, cplx, batched = _frames_bins(y, S, n_fft, hop_length, win_length, window, center, pad_mode)
    return _stat(data, cplx, batched, _freq(freq, sr, 2 * (data.shape[2] - 1) if S is not None else n_fft, data.device), _CENTROID)


def spectral_bandwidth(y=None, sr: int = 22050, S=None, n_fft: int = 2048, hop_length: int = 512,
                       win_length: int | None = None, window="hann", center: bool = True, pad_mode: str = "constant",
                       freq=None, centroid=None, p: float = 2.0, norm: bool = True) -> torch.Tensor:
    """(sum S |f - centroid|^p / (sum S + 1e-10))^(1/p) (reference features.py:137-271)."""
    if p <= 0:
        raise ValueError(f"p must be positive, got {p}")
    r = _fused_from_audio(y, S, n_fft, hop_length, win_length, window, center, pad_mode, sr, freq, _BANDWIDTH, p1=p, norm=norm,
                          centroid=centroid)
    if r is not None:
        return r
    data, cplx, batched = _frames_bins(y, S, n_fft, hop_length, win_length, window, center, pad_mode)
    B, T, _ = data.shape
    c = None
    if centroid is not None:
        c = to_tensor(centroid, torch.float32, data.device).contiguous()
        if c.numel() != B * T:
            raise ValueError(f"centroid has {c.numel()} entries for {B * T} frames")
    return _stat(data, cplx, batched, _freq(freq, sr, 2 * (data.shape[2] - 1) if S is not None else n_fft, data.device),
                 _BANDWIDTH, p1=p, norm=norm, centroid=c)


def spectral_rolloff(y=None, sr: int = 22050, S=None, n_fft: int = 2048, hop_length: int = 512,
                     win_length: int | None = None, window="hann", center: bool = True, pad_mode: str = "constant",
                     freq=None, roll_percent: float = 0.85, use_cpp: bool = True) -> torch.Tensor:
    """Frequency of the first bin whose cumulative magnitude reaches roll_percent of the frame's total
    (reference features.py:274-360).  ``use_cpp`` is accepted for signature compatibility."""
    if roll_percent < 0.0:
        raise ValueError(f"roll_percent must be >= 0.0, got {roll_percent}")
    if roll_percent > 1.0:
        raise ValueError(f"roll_percent must be <= 1.0, got {roll_percent}")
    r = _fused_from_audio(y, S, n_fft, hop_length, win_length, window, center, pad_mode, sr, freq, _ROLLOFF, p1=roll_percent)
    if r is not None:
        return r
    data, cplx, batched = _frames_bins(y, S, n_fft, hop_length, win_length, window, center, pad_mode)
    return _stat(data, cplx, batched, _freq(freq, sr, 2 * (data.shape[2] - 1) if S is not None else n_fft, data.device),
                 _ROLLOFF, p1=roll_percent)


def spectral_flatness(y=None, S=None, n_fft: int = 2048, hop_length: int = 512, win_length: int | None = None,
                      window="hann", center: bool = True, pad_mode: str = "constant", power: float = 2.0,
                      amin: float = 1e-10) -> torch.Tensor:
    """exp(mean log max(S, amin)) / (mean max(S, amin) + 1e-10) with S = |X|^power when computed from audio; a
    supplied S is used as it is (reference features.py:363-442)."""
    r = _fused_from_audio(y, S, n_fft, hop_length, win_length, window, center, pad_mode, 1, None, _FLATNESS, p1=power, p2=amin)
    if r is not None:
        return r
    data, cplx, batched = _frames_bins(y, S, n_fft, hop_length, win_length, window, center, pad_mode)
    f = fft_frequencies_device(1, 2 * (data.shape[2] - 1))  # unused by this statistic; only its length is checked
    return _stat(data, cplx, batched, f, _FLATNESS, p1=(power if S is None else 1.0), p2=amin)


def contrast_bands_host(freq: np.ndarray, fmin: float, n_bands: int, quantile: float) -> np.ndarray:
    """(n_bands + 1, 3) int32 {first bin, bin count, n_quantile} per octave band by the reference's edge rules
    (features.py:536-565): bins with f_low <= f <= f_high, the neighbour bin below for every band but the first,
    the last band extended to Nyquist, n_quantile from the count before the last bin is dropped."""
    freq = np.asarray(freq, dtype=np.float64)
    octa = np.zeros(n_bands + 2)
    octa[1:] = fmin * (2.0 ** np.arange(0, n_bands + 1))
    rows = []
    for k in range(n_bands + 1):
        inside = np.flatnonzero((freq >= octa[k]) & (freq <= octa[k + 1]))
        if inside.size == 0:
            rows.append((0, 0, 0))
            continue
        lo, hi = int(inside[0]), int(inside[-1])
        if k > 0 and lo > 0:
            lo -= 1
        if k == n_bands:
            hi = len(freq) - 1
        n = hi - lo + 1
        nq = int(max(np.rint(quantile * n), 1))
        if k < n_bands and n > 1:
            n -= 1
        rows.append((lo, n, nq))
    return np.asarray(rows, dtype=np.int32)


def spectral_contrast(y=None, sr: int = 22050, S=None, n_fft: int = 2048, hop_length: int = 512,
                      win_length: int | None = None, window="hann", center: bool = True, pad_mode: str = "constant",
                      freq=None, fmin: float = 200.0, n_bands: int = 6, quantile: float = 0.02,
                      linear: bool = False) -> torch.Tensor:
    """Octave-band peak/valley contrast, (n_bands + 1, T) / (B, n_bands + 1, T) (reference features.py:445-592,
    which does this on the host in NumPy): mean of the top quantile against mean of the bottom quantile of each
    band's magnitudes, as a difference of 10*log10 values or linearly.  One kernel over the physical (B, T, F)
    spectrum; the band edges follow the reference's (librosa's) rules on the host."""
    if n_bands <= 0:
        raise ValueError(f"n_bands must be positive, got {n_bands}")
    if not 0.0 <= quantile <= 1.0:
        raise ValueError(f"quantile must be in [0, 1], got {quantile}" if quantile < 0 else f"quantile must be <= 1.0, got {quantile}")
    data, cplx, batched = _frames_bins(y, S, n_fft, hop_length, win_length, window, center, pad_mode)
    B, T, F = data.shape
    f = _freq(freq, sr, 2 * (F - 1) if S is not None else n_fft, data.device)
    if f.numel() != F:
        raise ValueError(f"freq has {f.numel()} entries, the spectrogram has {F} bins")
    bands = torch.from_numpy(contrast_bands_host(f.cpu().numpy(), float(fmin), int(n_bands), float(quantile))).to(data.device)
    out = torch.empty((B, n_bands + 1, T), dtype=torch.float32, device=data.device)
    if B * T:
        check(_ext.mlxa_spectral_contrast_f32(ptr(data), int(cplx), B, T, F, ptr(bands), n_bands + 1, int(linear), ptr(out),
                                              stream_ptr(data)), "spectral_contrast")
    return out if batched else out[0]


def _frame_stat(y, frame_length, hop_length, center, pad_mode, kind, modes):
    if frame_length <= 0:
        raise ValueError(f"frame_length must be positive, got {frame_length}")
    if hop_length <= 0:
        raise ValueError(f"hop_length must be positive, got {hop_length}")
    y = f32c(y)
    one_d = y.ndim == 1
    if one_d:
        y = y[None, :]
    if y.ndim != 2:
        raise ValueError(f"y must be 1D or 2D, got {y.ndim}D")
    if center and pad_mode not in modes:
        raise ValueError(f"Unknown pad_mode: '{pad_mode}'. Supported: " + ", ".join(f"'{m}'" for m in modes))
    B, L = y.shape
    Lp = L + (2 * (frame_length // 2) if center else 0)
    if Lp < frame_length:
        raise ValueError(f"Signal length ({Lp}) must be >= frame_length ({frame_length}). Consider padding the signal.")
    T = 1 + (Lp - frame_length) // hop_length
    out = torch.empty((B, 1, T), dtype=torch.float32, device=y.device)
    check(_ext.mlxa_frame_stats_f32(ptr(y), B, L, y.stride(0), frame_length, hop_length, int(center),
                                    pad_mode_code(pad_mode) if center else 0, kind, ptr(out), stream_ptr(y)), "frame_stats")
    return out[0] if one_d else out


def zero_crossing_rate(y, frame_length: int = 2048, hop_length: int = 512, center: bool = True, pad_mode: str = "edge",
                       use_mlx: bool = True) -> torch.Tensor:
    """Fraction of sign changes per frame, sign = (x >= 0), the first sample of a frame never counts
    (reference features.py:625-720).  ``use_mlx`` is accepted for signature compatibility."""
    return _frame_stat(y, frame_length, hop_length, center, pad_mode, 1, ("constant", "edge"))
