"""Mel scale, triangular filterbanks and the fused mel spectrogram (reference ``mel.py``).

Filterbank matrices are built on the host exactly like the reference builds them (float64
slopes, +1e-10 in the denominators, float32 cast *before* the Slaney area scaling) and kept
resident per device in two forms: the dense (n_bands, F) matrix the public API returns, and the
band-sparse rows (contiguous support per band) that the fused kernel consumes.
"""
from __future__ import annotations

import ctypes as C
import math
import threading
from dataclasses import dataclass
from functools import lru_cache

import numpy as np
import torch

from . import _peaks
from ._extension import _ext, check
from ._tensor import f32c, ptr, publish, require_cuda, stream_ptr
from ._validation import validate_non_negative, validate_positive
from .windows import padded_window

# Slaney auditory-toolbox constants (reference mel.py:25-28)
_LIN_STEP_HZ = 200.0 / 3
_BREAK_HZ = 1000.0
_BREAK_MEL = _BREAK_HZ / _LIN_STEP_HZ
_LOG_STEP = math.log(6.4) / 27.0


def hz_to_mel(frequencies, htk: bool = False) -> np.ndarray:
    """Host NumPy in/out like the reference (mel.py:31-62)."""
    f = np.asarray(frequencies, dtype=np.float64)
    if htk:
        return 2595.0 * np.log10(1.0 + f / 700.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        log_part = _BREAK_MEL + np.log(f / _BREAK_HZ) / _LOG_STEP
    return np.where(f < _BREAK_HZ, f / _LIN_STEP_HZ, log_part)


def mel_to_hz(mels, htk: bool = False) -> np.ndarray:
    """Host NumPy in/out like the reference (mel.py:65-93)."""
    m = np.asarray(mels, dtype=np.float64)
    if htk:
        return 700.0 * (10.0 ** (m / 2595.0) - 1.0)
    return np.where(m < _BREAK_MEL, _LIN_STEP_HZ * m, _BREAK_HZ * np.exp(_LOG_STEP * (m - _BREAK_MEL)))


def triangular_bank_host(edges_hz: np.ndarray, sr: int, n_fft: int, norm) -> np.ndarray:
    """Triangles between consecutive edge triples (reference mel.py:137-165; the same pattern in
    filterbanks.py:126-152, 248-265)."""
    if norm not in ("slaney", None):
        raise ValueError(f"Unknown norm: '{norm}'. Supported: 'slaney', None")
    n_bands = edges_hz.shape[0] - 2
    bin_hz = np.linspace(0, sr / 2.0, 1 + n_fft // 2)[None, :]
    left, mid, right = edges_hz[:-2, None], edges_hz[1:-1, None], edges_hz[2:, None]
    up = (bin_hz - left) / (mid - left + 1e-10)
    down = (right - bin_hz) / (right - mid + 1e-10)
    bank = np.maximum(0, np.minimum(up, down)).astype(np.float32)
    if norm == "slaney":
        bank *= (2.0 / (edges_hz[2:n_bands + 2] - edges_hz[:n_bands]))[:, None]
    return bank


def check_band_args(n_bands, name: str, fmin, fmax, sr):
    validate_positive(n_bands, name)
    validate_non_negative(fmin, "fmin")
    if fmax is None:
        fmax = sr / 2.0
    if fmin >= fmax:
        raise ValueError(f"fmin ({fmin}) must be less than fmax ({fmax})")
    if fmax > sr / 2.0:
        raise ValueError(f"fmax ({fmax}) cannot exceed Nyquist frequency ({sr / 2.0})")
    return fmax


@lru_cache(maxsize=64)
def mel_filterbank_host(sr, n_fft, n_mels, fmin, fmax, htk, norm) -> np.ndarray:
    edges = mel_to_hz(np.linspace(hz_to_mel(fmin, htk), hz_to_mel(fmax, htk), n_mels + 2), htk)
    bank = triangular_bank_host(edges, sr, n_fft, norm)
    bank.setflags(write=False)
    return bank


@dataclass
class SparseBank:
    """A filterbank in the packed band-sparse form the kernels bulk-copy into shared memory
    (layout: include/mlxa_cuda.h, "packed filterbank")."""
    packed: torch.Tensor      # device, float32 words
    n_bands: int
    n_w4: int
    host: np.ndarray          # the same words on the host (for the host-buffer entry point)


def sparse_rows_host(bank: np.ndarray):
    """(start, length, weights-per-row) of each row's contiguous support -- used by the tests to
    show that the packed form reproduces the dense matrix exactly."""
    out = []
    for row in bank:
        nz = np.flatnonzero(row)
        out.append((int(nz[0]), int(nz[-1] - nz[0] + 1), row[nz[0]:nz[-1] + 1].copy()) if nz.size else (0, 0, row[:0]))
    return out


def pack_bank_host(bank: np.ndarray, group: int):
    """dense (n_bands, F) float32 -> (packed words, n_wt) via the library's own packer; ``group`` is the
    lane-group width of the kernels serving this n_fft (``mlxa_plan_group``)."""
    bank = np.ascontiguousarray(bank, dtype=np.float32)
    n_bands, F = bank.shape
    n_wt = C.c_int64(0)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    check(_ext.mlxa_pack_filterbank(vp(bank), n_bands, F, group, None, 0, C.byref(n_wt)), "pack_filterbank")
    words = int(_ext.mlxa_packed_bank_words(n_bands, n_wt.value, group))
    packed = np.zeros(words, dtype=np.float32)
    check(_ext.mlxa_pack_filterbank(vp(bank), n_bands, F, group, vp(packed), words, C.byref(n_wt)), "pack_filterbank")
    return packed, int(n_wt.value)


_lock = threading.RLock()
_dense_cache: dict[tuple, torch.Tensor] = {}
_sparse_cache: dict[tuple, SparseBank] = {}


def dense_bank_device(key: tuple, host_fn) -> torch.Tensor:
    require_cuda()
    k = key + (torch.cuda.current_device(),)
    with _lock:
        t = _dense_cache.get(k)
        if t is None:
            t = torch.from_numpy(np.array(host_fn())).cuda()
            _dense_cache[k] = publish(t)
        return t


def sparse_bank_device(key: tuple, host_fn) -> SparseBank:
    require_cuda()
    k = key + (torch.cuda.current_device(),)
    with _lock:
        sb = _sparse_cache.get(k)
        if sb is None:
            dense = np.asarray(host_fn())
            n_fft = 2 * (dense.shape[1] - 1)
            packed, n_w4 = pack_bank_host(dense, int(_ext.mlxa_plan_group(n_fft)))
            sb = SparseBank(torch.from_numpy(packed).cuda(), dense.shape[0], n_w4, packed)
            _sparse_cache[k] = publish(sb)
        return sb


def mel_filterbank(sr: int, n_fft: int, n_mels: int = 128, fmin: float = 0.0, fmax: float | None = None,
                   htk: bool = False, norm: str | None = "slaney") -> torch.Tensor:
    """(n_mels, n_fft // 2 + 1) float32 filterbank resident on the current device
    (reference mel.py:171-242)."""
    fmax = check_band_args(n_mels, "n_mels", fmin, fmax, sr)
    key = ("mel", sr, n_fft, n_mels, float(fmin), float(fmax), bool(htk), norm)
    return dense_bank_device(key, lambda: mel_filterbank_host(sr, n_fft, n_mels, float(fmin), float(fmax), bool(htk), norm))


def _resolve_stft_args(n_fft, hop_length, win_length):
    if hop_length is None:
        hop_length = n_fft // 4
    if win_length is None:
        win_length = n_fft
    if hop_length <= 0:
        raise ValueError(f"hop_length must be positive, got {hop_length}")
    if win_length <= 0:
        raise ValueError(f"win_length must be positive, got {win_length}")
    if win_length > n_fft:
        raise ValueError(f"win_length ({win_length}) must be <= n_fft ({n_fft})")
    if hop_length > n_fft:
        raise ValueError(f"hop_length ({hop_length}) should typically be <= n_fft ({n_fft})")
    return int(hop_length), int(win_length)


_PAD_MODES = {"constant": 0, "reflect": 1, "edge": 2}


def pad_mode_code(pad_mode: str) -> int:
    if pad_mode not in _PAD_MODES:
        raise ValueError(f"Unknown pad_mode: '{pad_mode}'. Supported: reflect, constant, edge")
    return _PAD_MODES[pad_mode]


def frames_or_raise(L: int, n_fft: int, hop: int, center: bool, pad_mode: str) -> int:
    """Frame count of the fused kernel, with the reference's error behaviour
    (_frame_impl.py:51-61; native reflect check pad_signal.cpp:102-105)."""
    pad = n_fft // 2 if center else 0
    if center and pad_mode == "reflect" and pad > L - 1:
        raise ValueError(f"reflect padding ({pad}) requires pad <= signal_length - 1 ({L - 1})")
    padded = L + 2 * pad
    if padded < n_fft:
        raise ValueError(f"Signal length ({padded}) must be >= frame_length ({n_fft}). Consider padding the signal.")
    return 1 + (padded - n_fft) // hop


def _melspec_from_bank(y, bank: SparseBank, n_fft, hop, win_length, window, center, pad_mode, power,
                       want_peak=True, fused_db=None):
    mode = pad_mode_code(pad_mode)
    y = f32c(y)
    one_d = y.ndim == 1
    if one_d:
        y = y[None, :]
    if y.ndim != 2:
        raise ValueError(f"y must be 1D or 2D, got {y.ndim}D")
    B, L = y.shape
    T = frames_or_raise(L, n_fft, hop, center, pad_mode)
    win = padded_window(window, win_length, n_fft)
    out = torch.empty((B, bank.n_bands, T), dtype=torch.float32, device=y.device)
    if B == 0:  # an empty batch gives an empty result (PyTorch semantics), not an error
        return out
    peak = torch.zeros(1, dtype=torch.float32, device=y.device) if want_peak else None
    db = fused_db or (0, 10.0, 1e-10, 1.0)
    check(_ext.mlxa_melspec_f32(ptr(y), B, L, y.stride(0), ptr(win), n_fft, hop, int(center), mode, float(power),
                                ptr(bank.packed), bank.n_bands, bank.n_w4, ptr(out), ptr(peak), int(db[0]), float(db[1]), float(db[2]),
                                float(db[3]), None, None, stream_ptr(y)), "melspectrogram")
    res = out[0] if one_d else out
    if want_peak:
        _peaks.remember(res, peak)
    return res


def melspectrogram(y, sr: int = 22050, n_fft: int = 2048, hop_length: int | None = None,
                   win_length: int | None = None, window="hann", center: bool = True,
                   pad_mode: str = "constant", power: float = 2.0, n_mels: int = 128, fmin: float = 0.0,
                   fmax: float | None = None, htk: bool = False, norm: str | None = "slaney") -> torch.Tensor:
    """Mel spectrogram (n_mels, T) / (B, n_mels, T) (reference mel.py:245-352) in ONE kernel:
    pad -> frame -> window -> rFFT -> |X|^power -> band-sparse projection; the complex spectrum
    stays on chip.  The kernel also leaves max(mel) on the device, so a following
    ``power_to_db(ref=max)`` / ``top_db`` clamp needs no extra reduction pass."""
    hop, win_length = _resolve_stft_args(n_fft, hop_length, win_length)
    fmax = check_band_args(n_mels, "n_mels", fmin, fmax, sr)
    key = ("mel", sr, n_fft, n_mels, float(fmin), float(fmax), bool(htk), norm)
    bank = sparse_bank_device(key, lambda: mel_filterbank_host(sr, n_fft, n_mels, float(fmin), float(fmax), bool(htk), norm))
    return _melspec_from_bank(y, bank, n_fft, hop, win_length, window, center, pad_mode, power)


def clear_caches() -> None:
    with _lock:
        _dense_cache.clear()
        _sparse_cache.clear()
    mel_filterbank_host.cache_clear()
