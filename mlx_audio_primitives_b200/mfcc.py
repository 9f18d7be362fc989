"""DCT-II and MFCC (reference ``mfcc.py``)."""
from __future__ import annotations

import threading
from functools import lru_cache

import numpy as np
import torch

from . import distributed
from ._extension import _ext, check
from ._tensor import f32c, ptr, require_cuda, stream_ptr, to_tensor
from ._validation import validate_positive
from .mel import _melspec_from_bank, _resolve_stft_args, check_band_args, mel_filterbank_host, sparse_bank_device


@lru_cache(maxsize=32)
def dct_matrix_host(n_out: int, n_in: int, norm) -> np.ndarray:
    """D[k, n] = cos(pi*k*(2n+1)/(2N)) in float64, orthonormal scaling when norm == 'ortho',
    rounded to float32 (reference mfcc.py:24-66)."""
    n = np.arange(n_in, dtype=np.float64)
    k = np.arange(n_out, dtype=np.float64)[:, None]
    D = np.cos(np.pi * k * (2 * n + 1) / (2 * n_in))
    if norm == "ortho":
        D[0] *= 1.0 / np.sqrt(n_in)
        D[1:] *= np.sqrt(2.0 / n_in)
    D = D.astype(np.float32)
    D.setflags(write=False)
    return D


_lock = threading.RLock()
_dct_device: dict[tuple, torch.Tensor] = {}


def dct_matrix(n_out: int, n_in: int, norm="ortho") -> torch.Tensor:
    """Device-resident (n_out, n_in) DCT-II basis (reference ``get_dct_matrix`` dct.cpp:24)."""
    require_cuda()
    key = (int(n_out), int(n_in), norm, torch.cuda.current_device())
    with _lock:
        t = _dct_device.get(key)
        if t is None:
            t = torch.from_numpy(np.array(dct_matrix_host(int(n_out), int(n_in), norm))).cuda()
            _dct_device[key] = t
        return t


def dct(x, type: int = 2, n: int | None = None, axis: int = -1, norm: str | None = "ortho") -> torch.Tensor:
    """DCT-II along ``axis`` as a product with the cached basis (reference mfcc.py:69-140)."""
    if type != 2:
        raise ValueError(f"Only DCT type 2 is supported, got {type}")
    x = to_tensor(x, torch.float32)
    size = x.shape[axis]
    if n is None:
        n = size
    moved = x.movedim(axis, -1).contiguous()
    rows = moved.numel() // size
    D = dct_matrix(n, size, norm)
    out = torch.empty(moved.shape[:-1] + (n,), dtype=torch.float32, device=x.device)
    if rows:
        check(_ext.mlxa_dct_f32(ptr(moved), rows, size, ptr(D), n, ptr(out), stream_ptr(x)), "dct")
    return out.movedim(-1, axis)


def _lifter_device(n_mfcc: int, lifter, device) -> torch.Tensor | None:
    if lifter <= 0:
        return None
    k = np.arange(n_mfcc)
    lift = 1 + (lifter / 2.0) * np.sin(np.pi * (k + 1) / lifter)  # reference mfcc.py:277-282
    return torch.from_numpy(lift.astype(np.float32)).to(device)


def mfcc(y=None, sr: int = 22050, S=None, n_mfcc: int = 20, dct_type: int = 2, norm: str | None = "ortho",
         lifter: float = 0, n_fft: int = 2048, hop_length: int = 512, win_length: int | None = None,
         window="hann", center: bool = True, pad_mode: str = "constant", power: float = 2.0, n_mels: int = 128,
         fmin: float = 0.0, fmax: float | None = None, htk: bool = False, mel_norm: str | None = "slaney"):
    """MFCC (n_mfcc, T) / (B, n_mfcc, T) (reference mfcc.py:143-287) in two kernels: the fused mel
    kernel (which also leaves the batch peak on the device) and one tail kernel doing
    power_to_db(ref=1, amin=1e-10, top_db=80) -> DCT-II over the mel axis -> lifter."""
    validate_positive(n_mfcc, "n_mfcc")
    if dct_type != 2:
        raise ValueError(f"Only DCT type 2 is supported, got {dct_type}")
    if S is None:
        hop, win_length = _resolve_stft_args(n_fft, hop_length, win_length)
        fmax_ = check_band_args(n_mels, "n_mels", fmin, fmax, sr)
        key = ("mel", sr, n_fft, n_mels, float(fmin), float(fmax_), bool(htk), mel_norm)
        bank = sparse_bank_device(key, lambda: mel_filterbank_host(sr, n_fft, n_mels, float(fmin), float(fmax_),
                                                                   bool(htk), mel_norm))
        M = _melspec_from_bank(y, bank, n_fft, hop, win_length, window, center, pad_mode, power)
        from . import _peaks
        peak = _peaks.lookup(M)
        if distributed.is_enabled():
            peak = distributed.all_reduce_max_(peak.clone())
        apply_db = 1
    else:
        M = f32c(S)  # caller-supplied log-power mel spectrogram: no dB step (reference mfcc.py:229,257)
        peak, apply_db = None, 0
    batched = M.ndim == 3
    if not batched:
        M = M[None]
    M = M.contiguous()
    B, n_in, T = M.shape
    D = dct_matrix(n_mfcc, n_in, norm)
    lift = _lifter_device(n_mfcc, lifter, M.device)
    out = torch.empty((B, n_mfcc, T), dtype=torch.float32, device=M.device)
    for b0 in range(0, B, 65535):
        nb = min(65535, B - b0)
        check(_ext.mlxa_mfcc_tail_f32(ptr(M[b0:]), nb, n_in, T, ptr(D), n_mfcc, ptr(lift), apply_db, 1e-10, 1.0,
                                      apply_db, 80.0, ptr(peak), ptr(out[b0:]), stream_ptr(M)), "mfcc")
    return out if batched else out[0]


# ---------------------------------------------------------------- delta features (reference mfcc.py:290-371)
_SG_MODES = {"interp": 0, "nearest": 1, "mirror": 2, "constant": 3, "wrap": 4}


@lru_cache(maxsize=64)
def savgol_operators_host(width: int, polyorder: int, deriv: int, delta: float):
    """Savitzky-Golay taps (correlation order) and the (width//2, width) edge operators of mode 'interp', float64 ->
    float32.  Restates scipy.signal.savgol_coeffs / _fit_edges_polyfit (the reference's host dependency,
    scipy >= 1.10): least-squares polynomial of degree `polyorder` through `width` samples, its `deriv`-th
    derivative at the centre (interior) or at the first / last width//2 positions of the edge windows."""
    from math import factorial
    if polyorder >= width:
        raise ValueError("polyorder must be less than window_length.")
    if deriv > polyorder:
        taps = np.zeros(width)
    else:
        h = width // 2
        x = np.arange(-h, width - h, dtype=np.float64)[::-1]
        A = x[None, :] ** np.arange(polyorder + 1)[:, None]
        y = np.zeros(polyorder + 1)
        y[deriv] = factorial(deriv) / (delta ** deriv)
        taps = np.linalg.lstsq(A, y, rcond=None)[0][::-1]  # 'conv' coefficients reversed = correlation taps
    h = width // 2
    t = np.arange(width, dtype=np.float64)
    P = np.linalg.pinv(t[:, None] ** np.arange(polyorder + 1)[None, :])  # polynomial coefficients = P @ samples

    def deriv_rows(pos):
        D = np.zeros((len(pos), polyorder + 1))
        for k in range(deriv, polyorder + 1):
            D[:, k] = factorial(k) / factorial(k - deriv) * pos ** (k - deriv)
        return D / (delta ** deriv)
    left = deriv_rows(np.arange(0, h, dtype=np.float64)) @ P
    right = deriv_rows(np.arange(width - h, width, dtype=np.float64)) @ P
    return (np.ascontiguousarray(taps, np.float32), np.ascontiguousarray(left, np.float32),
            np.ascontiguousarray(right, np.float32))


def delta(data, width: int = 9, order: int = 1, axis: int = -1, mode: str = "interp", **kwargs) -> torch.Tensor:
    """Delta (derivative) features by Savitzky-Golay filtering along ``axis`` (reference mfcc.py:290-371, which
    bounces to scipy.signal.savgol_filter on the host): one kernel on the device, taps and edge operators computed
    once per (width, order) on the host.  ``polyorder`` (default ``order``), ``delta`` and ``cval`` are honoured."""
    validate_positive(width, "width")
    validate_positive(order, "order")
    if width < 3:
        raise ValueError(f"width must be >= 3, got {width}")
    if width % 2 == 0:
        raise ValueError(f"width must be odd, got {width}")
    if mode not in _SG_MODES:
        raise ValueError("mode must be 'mirror', 'constant', 'nearest' 'wrap' or 'interp'.")
    x = to_tensor(data, torch.float32)
    if x.ndim == 0:
        x = x.reshape(1)
    n = x.shape[axis]
    if mode == "interp" and width > n:
        raise ValueError(f"when mode='interp', width={width} cannot exceed data.shape[axis]={n}")
    kwargs.pop("deriv", None)
    polyorder = int(kwargs.pop("polyorder", order))
    spacing = float(kwargs.pop("delta", 1.0))
    cval = float(kwargs.pop("cval", 0.0))
    if kwargs:
        raise TypeError(f"unexpected keyword arguments: {sorted(kwargs)}")
    taps, left, right = savgol_operators_host(int(width), polyorder, int(order), spacing)
    moved = x.movedim(axis, -1)
    xc = moved.contiguous()
    rows = xc.numel() // n if n else 0
    out = torch.empty_like(xc)
    if rows and n:
        dev = xc.device
        t_d, l_d, r_d = (torch.from_numpy(a).to(dev) for a in (taps, left, right))
        check(_ext.mlxa_savgol_f32(ptr(xc), rows, n, ptr(t_d), int(width), _SG_MODES[mode], cval, ptr(l_d), ptr(r_d), ptr(out),
                                   stream_ptr(xc)), "savgol")
    return out.movedim(-1, axis)
