"""Window functions (reference ``windows.py``): float64 on the host, rounded once to float32,
cached per (name, length, periodic) on the host and per device as resident tensors."""
from __future__ import annotations

import threading
from functools import lru_cache

import numpy as np
import torch

from ._tensor import publish, require_cuda, to_tensor

# generalized-cosine coefficient sets a0 - a1 cos + a2 cos ... (reference windows.py:63-67)
_COSINE_TERMS = {"hann": (0.5, 0.5), "hamming": (0.54, 0.46), "blackman": (0.42, 0.5, 0.08)}
_CANONICAL = {"hann": "hann", "hanning": "hann", "hamming": "hamming", "blackman": "blackman",
              "bartlett": "bartlett", "triangular": "bartlett", "rectangular": "rectangular",
              "boxcar": "rectangular", "ones": "rectangular"}  # reference windows.py:112-122


@lru_cache(maxsize=128)
def window_host(name: str, n_fft: int, fftbins: bool) -> np.ndarray:
    """float32 host window.  A periodic (DFT-even) window is the symmetric window of n_fft + 1
    points without its last sample (reference windows.py:169-185)."""
    key = name.lower()
    if key not in _CANONICAL:
        raise ValueError(f"Unknown window type: '{key}'. Supported: {', '.join(sorted(_CANONICAL))}")
    kind = _CANONICAL[key]
    n = n_fft + 1 if fftbins else n_fft
    if kind == "rectangular" or n <= 1:
        w = np.ones(n)
    else:
        pos = np.arange(n, dtype=np.float64)
        if kind == "bartlett":
            w = 1 - np.abs(2 * pos / (n - 1) - 1)  # reference windows.py:101-102
        else:
            terms = _COSINE_TERMS[kind]
            w = np.full(n, terms[0])
            for order in range(1, len(terms)):
                w = w + (-1.0) ** order * terms[order] * np.cos(2 * order * np.pi * pos / (n - 1))
            if kind == "blackman":
                w = np.maximum(w, 0.0)  # end points are -1e-17 in float64 (reference windows.py:55-56)
    out = w[:n_fft].astype(np.float32)
    out.setflags(write=False)
    return out


_lock = threading.RLock()
_device_windows: dict[tuple, torch.Tensor] = {}


def _resident(key: tuple, make) -> torch.Tensor:
    with _lock:
        t = _device_windows.get(key)
        if t is None:
            t = publish(make())
            if len(_device_windows) >= 256:
                _device_windows.pop(next(iter(_device_windows)))
            _device_windows[key] = t
        return t


def get_window(window, n_fft: int, fftbins: bool = True) -> torch.Tensor:
    """Window of shape (n_fft,), float32, resident on the current CUDA device
    (reference ``get_window`` windows.py:192-256)."""
    if isinstance(window, str):
        require_cuda()
        dev = torch.cuda.current_device()
        host = window_host(window, int(n_fft), bool(fftbins))
        return _resident(("w", window.lower(), int(n_fft), bool(fftbins), dev),
                         lambda: torch.from_numpy(host.copy()).to(torch.device("cuda", dev)))
    if isinstance(window, (torch.Tensor, np.ndarray)) or hasattr(window, "__dlpack__"):
        w = to_tensor(window)
        if w.shape[0] != n_fft:
            raise ValueError(f"Window array length ({w.shape[0]}) must match n_fft ({n_fft})")
        return w.to(torch.float32)
    raise TypeError(f"window must be str or array/tensor, got {type(window).__name__}")


def padded_window(window, win_length: int, n_fft: int) -> torch.Tensor:
    """Window zero-padded and centred to n_fft: left = (n_fft - win_length) // 2 (reference
    ``_get_padded_window`` stft.py:88-107).  Named windows are cached per device; caller-owned CUDA tensors
    are keyed on (storage pointer, version) -- with the tensor kept alive by the entry -- instead of the
    reference's device->host content hash (stft.py:42), so no synchronising copy happens per call; host
    arrays are uploaded per call and not cached."""
    def build(w: torch.Tensor) -> torch.Tensor:
        if win_length == n_fft:
            return w.contiguous()
        out = torch.zeros(n_fft, dtype=torch.float32, device=w.device)
        left = (n_fft - win_length) // 2
        out[left:left + win_length] = w
        return out

    if isinstance(window, str):
        require_cuda()
        dev = torch.cuda.current_device()
        return _resident(("p", window.lower(), int(win_length), int(n_fft), dev),
                         lambda: build(get_window(window, win_length, True)))
    w = get_window(window, win_length, True)
    if not (isinstance(window, torch.Tensor) and window.is_cuda):
        # NumPy / DLPack / CPU windows are uploaded afresh on every call; the upload is freed on return and the
        # allocator hands its block to the next upload, so a (pointer, version) key would alias different windows
        return build(w)
    # a caller-owned CUDA tensor: key on (pointer, version) and keep the tensor alive in the entry, so the
    # pointer cannot be recycled for another window while the entry exists
    return _resident(("pa", w.data_ptr(), w._version, int(win_length), int(n_fft), w.device.index),
                     lambda: (build(w).clone(), w))[0]


def clear_caches() -> None:
    with _lock:
        _device_windows.clear()
    window_host.cache_clear()
