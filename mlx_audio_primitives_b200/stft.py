"""STFT / ISTFT (reference ``stft.py``) on the fused sm_100a kernels.

``stft`` is one kernel (pad -> frame -> window -> real FFT); ``istft`` is one kernel (inverse real
FFT -> window -> overlap-add -> normalise -> trim) plus a cached window-sum-of-squares envelope.
Physical layout of spectra is (B, T, F); the public result is the transposed view (B, F, T),
exactly the logical shape/values of the reference (which also returns a transposed view of a
(B, T, F) buffer, stft.py:216, and transposes back first thing in istft, stft.py:292).
"""
from __future__ import annotations

import threading

import numpy as np
import torch

from ._extension import _ext, check
from ._tensor import dense_like, f32c, ptr, publish, stream_ptr, to_tensor
from .mel import _resolve_stft_args, frames_or_raise, pad_mode_code
from .windows import get_window, padded_window

_WINDOW_SUM_EPSILON = 1e-8  # applied inside the kernels (reference stft.py:20)


def _stft_physical(y2d: torch.Tensor, n_fft: int, hop: int, win: torch.Tensor, center: bool, pad_mode: str):
    """(B, L) float32 -> (B, T, F) complex64, physical layout."""
    mode = pad_mode_code(pad_mode)
    B, L = y2d.shape
    T = frames_or_raise(L, n_fft, hop, center, pad_mode)
    F = n_fft // 2 + 1
    out = torch.empty((B, T, F, 2), dtype=torch.float32, device=y2d.device)
    if B == 0:  # an empty batch gives an empty result (PyTorch semantics), not an error
        return torch.view_as_complex(out)
    check(_ext.mlxa_stft_f32(ptr(y2d), B, L, y2d.stride(0), ptr(win), n_fft, hop, int(center), mode, ptr(out),
                             stream_ptr(y2d)), "stft")
    return torch.view_as_complex(out)


def stft(y, n_fft: int = 2048, hop_length: int | None = None, win_length: int | None = None, window="hann",
         center: bool = True, pad_mode: str = "constant") -> torch.Tensor:
    """Short-time Fourier transform, complex64 (F, T) / (B, F, T) (reference stft.py:136-222)."""
    hop, win_length = _resolve_stft_args(n_fft, hop_length, win_length)
    y = f32c(y)
    one_d = y.ndim == 1
    if one_d:
        y = y[None, :]
    if y.ndim != 2:
        raise ValueError(f"y must be 1D or 2D, got {y.ndim}D")
    win = padded_window(window, win_length, n_fft)
    S = _stft_physical(y, n_fft, hop, win, center, pad_mode).transpose(1, 2)
    return S[0] if one_d else S


def _spectrum_physical(S: torch.Tensor) -> torch.Tensor:
    """logical (B, F, T) complex64 -> contiguous physical (B, T, F) (zero-copy when S came from stft)."""
    if S.dtype != torch.complex64:
        S = S.to(torch.complex64)
    P = S.transpose(1, 2)
    if P.is_contiguous():
        return P
    S = S.contiguous()
    B, F, T = S.shape
    out = torch.empty((B, T, F), dtype=torch.complex64, device=S.device)
    check(_ext.mlxa_transpose_c64(ptr(S), B, F, T, ptr(out), stream_ptr(S)), "transpose")
    return out


_wss_lock = threading.RLock()
_wss_cache: dict[tuple, torch.Tensor] = {}


def _window_sumsquare(win: torch.Tensor, n_fft: int, hop: int, T: int, ola_len: int) -> torch.Tensor:
    key = (win.data_ptr(), win._version, n_fft, hop, T, ola_len, win.device.index)
    with _wss_lock:
        hit = _wss_cache.get(key)
        if hit is not None:
            return hit[0]
        w = torch.empty(ola_len, dtype=torch.float32, device=win.device)
        check(_ext.mlxa_window_sumsquare_f32(ptr(win), n_fft, hop, T, ola_len, ptr(w), stream_ptr(win)),
              "window_sumsquare")
        if len(_wss_cache) >= 32:
            _wss_cache.pop(next(iter(_wss_cache)))
        _wss_cache[key] = (publish(w), win)  # keep the window alive so its pointer cannot be recycled
        return w


def _istft_geometry(T: int, n_fft: int, hop: int, center: bool, length):
    """(ola_len, trim, out_len) following reference stft.py:300-338."""
    if length is not None:
        ola_len = length + n_fft if center else length
    else:
        ola_len = n_fft + (T - 1) * hop
    if center:
        trim = n_fft // 2
        out_len = length if length is not None else max(ola_len - 2 * trim, 0)
    else:
        trim = 0
        out_len = length if length is not None else ola_len
    return ola_len, trim, out_len


def _istft_physical(P: torch.Tensor, n_fft: int, hop: int, win: torch.Tensor, center: bool, length,
                    out: torch.Tensor | None = None, u_prev: torch.Tensor | None = None,
                    momentum: float = 0.0, u_out: torch.Tensor | None = None) -> torch.Tensor:
    """physical (B, T, F_in) complex64 -> (B, out_len) float32.  With ``u_out`` / ``u_prev`` (Griffin-Lim) the
    kernel also stores u = istft(P) and returns u + momentum*(u - u_prev): the momentum step of
    griffinlim.py:176-178 taken through the (linear) inverse transform."""
    B, T, F_in = P.shape
    ola_len, trim, out_len = _istft_geometry(T, n_fft, hop, center, length)
    if out_len <= 0 or ola_len <= 0:
        return torch.zeros((B, 0), dtype=torch.float32, device=P.device)
    wss = _window_sumsquare(win, n_fft, hop, T, ola_len)
    if out is None:
        out = torch.empty((B, out_len), dtype=torch.float32, device=P.device)
    if B == 0:
        return out
    if u_out is not None or (u_prev is not None and momentum != 0.0):
        check(_ext.mlxa_istft_momentum_f32(ptr(P), ptr(u_prev), float(momentum), ptr(u_out), B, T, F_in, ptr(win), ptr(wss),
                                           n_fft, hop, ola_len, trim, out_len, ptr(out), out.stride(0), stream_ptr(P)),
              "istft")
        return out
    check(_ext.mlxa_istft_f32(ptr(P), B, T, F_in, ptr(win), ptr(wss), n_fft, hop, ola_len, trim, out_len,
                              ptr(out), out.stride(0), stream_ptr(P)), "istft")
    return out


def istft(stft_matrix, hop_length: int | None = None, win_length: int | None = None, n_fft: int | None = None,
          window="hann", center: bool = True, length: int | None = None) -> torch.Tensor:
    """Inverse STFT, float32 (samples,) / (B, samples) (reference stft.py:225-344)."""
    S = to_tensor(stft_matrix)
    if S.ndim not in (2, 3):
        raise ValueError(f"stft_matrix must be 2D or 3D, got {S.ndim}D")
    two_d = S.ndim == 2
    if two_d:
        S = S[None]
    F = S.shape[1]
    if n_fft is None:
        n_fft = 2 * (F - 1)
    if hop_length is None:
        hop_length = n_fft // 4
    if win_length is None:
        win_length = n_fft
    win = padded_window(window, win_length, n_fft)
    y = _istft_physical(_spectrum_physical(S), int(n_fft), int(hop_length), win, center, length)
    return y[0] if two_d else y


def magnitude(stft_matrix) -> torch.Tensor:
    """|S| with the shape/strides of the input (reference stft.py:347-362)."""
    S = to_tensor(stft_matrix)
    if not S.is_complex():
        return S.abs()
    S, out = dense_like(S.to(torch.complex64), torch.float32)
    if S.numel():
        check(_ext.mlxa_magnitude_f32(ptr(S), S.numel(), ptr(out), stream_ptr(S)), "magnitude")
    return out


def phase(stft_matrix) -> torch.Tensor:
    """atan2(im, re) (reference stft.py:365-379)."""
    S = to_tensor(stft_matrix)
    S, out = dense_like(S.to(torch.complex64), torch.float32)
    if S.numel():
        check(_ext.mlxa_phase_f32(ptr(S), S.numel(), ptr(out), stream_ptr(S)), "phase")
    return out


def check_nola(window, hop_length: int, n_fft: int, tol: float = 1e-10) -> bool:
    """Nonzero-overlap-add test on the host (reference stft.py:382-431)."""
    w = get_window(window, n_fft, True).detach().cpu().numpy().astype(np.float32)
    sums = np.zeros(hop_length, dtype=np.float32)
    for s in range(n_fft // hop_length):
        sums += w[s * hop_length:(s + 1) * hop_length] ** 2
    rem = n_fft % hop_length
    if rem:
        sums[:rem] += w[-rem:] ** 2
    return bool(sums.min() > tol)


# ---- reference-granularity primitives (the functions of the nanobind module, bindings.cpp) ----
def pad_signal(signal, pad_length: int, mode: str = "constant") -> torch.Tensor:
    """(B, L) -> (B, L + 2*pad_length) (reference bindings.cpp:77, pad_signal.cpp:133)."""
    x = f32c(signal)
    if x.ndim != 2:
        raise ValueError("signal must be 2D (batch, samples)")
    if pad_length < 0:
        raise ValueError("pad_length must be non-negative")
    code = pad_mode_code(mode)
    if pad_length == 0:
        return x
    B, L = x.shape
    if mode == "reflect" and pad_length > L - 1:
        raise ValueError(f"reflect padding ({pad_length}) requires pad <= signal_length - 1 ({L - 1})")
    out = torch.empty((B, L + 2 * pad_length), dtype=torch.float32, device=x.device)
    check(_ext.mlxa_pad_signal_f32(ptr(x), B, L, pad_length, code, ptr(out), stream_ptr(x)), "pad_signal")
    return out


def overlap_add(frames, window, hop_length: int, output_length: int) -> torch.Tensor:
    """(B, T, n_fft) raw frames -> (B, output_length) (reference bindings.cpp:16, overlap_add.cpp:195)."""
    fr = f32c(frames)
    if fr.ndim != 3:
        raise ValueError("frames must be 3D (batch, n_frames, n_fft)")
    B, T, n_fft = fr.shape
    w = f32c(window)
    if w.ndim != 1 or w.shape[0] != n_fft:
        raise ValueError("window length must match n_fft")
    if hop_length <= 0:
        raise ValueError("hop_length must be positive")
    if output_length <= 0:
        raise ValueError("output_length must be positive")
    out = torch.empty((B, output_length), dtype=torch.float32, device=fr.device)
    check(_ext.mlxa_overlap_add_f32(ptr(fr), ptr(w), B, T, n_fft, hop_length, output_length, ptr(out),
                                    stream_ptr(fr)), "overlap_add")
    return out
