"""Resampling (reference ``resample.py``).

``resample_poly`` (scipy.signal.resample_poly on the host in the reference) and ``resample(res_type="linear")`` run
on the device: the Kaiser low-pass of the polyphase resampler is designed on the host exactly as SciPy does
(firwin, beta 5, 20*max(up, down) + 1 taps, cast to float32, scaled by ``up``) and applied by one kernel.
``resample(res_type="fft")`` (scipy.signal.resample: rfft of the whole signal, spectrum kept / zero-extended, irfft) is
two Bluestein chirp transforms over power-of-two Stockham FFTs in global memory (``bigfft.cu``): any length up to 2^24
samples, chirps and their transforms cached per length on the device."""
from __future__ import annotations

import math
from functools import lru_cache

import numpy as np
import torch

from ._extension import _ext, check
from ._tensor import ptr, stream_ptr, to_tensor
from ._validation import validate_positive


@lru_cache(maxsize=64)
def poly_filter_host(up: int, down: int, n_in: int):
    """(padded taps float32, n_pre_remove, n_out) of scipy.signal.resample_poly(x, up, down) for a length-n_in axis
    (window ('kaiser', 5.0), padtype 'constant'); up / down already reduced by their gcd."""
    max_rate = max(up, down)
    f_c, half_len = 1.0 / max_rate, 10 * max_rate
    numtaps = 2 * half_len + 1
    m = np.arange(numtaps) - 0.5 * (numtaps - 1)
    h = f_c * np.sinc(f_c * m) * np.kaiser(numtaps, 5.0)  # scipy.signal.firwin(numtaps, f_c, window=('kaiser', 5.0))
    h /= h.sum()
    h = h.astype(np.float32)
    h *= up
    n_out = n_in * up
    n_out = n_out // down + bool(n_out % down)
    n_pre_pad = down - half_len % down
    n_pre_remove = (half_len + n_pre_pad) // down
    out_len = lambda len_h: (((n_in - 1) * up + len_h) - 1) // down + 1
    n_post_pad = 0
    while out_len(numtaps + n_pre_pad + n_post_pad) < n_out + n_pre_remove:
        n_post_pad += 1
    taps = np.concatenate([np.zeros(n_pre_pad, np.float32), h, np.zeros(n_post_pad, np.float32)])
    return taps, int(n_pre_remove), int(n_out)


_FFT_WORK_BYTES = 2 << 30  # scratch for the global-memory transforms of resample(res_type="fft"); clips go in slabs


def resample_poly(y, up: int, down: int, axis: int = -1, padtype: str = "constant") -> torch.Tensor:
    """Polyphase resampling by the rational factor up/down (reference resample.py:215-300)."""
    validate_positive(up, "up")
    validate_positive(down, "down")
    if padtype != "constant":
        raise ValueError(f"padtype '{padtype}' is not supported on the device path (only 'constant')")
    g = math.gcd(int(up), int(down))
    up, down = int(up) // g, int(down) // g
    y = to_tensor(y, torch.float32)
    if up == 1 and down == 1:
        return y
    moved = y.movedim(axis, -1).contiguous()
    n_in = moved.shape[-1]
    rows = moved.numel() // n_in if n_in else 0
    taps, pre_remove, n_out = poly_filter_host(up, down, int(n_in))
    out = torch.empty(moved.shape[:-1] + (n_out,), dtype=torch.float32, device=moved.device)
    if rows and n_out:
        h = torch.from_numpy(taps).to(moved.device)
        check(_ext.mlxa_resample_poly_f32(ptr(moved), rows, n_in, ptr(h), taps.size, up, down, pre_remove, n_out, ptr(out),
                                          stream_ptr(moved)), "resample_poly")
    return out.movedim(-1, axis)


def resample(y, orig_sr: int, target_sr: int, res_type: str = "fft", fix: bool = True, scale: bool = False,
             axis: int = -1) -> torch.Tensor:
    """Resample to another rate (reference resample.py:21-212).  ``res_type="fft"`` is the Fourier method of
    scipy.signal.resample (two chirp-z transforms over global-memory FFTs, any length up to 2^24 samples);
    ``"linear"`` blends neighbouring samples."""
    validate_positive(orig_sr, "orig_sr")
    validate_positive(target_sr, "target_sr")
    if orig_sr == target_sr:
        return to_tensor(y)
    if res_type not in ("fft", "linear"):
        raise ValueError(f"Unknown res_type: '{res_type}'. Supported: 'fft', 'linear'")
    y = to_tensor(y, torch.float32)
    moved = y.movedim(axis, -1).contiguous()
    n_in = moved.shape[-1]
    ratio = target_sr / orig_sr
    n_out = int(np.round(n_in * ratio)) if fix else int(np.ceil(n_in * ratio))
    if n_out == n_in:
        return y
    rows = moved.numel() // n_in if n_in else 0
    out = torch.empty(moved.shape[:-1] + (n_out,), dtype=torch.float32, device=moved.device)
    if res_type == "fft":
        if rows and n_out:
            if max(n_in, n_out) > (1 << 24):
                raise ValueError(f"res_type='fft' serves signals of up to 2^24 samples, got {max(n_in, n_out)}")
            flat_in, flat_out = moved.reshape(rows, n_in), out.view(rows, n_out)
            per_row = _ext.mlxa_resample_fft_work_bytes(1, n_in, n_out)
            step = int(max(1, min(rows, 65535, _FFT_WORK_BYTES // per_row)))
            work = torch.empty(per_row * step, dtype=torch.uint8, device=moved.device)
            for r0 in range(0, rows, step):
                nb = min(step, rows - r0)
                check(_ext.mlxa_resample_fft_f32(ptr(flat_in[r0:]), nb, n_in, n_in, n_out, float(ratio) if scale else 1.0,
                                                 ptr(flat_out[r0:]), n_out, ptr(work), work.numel(), stream_ptr(moved)),
                      "resample_fft")
        return out.movedim(-1, axis)
    if rows and n_out:
        check(_ext.mlxa_resample_linear_f32(ptr(moved), rows, n_in, n_out, float(ratio), int(bool(scale)), ptr(out),
                                            stream_ptr(moved)), "resample_linear")
    return out.movedim(-1, axis)
