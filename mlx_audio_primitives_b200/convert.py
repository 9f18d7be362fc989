"""Decibel conversions (reference ``convert.py``)."""
from __future__ import annotations

import builtins

import numpy as np
import torch

from . import _peaks, distributed
from ._extension import _ext, check
from ._tensor import dense_like, ptr, stream_ptr, to_tensor

_MAX_CALLABLES = {torch.max, torch.amax, np.max, np.amax, builtins.max}


def _global_peak(S: torch.Tensor) -> torch.Tensor:
    """max(S) as a 1-element device tensor: taken from the producing kernel when S is an untouched
    melspectrogram output, otherwise one reduction kernel; all-reduced over ranks when enabled."""
    peak = _peaks.lookup(S)
    if peak is None:
        peak = torch.full((1,), float("-inf"), dtype=torch.float32, device=S.device)
        check(_ext.mlxa_max_f32(ptr(S), S.numel(), ptr(peak), stream_ptr(S)), "max")
    elif distributed.is_enabled():
        peak = peak.clone()
    return distributed.all_reduce_max_(peak)


def _to_db(S, ref, coefficient: float, amin: float, top_db):
    """out = coef*log10(max(S, amin)/max(ref, amin)), then max(out, max(out) - top_db) over the whole
    array (reference convert.py:14-60).  A callable ref sees the unclamped input."""
    S = to_tensor(S, torch.float32)
    if top_db is not None and top_db <= 0:
        raise ValueError(f"top_db must be positive, got {top_db}")
    S, out = dense_like(S)
    if S.numel() == 0:
        return out
    peak = None
    ref_host, ref_dev = 1.0, None
    if callable(ref):
        if ref in _MAX_CALLABLES:
            peak = _global_peak(S)
            ref_dev = peak
        else:
            r = ref(S)
            if isinstance(r, torch.Tensor):
                ref_dev = r.to(device=S.device, dtype=torch.float32).reshape(1)
            else:
                ref_host = float(r)
    else:
        ref_host = float(ref)
    if top_db is not None and peak is None:
        peak = _global_peak(S)
    check(_ext.mlxa_to_db_f32(ptr(S), S.numel(), float(coefficient), float(amin), ref_host, ptr(ref_dev),
                              int(top_db is not None), float(top_db or 0.0), ptr(peak), ptr(out), None, None, stream_ptr(S)),
          "to_db")
    return out


def power_to_db(S, ref=1.0, amin: float = 1e-10, top_db: float | None = 80.0) -> torch.Tensor:
    """10*log10(S/ref) (reference convert.py:63-97)."""
    return _to_db(S, ref, 10.0, amin, top_db)


def amplitude_to_db(S, ref=1.0, amin: float = 1e-5, top_db: float | None = 80.0) -> torch.Tensor:
    """20*log10(S/ref) (reference convert.py:132-166)."""
    return _to_db(S, ref, 20.0, amin, top_db)


def _from_db(S_db, ref: float, div: float):
    S_db = to_tensor(S_db, torch.float32)
    S_db, out = dense_like(S_db)
    if S_db.numel():
        check(_ext.mlxa_from_db_f32(ptr(S_db), S_db.numel(), float(ref), div, ptr(out), stream_ptr(S_db)), "from_db")
    return out


def db_to_power(S_db, ref: float = 1.0) -> torch.Tensor:
    """ref * 10^(S_db/10) (reference convert.py:100-129)."""
    return _from_db(S_db, ref, 10.0)


def db_to_amplitude(S_db, ref: float = 1.0) -> torch.Tensor:
    """ref * 10^(S_db/20) (reference convert.py:169-198)."""
    return _from_db(S_db, ref, 20.0)
