"""Device-side max values remembered for tensors this library produced.

``melspectrogram`` computes max(mel) in its epilogue; ``power_to_db`` needs exactly that number
for ``ref=max`` and for the ``top_db`` clamp (reference convert.py:42-58).  Keeping the scalar on
the device, keyed by the producing tensor's identity and version counter, fuses the reduction
into the producer without changing the two-call public API.
"""
from __future__ import annotations

import weakref

import torch

_MAX_ENTRIES = 16
_table: dict[int, tuple] = {}


def remember(t: torch.Tensor, peak: torch.Tensor) -> None:
    if len(_table) >= _MAX_ENTRIES:
        for k in list(_table)[: _MAX_ENTRIES // 2]:
            _table.pop(k, None)
    key = id(t)
    _table[key] = (weakref.ref(t, lambda _r, k=key: _table.pop(k, None)), t._version, t.data_ptr(), peak)


def lookup(t: torch.Tensor):
    """The remembered peak (1-element device tensor) or None if t is not the unmodified producer output."""
    e = _table.get(id(t))
    if e is None:
        return None
    ref, version, data_ptr, peak = e
    if ref() is t and t._version == version and t.data_ptr() == data_ptr:
        return peak
    return None


def clear() -> None:
    _table.clear()
