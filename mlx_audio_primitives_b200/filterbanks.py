"""Linear and Bark triangular filterbanks (reference ``filterbanks.py``): host-built once, then
device-resident through the same caches as the mel bank."""
from __future__ import annotations

from functools import lru_cache

import numpy as np
import torch

from .mel import check_band_args, dense_bank_device, triangular_bank_host


def hz_to_bark(frequencies, formula: str = "zwicker") -> np.ndarray:
    """reference filterbanks.py:17-55"""
    f = np.asarray(frequencies, dtype=np.float64)
    if formula == "zwicker":
        return 13.0 * np.arctan(0.00076 * f) + 3.5 * np.arctan((f / 7500.0) ** 2)
    if formula == "traunmuller":
        z = (26.81 * f) / (1960.0 + f) - 0.53
        z = np.where(z < 2, z + 0.15 * (2 - z), z)
        return np.where(z > 20.1, z + 0.22 * (z - 20.1), z)
    raise ValueError(f"Unknown formula: '{formula}'. Supported: 'zwicker', 'traunmuller'")


def bark_to_hz(bark, formula: str = "zwicker") -> np.ndarray:
    """reference filterbanks.py:58-104 (Newton inversion for Zwicker, closed form for Traunmuller)."""
    z = np.asarray(bark, dtype=np.float64)
    if formula == "zwicker":
        hz = 600.0 * np.sinh(z / 6.0)
        for _ in range(5):
            est = hz_to_bark(hz, "zwicker")
            slope = np.maximum((hz_to_bark(hz + 1e-6, "zwicker") - est) / 1e-6, 1e-10)
            hz = np.maximum(hz - (est - z) / slope, 0)
        return hz
    if formula == "traunmuller":
        z = np.where(z < 2, z - 0.15 * (2 - z) / 1.15, z)
        z = np.where(z > 20.1, z - 0.22 * (z - 20.1) / 1.22, z)
        return 1960.0 * (z + 0.53) / (26.28 - z)
    raise ValueError(f"Unknown formula: '{formula}'. Supported: 'zwicker', 'traunmuller'")


@lru_cache(maxsize=64)
def _linear_host(sr, n_fft, n_bands, fmin, fmax, norm):
    return triangular_bank_host(np.linspace(fmin, fmax, n_bands + 2), sr, n_fft, norm)


@lru_cache(maxsize=64)
def _bark_host(sr, n_fft, n_bands, fmin, fmax, formula, norm):
    lo = hz_to_bark(np.array([fmin]), formula)[0]
    hi = hz_to_bark(np.array([fmax]), formula)[0]
    return triangular_bank_host(bark_to_hz(np.linspace(lo, hi, n_bands + 2), formula), sr, n_fft, norm)


def linear_filterbank(sr: int, n_fft: int, n_bands: int = 64, fmin: float = 0.0, fmax: float | None = None,
                      norm: str | None = "slaney") -> torch.Tensor:
    """(n_bands, n_fft//2 + 1) equal-width bands (reference filterbanks.py:273-342)."""
    fmax = check_band_args(n_bands, "n_bands", fmin, fmax, sr)
    key = ("lin", sr, n_fft, n_bands, float(fmin), float(fmax), norm)
    return dense_bank_device(key, lambda: _linear_host(sr, n_fft, n_bands, float(fmin), float(fmax), norm))


def bark_filterbank(sr: int, n_fft: int, n_bands: int = 24, fmin: float = 0.0, fmax: float | None = None,
                    formula: str = "zwicker", norm: str | None = "slaney") -> torch.Tensor:
    """(n_bands, n_fft//2 + 1) Bark-spaced bands (reference filterbanks.py:159-231)."""
    fmax = check_band_args(n_bands, "n_bands", fmin, fmax, sr)
    key = ("bark", sr, n_fft, n_bands, float(fmin), float(fmax), formula, norm)
    return dense_bank_device(key, lambda: _bark_host(sr, n_fft, n_bands, float(fmin), float(fmax), formula, norm))
