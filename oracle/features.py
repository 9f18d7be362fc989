"""CPU oracle for the 'next' rows of SURVEY section 8(f): spectral features and the time-domain
frame statistics that sit either side of the spectral hot path (TEST INFRASTRUCTURE ONLY).

NumPy restatement of reference ``features.py`` (centroid :57, bandwidth :137, rolloff :274,
flatness :363, zero_crossing_rate :625) and ``framing.py`` (rms :81, preemphasis :154-295).
Pinned by tests/golden/reference_features.npz (the reference's own code run on the MLX stand-in).
"""
from __future__ import annotations

import numpy as np

from . import spectral as sp


def fft_frequencies(sr, n_fft):
    """features.py:19-21 : linspace(0, sr/2, n_fft//2 + 1) in float32."""
    return np.linspace(0, sr / 2.0, n_fft // 2 + 1).astype(np.float32)


def _spectrogram(y, S, n_fft, hop_length, win_length, window, center, pad_mode, power=1.0, dtype=np.float32):
    """features.py:24-54 : magnitude STFT (to the given power) unless S is supplied."""
    if S is not None:
        return np.asarray(S).astype(dtype)
    if y is None:
        raise ValueError("Either y (audio) or S (spectrogram) must be provided")
    mag = np.abs(sp.stft(y, n_fft, hop_length, win_length, window, center, pad_mode, dtype)).astype(dtype)
    return np.power(mag, dtype(power)) if power != 1.0 else mag


def _batched(S):
    return (S, True) if S.ndim == 3 else (S[None], False)


def spectral_centroid(y=None, sr=22050, S=None, n_fft=2048, hop_length=512, win_length=None, window="hann",
                      center=True, pad_mode="constant", freq=None, dtype=np.float32):
    """sum_k f_k S[k] / (sum_k S[k] + 1e-10) per frame (features.py:120-134)."""
    S, batched = _batched(_spectrogram(y, S, n_fft, hop_length, win_length, window, center, pad_mode, dtype=dtype))
    f = (fft_frequencies(sr, n_fft) if freq is None else np.asarray(freq)).astype(dtype)
    out = (f[None, :, None] * S).sum(1, keepdims=True) / (S.sum(1, keepdims=True) + dtype(1e-10))
    return out if batched else out[0]


def spectral_bandwidth(y=None, sr=22050, S=None, n_fft=2048, hop_length=512, win_length=None, window="hann",
                       center=True, pad_mode="constant", freq=None, centroid=None, p=2.0, norm=True, dtype=np.float32):
    """(sum S |f - centroid|^p / (sum S + 1e-10))^(1/p) (features.py:226-271)."""
    S, batched = _batched(_spectrogram(y, S, n_fft, hop_length, win_length, window, center, pad_mode, dtype=dtype))
    f = (fft_frequencies(sr, n_fft) if freq is None else np.asarray(freq)).astype(dtype)
    if centroid is None:
        centroid = spectral_centroid(S=S, sr=sr, n_fft=n_fft, freq=f, dtype=dtype)
    centroid = np.asarray(centroid).astype(dtype)
    if centroid.ndim == 2:
        centroid = centroid[None]
    dev = np.abs(f[None, :, None] - centroid)
    w = (S * np.power(dev, dtype(p))).sum(1, keepdims=True)
    out = np.power(w / (S.sum(1, keepdims=True) + dtype(1e-10)), dtype(1.0 / p)) if norm else np.power(w, dtype(1.0 / p))
    return out if batched else out[0]


def spectral_rolloff(y=None, sr=22050, S=None, n_fft=2048, hop_length=512, win_length=None, window="hann",
                     center=True, pad_mode="constant", freq=None, roll_percent=0.85, dtype=np.float32):
    """first bin whose cumulative sum reaches roll_percent of the total (features.py:274-360)."""
    if not 0.0 <= roll_percent <= 1.0:
        raise ValueError(f"roll_percent must be >= 0.0, got {roll_percent}" if roll_percent < 0 else
                         f"roll_percent must be <= 1.0, got {roll_percent}")
    S, batched = _batched(_spectrogram(y, S, n_fft, hop_length, win_length, window, center, pad_mode, dtype=dtype))
    f = (fft_frequencies(sr, n_fft) if freq is None else np.asarray(freq)).astype(np.float32)
    cs = np.cumsum(S, axis=1, dtype=dtype)
    thr = dtype(roll_percent) * cs[:, -1:, :]
    idx = np.minimum((cs < thr).sum(1), S.shape[1] - 1)  # searchsorted(side='left') on a non-decreasing sequence
    out = f[idx][:, None, :].astype(np.float32)
    return out if batched else out[0]


def spectral_flatness(y=None, S=None, n_fft=2048, hop_length=512, win_length=None, window="hann", center=True,
                      pad_mode="constant", power=2.0, amin=1e-10, dtype=np.float32):
    """exp(mean log max(S, amin)) / (mean max(S, amin) + 1e-10), S = |X|^power (features.py:430-442)."""
    S, batched = _batched(_spectrogram(y, S, n_fft, hop_length, win_length, window, center, pad_mode, power, dtype))
    S = np.maximum(S, dtype(amin))
    out = np.exp(np.log(S).mean(1, keepdims=True)) / (S.mean(1, keepdims=True) + dtype(1e-10))
    return out.astype(dtype) if batched else out[0].astype(dtype)


def contrast_bands(freq, fmin, n_bands, quantile):
    """Per octave band k = 0..n_bands: (first bin, bin count after the edge rules, n_quantile) exactly as
    features.py:536-565 (librosa's rules): bins with f_low <= f <= f_high, plus the neighbour below for k > 0,
    the last band extended to Nyquist, n_quantile from the count BEFORE the last bin is dropped (all bands but
    the last drop it when they have more than one bin).  Bands without bins are (0, 0, 0)."""
    freq = np.asarray(freq, dtype=np.float64)
    octa = np.zeros(n_bands + 2)
    octa[1:] = fmin * (2.0 ** np.arange(0, n_bands + 1))
    out = []
    for k, (f_low, f_high) in enumerate(zip(octa[:-1], octa[1:])):
        band = np.logical_and(freq >= f_low, freq <= f_high)
        idx = np.flatnonzero(band)
        if len(idx) == 0:
            out.append((0, 0, 0))
            continue
        if k > 0 and idx[0] > 0:
            band[idx[0] - 1] = True
        if k == n_bands and idx[-1] + 1 < len(band):
            band[idx[-1] + 1:] = True
        n = int(band.sum())
        nq = int(np.maximum(np.rint(quantile * n), 1))
        lo = int(np.flatnonzero(band)[0])
        if k < n_bands and n > 1:
            n -= 1
        out.append((lo, n, nq))
    return out


def spectral_contrast(y=None, sr=22050, S=None, n_fft=2048, hop_length=512, win_length=None, window="hann", center=True,
                      pad_mode="constant", freq=None, fmin=200.0, n_bands=6, quantile=0.02, linear=False, dtype=np.float32,
                      parts=False):
    """peak (mean of the top quantile) against valley (mean of the bottom quantile) of every octave band, as a
    difference of 10*log10 values or linearly (features.py:445-592)."""
    if n_bands <= 0:
        raise ValueError(f"n_bands must be positive, got {n_bands}")
    if not 0.0 <= quantile <= 1.0:
        raise ValueError(f"quantile must be in [0, 1], got {quantile}")
    S, batched = _batched(_spectrogram(y, S, n_fft, hop_length, win_length, window, center, pad_mode, dtype=dtype))
    f = fft_frequencies(sr, n_fft) if freq is None else np.asarray(freq)
    B, F, T = S.shape
    valley = np.zeros((B, n_bands + 1, T), dtype=np.float32)
    peak = np.zeros_like(valley)
    for k, (lo, n, nq) in enumerate(contrast_bands(f, fmin, n_bands, quantile)):
        if n == 0:
            continue
        srt = np.sort(S[:, lo:lo + n, :], axis=1)
        valley[:, k, :] = srt[:, :nq, :].mean(1)
        peak[:, k, :] = srt[:, -nq:, :].mean(1)
    if linear:
        out = peak - valley
    else:
        out = 10.0 * np.log10(np.maximum(peak, 1e-10)) - 10.0 * np.log10(np.maximum(valley, 1e-10))
    out = out.astype(np.float32)
    if parts:  # (contrast, peak, valley, largest magnitude per frame): lets a test judge the conditioning of the dB values
        return out, peak, valley, S.max(1, keepdims=True)
    return out if batched else out[0]


def _padded_frames(y, frame_length, hop_length, center, pad_mode):
    if frame_length <= 0:
        raise ValueError(f"frame_length must be positive, got {frame_length}")
    if hop_length <= 0:
        raise ValueError(f"hop_length must be positive, got {hop_length}")
    y = np.asarray(y, dtype=np.float32)
    one_d = y.ndim == 1
    if one_d:
        y = y[None]
    if center:
        if pad_mode not in ("constant", "edge"):
            raise ValueError(f"Unknown pad_mode: '{pad_mode}'. Supported: 'constant', 'edge'")
        y = sp.pad_signal(y, frame_length // 2, pad_mode)
    return sp.frame_signal(y, frame_length, hop_length), one_d


def rms(y, frame_length=2048, hop_length=512, center=True, pad_mode="constant"):
    """sqrt(mean(frame^2)) -> (B, 1, T) (framing.py:81-151)."""
    fr, one_d = _padded_frames(y, frame_length, hop_length, center, pad_mode)
    out = np.sqrt((fr.astype(np.float32) ** 2).mean(-1))[:, None, :].astype(np.float32)
    return out[0] if one_d else out


def zero_crossing_rate(y, frame_length=2048, hop_length=512, center=True, pad_mode="edge"):
    """mean over the frame of sign(x[i]) != sign(x[i-1]) with sign = (x >= 0); the first sample of a
    frame never counts (features.py:594-720)."""
    fr, one_d = _padded_frames(y, frame_length, hop_length, center, pad_mode)
    s = fr >= 0
    cross = np.concatenate([np.zeros_like(s[..., :1]), s[..., 1:] != s[..., :-1]], axis=-1)
    out = cross.astype(np.float32).mean(-1)[:, None, :].astype(np.float32)
    return out[0] if one_d else out


def preemphasis(y, coef=0.97, zi=None, return_zf=False):
    """y[n] - coef*y[n-1]; first sample y[0] + zi with zi = 2 y[0] - y[1] by default; zf = y[-1]
    (framing.py:154-295, the default ``use_mlx`` path)."""
    if not 0.0 <= coef <= 1.0:
        raise ValueError(f"coef must be in [0, 1], got {coef}")
    y = np.asarray(y, dtype=np.float32)
    one_d = y.ndim == 1
    if one_d:
        y = y[None]
    B = y.shape[0]
    if zi is None:
        z = 2 * y[:, 0:1] - y[:, 1:2]
    else:
        z = np.asarray(zi, dtype=np.float32)
        z = np.broadcast_to(z.reshape(-1, 1) if z.ndim == 1 and z.shape[0] == B else z.reshape(1, -1)[:, :1], (B, 1))
    out = y.copy()
    out[:, 1:] = y[:, 1:] - np.float32(coef) * y[:, :-1]
    out[:, 0:1] = y[:, 0:1] + z
    zf = y[:, -1:]
    if one_d:
        out, zf = out[0], zf[0]
    return (out, zf) if return_zf else out


def delta(data, width=9, order=1, axis=-1, mode="interp", **kwargs):
    """mfcc.py:290-371: the reference validates, then calls scipy.signal.savgol_filter (scipy >= 1.10; 1.18.1 here)
    with deriv = order and polyorder defaulting to order, and casts to float32.  The filter is the third-party
    algorithm itself, so the oracle calls it too (on float64 when a yardstick is wanted)."""
    from scipy.signal import savgol_filter
    if width <= 0:
        raise ValueError(f"width must be positive, got {width}")
    if order <= 0:
        raise ValueError(f"order must be positive, got {order}")
    if width < 3:
        raise ValueError(f"width must be >= 3, got {width}")
    if width % 2 == 0:
        raise ValueError(f"width must be odd, got {width}")
    x = np.atleast_1d(np.asarray(data))
    if mode == "interp" and width > x.shape[axis]:
        raise ValueError(f"when mode='interp', width={width} cannot exceed data.shape[axis]={x.shape[axis]}")
    kwargs.pop("deriv", None)
    kwargs.setdefault("polyorder", order)
    return savgol_filter(x, width, deriv=order, axis=axis, mode=mode, **kwargs).astype(np.float32)


def pitch_detect_acf(y, sr=22050, fmin=50.0, fmax=2000.0, frame_length=2048, hop_length=512, threshold=0.1, center=True,
                     dtype=np.float64, return_acf=False):
    """pitch.py:118-260: per frame, r = irfft(|rfft(frame - mean, n_fft)|^2) with n_fft the power of two >= 2*frame_length - 1,
    normalised by r[0] (frames with r[0] <= 1e-10 are unvoiced); f0 = sr / lag of the FIRST local maximum above
    `threshold` in [sr/fmax, sr/fmin] (lags as int()), else of the global maximum of that range if above threshold.
    -> (f0 float32, voiced bool), (T,) or (B, T).  return_acf adds the normalised r[:max_lag + 2] per frame."""
    if frame_length <= 0:
        raise ValueError(f"frame_length must be positive, got {frame_length}")
    if hop_length <= 0:
        raise ValueError(f"hop_length must be positive, got {hop_length}")
    if fmin >= fmax:
        raise ValueError(f"fmin ({fmin}) must be less than fmax ({fmax})")
    min_lag, max_lag = int(sr / fmax), int(sr / fmin)
    y = np.asarray(y, dtype=np.float32)
    one_d = y.ndim == 1
    if one_d:
        y = y[None]
    if center:
        y = np.pad(y, [(0, 0), (frame_length // 2, frame_length // 2)])
    B, Lp = y.shape
    T = 1 + (Lp - frame_length) // hop_length
    n_fft = 2 ** int(np.ceil(np.log2(2 * frame_length - 1)))
    f0 = np.zeros((B, T), np.float32)
    voiced = np.zeros((B, T), bool)
    acf = np.zeros((B, T, max_lag + 2), dtype) if return_acf else None
    for b in range(B):
        for t in range(T):
            fr = y[b, t * hop_length:t * hop_length + frame_length].astype(dtype)
            fr = fr - fr.mean()
            Y = np.fft.rfft(fr, n=n_fft)
            r = np.fft.irfft(Y * np.conj(Y), n=n_fft)
            if not r[0] > 1e-10:
                continue
            r = r / r[0]
            if return_acf:
                acf[b, t, :] = r[:max_lag + 2]
            sr_ = r[min_lag:max_lag + 1]
            if len(sr_) == 0:
                continue
            lag = -1
            for i in range(1, len(sr_) - 1):
                if sr_[i] > sr_[i - 1] and sr_[i] > sr_[i + 1] and sr_[i] > threshold:
                    lag = min_lag + i
                    break
            if lag < 0:
                i = int(np.argmax(sr_))
                if sr_[i] > threshold:
                    lag = min_lag + i
            if lag > 0:
                f0[b, t] = sr / lag
                voiced[b, t] = True
    if one_d:
        f0, voiced = f0[0], voiced[0]
        acf = acf[0] if return_acf else None
    return (f0, voiced, acf) if return_acf else (f0, voiced)


def resample_poly(y, up, down, axis=-1, padtype="constant"):
    """resample.py:215-300: gcd reduction, then scipy.signal.resample_poly (the third-party algorithm itself) and a
    float32 cast."""
    import math
    from scipy.signal import resample_poly as sp
    if up <= 0:
        raise ValueError(f"up must be positive, got {up}")
    if down <= 0:
        raise ValueError(f"down must be positive, got {down}")
    g = math.gcd(up, down)
    up, down = up // g, down // g
    y = np.asarray(y)
    if up == 1 and down == 1:
        return y
    return sp(y, up, down, axis=axis, padtype=padtype).astype(np.float32)


def resample_linear(y, orig_sr, target_sr, fix=True, scale=False, axis=-1):
    """resample.py:142-212 (res_type='linear'): positions linspace(0, n - 1, target), blend in float64, float32 cast."""
    y = np.asarray(y, dtype=np.float32)
    if orig_sr == target_sr:
        return y
    y = np.moveaxis(y, axis, -1)
    n = y.shape[-1]
    ratio = target_sr / orig_sr
    m = int(np.round(n * ratio)) if fix else int(np.ceil(n * ratio))
    if m == n:
        return np.moveaxis(y, -1, axis)
    t = np.linspace(0, n - 1, m)
    lo = np.floor(t).astype(np.int32)
    hi = np.minimum(lo + 1, n - 1)
    fr = t - lo
    out = (1 - fr) * y[..., lo] + fr * y[..., hi]
    if scale:
        out = out * ratio
    return np.moveaxis(out.astype(np.float32), -1, axis)


def autocorrelation(y, max_lag=None, normalize=True, center=True, dtype=np.float64):
    """pitch.py:16-116 (the Python path): r = irfft(|rfft(y - mean, n_fft)|^2)[:max_lag], n_fft the power of two >= 2n - 1,
    divided by max(r[0], 1e-10) when normalize; float32 result."""
    y = np.asarray(y, dtype=np.float32)
    one_d = y.ndim == 1
    if one_d:
        y = y[None]
    n = y.shape[1]
    lag = n if max_lag is None else min(max_lag, n)
    x = y.astype(dtype)
    if center:
        x = x - x.mean(-1, keepdims=True)
    n_fft = 2 ** int(np.ceil(np.log2(2 * n - 1)))
    Y = np.fft.rfft(x, n=n_fft, axis=-1)
    r = np.fft.irfft(Y * np.conj(Y), n=n_fft, axis=-1)[:, :lag]
    if normalize:
        r = r / np.maximum(r[:, :1], 1e-10)
    r = r.astype(np.float32)
    return r[0] if one_d else r


def deemphasis(y, coef=0.97, zi=None):
    """framing.py:298-392: scipy.signal.lfilter([1], [1, -coef]) in float32 along the last axis; zi=None runs from a zero
    state and subtracts corr * coef^n, corr = ((2 - coef) y[0] - y[1]) / (3 - coef).  Returns (out, zf)."""
    from scipy import signal
    y = np.asarray(y, dtype=np.float32)
    one_d = y.ndim == 1
    if one_d:
        y = y[None]
    B, L = y.shape
    b = np.array([1.0], dtype=np.float32)
    a = np.array([1.0, -coef], dtype=np.float32)
    if zi is not None:
        z = np.asarray(zi, dtype=np.float32).reshape(-1)
        z = (z if z.size == B else np.broadcast_to(z[:1], (B,))).reshape(B, 1)
        out, zf = signal.lfilter(b, a, y, zi=z, axis=-1)
    else:
        out, zf = signal.lfilter(b, a, y, zi=np.zeros((B, 1), dtype=np.float32), axis=-1)
        corr = ((2 - coef) * y[:, 0:1] - y[:, 1:2]) / (3 - coef)
        out = out - corr * (coef ** np.arange(L, dtype=np.float32))
    out, zf = out.astype(np.float32), zf.astype(np.float32)
    return (out[0], zf[0]) if one_d else (out, zf)


def periodicity(y, sr=22050, fmin=50.0, fmax=2000.0, frame_length=2048, hop_length=512, center=True):
    """pitch.py:267-383: per frame, max over lags [int(sr/fmax), int(sr/fmin)] of the normalised autocorrelation of the
    mean-removed frame (one zero-padded FFT), 0 where r[0] <= 1e-10."""
    y = np.asarray(y, dtype=np.float32)
    one_d = y.ndim == 1
    if one_d:
        y = y[None]
    lo, hi = int(sr / fmax), int(sr / fmin)
    if center:
        y = np.pad(y, [(0, 0), (frame_length // 2, frame_length // 2)])
    T = 1 + (y.shape[1] - frame_length) // hop_length
    n_fft = 2 ** int(np.ceil(np.log2(2 * frame_length - 1)))
    out = np.zeros((y.shape[0], 1, T), dtype=np.float32)
    for b in range(y.shape[0]):
        for t in range(T):
            fr = y[b, t * hop_length: t * hop_length + frame_length]
            fr = fr - np.mean(fr)
            Y = np.fft.rfft(fr, n=n_fft)
            r = np.fft.irfft(Y * np.conj(Y), n=n_fft)
            if r[0] > 1e-10:
                rng = (r / r[0])[lo: hi + 1]
                if len(rng):
                    out[b, 0, t] = np.max(rng)
    return out[0] if one_d else out


def resample_fft(y, orig_sr, target_sr, fix=True, scale=False, dtype=np.float64):
    """resample.py:84-139 -> scipy.signal.resample 1.18.1 (real input, no window), restated: X = rfft(x)[:m2] with
    m = min(num, n), m2 = m // 2 + 1; the unpaired bin m / 2 (m even, num != n) doubled when shrinking, halved when growing;
    irfft(X * num / n, num); times target_sr / orig_sr when scale."""
    y = np.asarray(y, dtype=np.float32)
    n = y.shape[-1]
    ratio = target_sr / orig_sr
    num = int(np.round(n * ratio)) if fix else int(np.ceil(n * ratio))
    if num == n:
        return y
    X = np.fft.rfft(y.astype(dtype), axis=-1)
    m = min(num, n)
    X = X[..., : m // 2 + 1].copy()
    if m % 2 == 0:
        X[..., m // 2] *= 2 if num < n else 0.5
    out = np.fft.irfft(X * (num / n), n=num, axis=-1)
    if scale:
        out = out * ratio
    return out.astype(np.float32)
