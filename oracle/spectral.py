"""CPU oracle for the spectral hot path (TEST INFRASTRUCTURE, not product code).

This module is a NumPy restatement of the algorithm the reference
(zkeown/mlx-audio-primitives, /root/reference) runs for
pad -> frame -> window -> rFFT -> |X|^p -> mel -> dB -> DCT, the inverse
irFFT -> overlap-add chain and Griffin-Lim.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it; the product package never does.

The arithmetic of the reference lives in a third-party dependency that is not
vendored and not installable here: ``mlx>=0.30.0,<1.0.0``
(reference ``pyproject.toml:6,37``).  The restatement therefore follows the
reference's own call sites (cited per function) and the published semantics of
the MLX ops they call (``mx.fft.rfft`` = unnormalised forward DFT with
e^{-j...}; ``mx.fft.irfft(n=)`` = 1/n-normalised Hermitian inverse; ``mx.pad``;
``mx.matmul``; elementwise float32 math).

Pinning (see tests/test_oracle_golden.py, tests/golden/README.md):
  * the reference's own golden vectors (tests/test_cpp_extension.py:525-546),
  * outputs of the reference's *own Python code* executed in this container on
    a NumPy stand-in for ``mlx.core`` (tests/golden/generate_golden.py), and
  * the third-party libraries the reference's tests use as their oracle and
    that exist here (torch.stft/istft, torchaudio, scipy windows / DCT).
``power_to_db(ref=callable)`` and n_fft=400 are pinned by no reference test
("parity unpinned" upstream); for those the shim-executed reference code is
the only authority.

Two precisions are offered everywhere a transform is involved:
``dtype=np.float32`` mimics the reference (float32 pocketfft, float32
elementwise) and ``dtype=np.float64`` is the tight yardstick used for
tolerance checks of the CUDA path (inputs are still float32 values).
"""

from __future__ import annotations

import math
from typing import Callable

import numpy as np

try:  # scipy's pocketfft keeps float32 precision like MLX's CPU FFT does
    import scipy.fft as _fft
except Exception:  # pragma: no cover - numpy>=2 also preserves float32
    _fft = np.fft

# --------------------------------------------------------------------------
# windows  (reference windows.py:19-122, 139-189, 192-256)
# --------------------------------------------------------------------------

_COSINE_SUMS = {  # a0, a1, a2 ... with alternating signs (windows.py:63-67)
    "hann": (0.5, 0.5),
    "hamming": (0.54, 0.46),
    "blackman": (0.42, 0.5, 0.08),
}
_ALIASES = {  # windows.py:112-122
    "hanning": "hann",
    "triangular": "bartlett",
    "boxcar": "rectangular",
    "ones": "rectangular",
}
WINDOW_NAMES = sorted(set(_COSINE_SUMS) | set(_ALIASES) | {"bartlett", "rectangular"})


def _symmetric_window(kind: str, n: int) -> np.ndarray:
    """float64 symmetric window of n points (windows.py:19-58, 92-108)."""
    if kind == "rectangular" or n <= 1:
        return np.ones(n, dtype=np.float64)
    k = np.arange(n, dtype=np.float64)
    if kind == "bartlett":
        return 1 - np.abs(2 * k / (n - 1) - 1)
    coeffs = _COSINE_SUMS[kind]
    w = np.full(n, coeffs[0], dtype=np.float64)
    for i in range(1, len(coeffs)):
        sign = -1.0 if i % 2 else 1.0
        w = w + sign * coeffs[i] * np.cos(2 * i * np.pi * k / (n - 1))
    if kind == "blackman":  # tiny negative end points are clamped (windows.py:55-56)
        w = np.maximum(w, 0.0)
    return w


def get_window(window, n_fft: int, fftbins: bool = True) -> np.ndarray:
    """float32 window of length n_fft (windows.py:139-189, 192-256).

    Periodic windows are the first n_fft points of the (n_fft+1)-point
    symmetric window (windows.py:169-185).  Array windows are length-checked
    and cast (windows.py:234-239).
    """
    if isinstance(window, np.ndarray):
        if window.shape[0] != n_fft:
            raise ValueError(
                f"Window array length ({window.shape[0]}) must match n_fft ({n_fft})"
            )
        return window.astype(np.float32)
    if not isinstance(window, str):
        raise TypeError(f"window must be str or array, got {type(window).__name__}")
    kind = window.lower()
    kind = _ALIASES.get(kind, kind)
    if kind not in _COSINE_SUMS and kind not in ("bartlett", "rectangular"):
        raise ValueError(
            f"Unknown window type: '{window.lower()}'. Supported: {', '.join(WINDOW_NAMES)}"
        )
    n = n_fft + 1 if fftbins else n_fft
    return _symmetric_window(kind, n)[:n_fft].astype(np.float32)


def padded_window(window, win_length: int, n_fft: int) -> np.ndarray:
    """Window centred inside n_fft zeros (stft.py:88-107): left=(n_fft-win)//2."""
    w = get_window(window, win_length, True)
    if win_length < n_fft:
        left = (n_fft - win_length) // 2
        out = np.zeros(n_fft, dtype=np.float32)
        out[left : left + win_length] = w
        return out
    return w


# --------------------------------------------------------------------------
# padding / framing indices (bit-exact integer work)
# --------------------------------------------------------------------------

PAD_MODES = ("constant", "reflect", "edge")


def pad_source_index(i, L: int, pad: int, mode: str):
    """Source sample index for padded position(s) i, or -1 for a zero.

    constant: pad_signal.metal:92 ; edge: pad_signal.metal:53 (clamp) ;
    reflect: pad_signal.metal:11-37 (left src = pad - i ; right src =
    L - 2 - (i - pad - L)), identical to stft.py:445-464.
    """
    i = np.asarray(i, dtype=np.int64)
    j = i - pad
    if mode == "constant":
        return np.where((j >= 0) & (j < L), j, -1)
    if mode == "edge":
        return np.clip(j, 0, L - 1)
    if mode == "reflect":
        right = L - 2 - (j - L)
        return np.where(j < 0, -j, np.where(j >= L, right, j))
    raise ValueError(f"Unknown pad_mode: '{mode}'. Supported: reflect, constant, edge")


def pad_signal(y: np.ndarray, pad: int, mode: str) -> np.ndarray:
    """(B, L) -> (B, L + 2*pad) (stft.py:434-468, pad_signal.cpp:133-164)."""
    if mode not in PAD_MODES:
        raise ValueError(f"Unknown pad_mode: '{mode}'. Supported: reflect, constant, edge")
    B, L = y.shape
    if pad == 0:
        return y.copy()
    if mode == "reflect" and pad > L - 1:
        # native path throws (pad_signal.cpp:102-105); Python slices silently mis-size
        raise ValueError(f"reflect padding ({pad}) requires pad <= signal_length - 1 ({L - 1})")
    src = pad_source_index(np.arange(L + 2 * pad), L, pad, mode)
    out = y[:, np.maximum(src, 0)]
    if mode == "constant":
        out = np.where(src[None, :] >= 0, out, 0).astype(y.dtype)
    return out


def n_frames_for(length: int, frame_length: int, hop: int) -> int:
    """T = 1 + (L - frame_length) // hop (_frame_impl.py:61)."""
    return 1 + (length - frame_length) // hop


def frame_signal(y: np.ndarray, frame_length: int, hop: int) -> np.ndarray:
    """(B, L) -> (B, T, frame_length), frames[b,t,s] = y[b, t*hop+s]
    (_frame_impl.py:48-82, frame_signal.metal:29-35)."""
    if frame_length <= 0:
        raise ValueError(f"frame_length must be positive, got {frame_length}")
    if hop <= 0:
        raise ValueError(f"hop_length must be positive, got {hop}")
    B, L = y.shape
    if L < frame_length:
        raise ValueError(
            f"Signal length ({L}) must be >= frame_length ({frame_length}). "
            "Consider padding the signal."
        )
    T = n_frames_for(L, frame_length, hop)
    idx = (np.arange(T) * hop)[:, None] + np.arange(frame_length)[None, :]
    return y[:, idx]


# --------------------------------------------------------------------------
# STFT / ISTFT
# --------------------------------------------------------------------------


def _resolve(n_fft, hop, win):
    hop = n_fft // 4 if hop is None else hop
    win = n_fft if win is None else win
    return hop, win


def stft(y, n_fft=2048, hop_length=None, win_length=None, window="hann",
         center=True, pad_mode="constant", dtype=np.float32):
    """Complex STFT, logical layout (F, T) / (B, F, T)  (stft.py:136-222).

    pad by n_fft//2 when centred -> frame -> multiply by the padded window ->
    unnormalised forward real DFT along the frame axis (stft.py:118-130) ->
    swap the last two axes (stft.py:216).
    """
    hop, win_length = _resolve(n_fft, hop_length, win_length)
    if hop <= 0:
        raise ValueError(f"hop_length must be positive, got {hop}")
    if win_length <= 0:
        raise ValueError(f"win_length must be positive, got {win_length}")
    if win_length > n_fft:
        raise ValueError(f"win_length ({win_length}) must be <= n_fft ({n_fft})")
    if hop > n_fft:
        raise ValueError(f"hop_length ({hop}) should typically be <= n_fft ({n_fft})")
    y = np.asarray(y, dtype=np.float32)
    one_d = y.ndim == 1
    if one_d:
        y = y[None, :]
    w = padded_window(window, win_length, n_fft)
    if center:
        y = pad_signal(y, n_fft // 2, pad_mode)
    frames = frame_signal(y, n_fft, hop)
    frames = frames.astype(dtype) * w.astype(dtype)
    spec = _fft.rfft(frames, axis=-1)
    spec = spec.astype(np.complex64 if dtype == np.float32 else np.complex128)
    spec = np.swapaxes(spec, 1, 2)
    return spec[0] if one_d else spec


def window_sumsquare(window_nfft: np.ndarray, n_frames: int, hop: int, out_len: int,
                     dtype=np.float32) -> np.ndarray:
    """sum_f w[i - f*hop]^2 over existing frames, ascending f
    (overlap_add.metal:36-50 / stft.py:563-588)."""
    n_fft = window_nfft.shape[0]
    w2 = (window_nfft.astype(dtype) * window_nfft.astype(dtype)).astype(dtype)
    acc = np.zeros(out_len, dtype=dtype)
    for f in range(n_frames):
        lo = f * hop
        if lo >= out_len:
            break
        hi = min(lo + n_fft, out_len)
        acc[lo:hi] += w2[: hi - lo]
    return acc


def overlap_add(frames: np.ndarray, window_nfft: np.ndarray, hop: int, out_len: int,
                dtype=np.float32) -> np.ndarray:
    """Gather overlap-add with fused normalisation.

    y[i] = sum_f w[i-f*hop]*frames[f, i-f*hop] / max(sum_f w[i-f*hop]^2, 1e-8),
    f ascending over the frames that cover i (overlap_add.metal:27-54,
    identical to the inline kernel stft.py:548-596).  Adding whole frames in
    ascending order reproduces the per-sample accumulation order exactly.
    """
    B, T, n_fft = frames.shape
    w = window_nfft.astype(dtype)
    acc = np.zeros((B, out_len), dtype=dtype)
    for f in range(T):
        lo = f * hop
        if lo >= out_len:
            break
        hi = min(lo + n_fft, out_len)
        acc[:, lo:hi] += (w[: hi - lo] * frames[:, f, : hi - lo].astype(dtype)).astype(dtype)
    norm = np.maximum(window_sumsquare(window_nfft, T, hop, out_len, dtype), dtype(1e-8))
    return (acc / norm).astype(dtype)


def istft(stft_matrix, hop_length=None, win_length=None, n_fft=None, window="hann",
          center=True, length=None, dtype=np.float32):
    """Inverse STFT (stft.py:225-344): irfft(n=n_fft) per frame, windowed
    gather-OLA over padded_length (stft.py:300-309), centre trim / length
    handling (stft.py:315-338)."""
    S = np.asarray(stft_matrix)
    if S.ndim not in (2, 3):
        raise ValueError(f"stft_matrix must be 2D or 3D, got {S.ndim}D")
    two_d = S.ndim == 2
    if two_d:
        S = S[None]
    B, F, T = S.shape
    if n_fft is None:
        n_fft = 2 * (F - 1)
    hop, win_length = _resolve(n_fft, hop_length, win_length)
    w = padded_window(window, win_length, n_fft)
    cdt = np.complex64 if dtype == np.float32 else np.complex128
    frames = _fft.irfft(np.swapaxes(S, 1, 2).astype(cdt), n=n_fft, axis=-1).astype(dtype)
    if length is not None:
        padded_length = length + n_fft if center else length
    else:
        padded_length = n_fft + (T - 1) * hop
    y = overlap_add(frames, w, hop, padded_length, dtype)
    if center:
        p = n_fft // 2
        if length is not None:
            y = y[:, p : p + length]
        else:
            end = y.shape[1] - p
            y = y[:, p:end] if end > p else y[:, :0]
    elif length is not None:
        cur = y.shape[1]
        if length < cur:
            y = y[:, :length]
        elif length > cur:
            y = np.concatenate([y, np.zeros((B, length - cur), dtype=y.dtype)], axis=1)
    return y[0] if two_d else y


def magnitude(S):
    """|S| (stft.py:347-362)."""
    return np.abs(S)


def phase(S):
    """atan2(im, re) (stft.py:365-379)."""
    return np.arctan2(S.imag, S.real)


def check_nola(window, hop_length: int, n_fft: int, tol: float = 1e-10) -> bool:
    """NOLA test on hop-sized bins of w^2 (stft.py:382-431)."""
    w = get_window(window, n_fft, True).astype(np.float32)
    sums = np.zeros(hop_length, dtype=np.float32)
    for s in range(n_fft // hop_length):
        sums += w[s * hop_length : (s + 1) * hop_length] ** 2
    rem = n_fft % hop_length
    if rem:
        sums[:rem] += w[-rem:] ** 2
    return bool(sums.min() > tol)


# --------------------------------------------------------------------------
# mel scale, filterbanks
# --------------------------------------------------------------------------

_F_SP = 200.0 / 3
_MIN_LOG_HZ = 1000.0
_MIN_LOG_MEL = _MIN_LOG_HZ / _F_SP
_LOGSTEP = math.log(6.4) / 27.0


def hz_to_mel(f, htk=False):
    """mel.py:31-62 (Slaney piecewise linear/log, or HTK)."""
    f = np.asarray(f, dtype=np.float64)
    if htk:
        return 2595.0 * np.log10(1.0 + f / 700.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(f < _MIN_LOG_HZ, f / _F_SP,
                        _MIN_LOG_MEL + np.log(f / _MIN_LOG_HZ) / _LOGSTEP)


def mel_to_hz(m, htk=False):
    """mel.py:65-93."""
    m = np.asarray(m, dtype=np.float64)
    if htk:
        return 700.0 * (10.0 ** (m / 2595.0) - 1.0)
    return np.where(m < _MIN_LOG_MEL, _F_SP * m,
                    _MIN_LOG_HZ * np.exp(_LOGSTEP * (m - _MIN_LOG_MEL)))


def _triangular_bank(hz_points: np.ndarray, sr: int, n_fft: int, norm) -> np.ndarray:
    """Shared triangle construction (mel.py:137-165, filterbanks.py:126-152,
    248-265): float64 slopes with +1e-10 in the denominators, clip at 0,
    cast to float32 BEFORE the Slaney area scaling (which is then applied in
    place, float32 x float64 -> float32)."""
    n = len(hz_points) - 2
    bins = np.linspace(0, sr / 2.0, 1 + n_fft // 2)[None, :]
    lo, ce, up = (hz_points[:-2, None], hz_points[1:-1, None], hz_points[2:, None])
    rise = (bins - lo) / (ce - lo + 1e-10)
    fall = (up - bins) / (up - ce + 1e-10)
    fb = np.maximum(0, np.minimum(rise, fall)).astype(np.float32)
    if norm == "slaney":
        scale = 2.0 / (hz_points[2 : n + 2] - hz_points[:n])
        fb *= scale[:, None]
    elif norm is not None:
        raise ValueError(f"Unknown norm: '{norm}'. Supported: 'slaney', None")
    return fb


def _check_band_args(n, name, fmin, fmax, sr):
    if n <= 0:
        raise ValueError(f"{name} must be positive, got {n}")
    if fmin < 0:
        raise ValueError(f"fmin must be non-negative, got {fmin}")
    if fmax is None:
        fmax = sr / 2.0
    if fmin >= fmax:
        raise ValueError(f"fmin ({fmin}) must be less than fmax ({fmax})")
    if fmax > sr / 2.0:
        raise ValueError(f"fmax ({fmax}) cannot exceed Nyquist frequency ({sr / 2.0})")
    return fmax


def mel_filterbank(sr, n_fft, n_mels=128, fmin=0.0, fmax=None, htk=False, norm="slaney"):
    """(n_mels, F) float32 (mel.py:101-168, 171-242)."""
    fmax = _check_band_args(n_mels, "n_mels", fmin, fmax, sr)
    pts = mel_to_hz(np.linspace(hz_to_mel(fmin, htk), hz_to_mel(fmax, htk), n_mels + 2), htk)
    return _triangular_bank(pts, sr, n_fft, norm)


def linear_filterbank(sr, n_fft, n_bands=64, fmin=0.0, fmax=None, norm="slaney"):
    """filterbanks.py:234-342."""
    fmax = _check_band_args(n_bands, "n_bands", fmin, fmax, sr)
    return _triangular_bank(np.linspace(fmin, fmax, n_bands + 2), sr, n_fft, norm)


def melspectrogram(y, sr=22050, n_fft=2048, hop_length=None, win_length=None,
                   window="hann", center=True, pad_mode="constant", power=2.0,
                   n_mels=128, fmin=0.0, fmax=None, htk=False, norm="slaney",
                   dtype=np.float32):
    """mel.py:245-352: stft -> abs -> pow(power) unless power == 1 -> dense
    (n_mels, F) @ (F, T) product."""
    S = np.abs(stft(y, n_fft, hop_length, win_length, window, center, pad_mode, dtype))
    S = S.astype(dtype)
    if power != 1.0:
        S = np.power(S, dtype(power))
    fb = mel_filterbank(sr, n_fft, n_mels, fmin, fmax, htk, norm).astype(dtype)
    return np.matmul(fb, S).astype(dtype)


# --------------------------------------------------------------------------
# dB conversions (convert.py:14-60, 63-198)
# --------------------------------------------------------------------------


def _to_db(S, ref, coef, amin, top_db, dtype=np.float32):
    S = np.asarray(S).astype(dtype)
    # a callable ref sees the *unclamped* input (convert.py:42-43)
    ref_value = dtype(ref(S)) if callable(ref) else dtype(ref)
    S = np.maximum(S, dtype(amin))
    ref_value = np.maximum(ref_value, dtype(amin))
    out = (dtype(coef) * np.log10(S / ref_value)).astype(dtype)  # divide, then log
    if top_db is not None:
        if top_db <= 0:
            raise ValueError(f"top_db must be positive, got {top_db}")
        out = np.maximum(out, out.max() - dtype(top_db))  # max over the WHOLE array
    return out


def power_to_db(S, ref=1.0, amin=1e-10, top_db=80.0, dtype=np.float32):
    return _to_db(S, ref, 10.0, amin, top_db, dtype)


def amplitude_to_db(S, ref=1.0, amin=1e-5, top_db=80.0, dtype=np.float32):
    return _to_db(S, ref, 20.0, amin, top_db, dtype)


def db_to_power(S_db, ref=1.0, dtype=np.float32):
    """ref * 10^(S_db/10) (convert.py:100-129)."""
    return (dtype(ref) * np.power(dtype(10.0), np.asarray(S_db, dtype=dtype) / dtype(10.0))).astype(dtype)


def db_to_amplitude(S_db, ref=1.0, dtype=np.float32):
    """ref * 10^(S_db/20) (convert.py:169-198)."""
    return (dtype(ref) * np.power(dtype(10.0), np.asarray(S_db, dtype=dtype) / dtype(20.0))).astype(dtype)


# --------------------------------------------------------------------------
# DCT-II / MFCC (mfcc.py:24-66, 69-140, 143-287)
# --------------------------------------------------------------------------


def dct_matrix(n_out: int, n_in: int, norm="ortho") -> np.ndarray:
    """D[k, n] = cos(pi k (2n+1) / (2N)); ortho: row 0 / sqrt(N), others
    * sqrt(2/N); float64 then float32 (mfcc.py:54-66)."""
    n = np.arange(n_in)
    k = np.arange(n_out)[:, None]
    D = np.cos(np.pi * k * (2 * n + 1) / (2 * n_in))
    if norm == "ortho":
        D[0] *= 1.0 / np.sqrt(n_in)
        D[1:] *= np.sqrt(2.0 / n_in)
    return D.astype(np.float32)


def dct(x, type=2, n=None, axis=-1, norm="ortho", dtype=np.float32):
    if type != 2:
        raise ValueError(f"Only DCT type 2 is supported, got {type}")
    x = np.asarray(x).astype(dtype)
    size = x.shape[axis]
    n = size if n is None else n
    D = dct_matrix(n, size, norm).astype(dtype)
    moved = np.moveaxis(x, axis, -1)
    return np.moveaxis(np.matmul(moved, D.T), -1, axis).astype(dtype)


def mfcc(y=None, sr=22050, S=None, n_mfcc=20, dct_type=2, norm="ortho", lifter=0,
         n_fft=2048, hop_length=512, win_length=None, window="hann", center=True,
         pad_mode="constant", power=2.0, n_mels=128, fmin=0.0, fmax=None, htk=False,
         mel_norm="slaney", dtype=np.float32):
    """mfcc.py:226-287: mel -> power_to_db(ref=1, amin=1e-10, top_db=80) ->
    DCT-II over the mel axis -> optional sinusoidal lifter.  A caller-supplied
    S is taken as log-power already."""
    if n_mfcc <= 0:
        raise ValueError(f"n_mfcc must be positive, got {n_mfcc}")
    if S is None:
        M = melspectrogram(y, sr, n_fft, hop_length, win_length, window, center, pad_mode,
                           power, n_mels, fmin, fmax, htk, mel_norm, dtype)
        M = power_to_db(M, 1.0, 1e-10, 80.0, dtype)
    else:
        M = np.asarray(S).astype(dtype)
    out = dct(M, dct_type, n_mfcc, axis=-2, norm=norm, dtype=dtype)
    if lifter > 0:
        k = np.arange(n_mfcc)
        lift = (1 + (lifter / 2.0) * np.sin(np.pi * (k + 1) / lifter)).astype(np.float32)
        out = (out * lift[:, None].astype(dtype)).astype(dtype)
    return out


# --------------------------------------------------------------------------
# Griffin-Lim (griffinlim.py:92-196)
# --------------------------------------------------------------------------


def griffinlim_init_angles(shape, init="random", random_state=None) -> np.ndarray:
    """Host RNG draw in (B, F, T) C order, float64 -> float32 (griffinlim.py:112-119)."""
    if init == "random":
        rng = np.random.default_rng(random_state)
        return rng.uniform(-np.pi, np.pi, shape).astype(np.float32)
    if init == "zeros":
        return np.zeros(shape, dtype=np.float32)
    raise ValueError(f"Unknown init: '{init}'. Supported: 'random', 'zeros'")


def griffinlim(S, n_iter=32, hop_length=None, win_length=None, n_fft=None, window="hann",
               center=True, length=None, pad_mode="constant", momentum=0.99,
               init="random", random_state=None, dtype=np.float32, angles=None):
    """The reference's fast-Griffin-Lim variant: rebuilt = new + m*(new - tprev),
    tprev = new, with new = S*exp(j*angle(stft(istft(rebuilt)))) (griffinlim.py:129-183)."""
    if n_iter <= 0:
        raise ValueError(f"n_iter must be positive, got {n_iter}")
    if momentum < 0.0:
        raise ValueError(f"momentum must be >= 0.0, got {momentum}")
    if momentum >= 1.0:
        raise ValueError(f"momentum must be < 1.0, got {momentum}")
    S = np.asarray(S).astype(dtype)
    batched = S.ndim == 3
    if not batched:
        S = S[None]
    B, F, T = S.shape
    if n_fft is None:
        n_fft = 2 * (F - 1)
    hop, win_length = _resolve(n_fft, hop_length, win_length)
    cdt = np.complex64 if dtype == np.float32 else np.complex128
    if angles is None:
        angles = griffinlim_init_angles((B, F, T), init, random_state)
    rebuilt = (S * np.exp(1j * angles.astype(cdt))).astype(cdt)
    tprev = rebuilt
    for _ in range(n_iter):
        y = istft(rebuilt, hop, win_length, n_fft, window, center, length, dtype)
        new = stft(y, n_fft, hop, win_length, window, center, pad_mode, dtype)
        t_new = new.shape[-1]
        if t_new > T:
            new = new[..., :T]
        elif t_new < T:
            new = np.concatenate([new, np.zeros((B, F, T - t_new), dtype=new.dtype)], axis=-1)
        ang = np.arctan2(new.imag, new.real).astype(dtype)
        new = (S * np.exp(1j * ang.astype(cdt))).astype(cdt)
        if momentum > 0:
            rebuilt = (new + dtype(momentum) * (new - tprev)).astype(cdt)
            tprev = new
        else:
            rebuilt = new
    y = istft(rebuilt, hop, win_length, n_fft, window, center, length, dtype)
    return y if batched else y[0]
