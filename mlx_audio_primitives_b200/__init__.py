"""mlx_audio_primitives_b200 -- the spectral hot path of zkeown/mlx-audio-primitives
(stft, istft, magnitude, melspectrogram, mel_filterbank, get_window, power_to_db, mfcc, griffinlim
and friends) rebuilt from scratch for NVIDIA B200 (sm_100a).

Same librosa-compatible signatures as the reference; arrays are ``torch`` CUDA tensors (or any
``__dlpack__`` / NumPy input, which is moved to the current CUDA device).  All arithmetic runs in
hand-written CUDA kernels behind the C ABI of ``include/mlxa_cuda.h``; there is no CPU fallback.
"""
from ._extension import HAS_CPP_EXT as _HAS_CPP_EXT  # noqa: F401  (loads the CUDA library or raises)
from .convert import amplitude_to_db, db_to_amplitude, db_to_power, power_to_db
from .filterbanks import bark_filterbank, bark_to_hz, hz_to_bark, linear_filterbank
from .features import (spectral_bandwidth, spectral_centroid, spectral_contrast, spectral_flatness, spectral_rolloff,
                       zero_crossing_rate)
from .framing import deemphasis, frame, preemphasis, rms
from .griffinlim import griffinlim, griffinlim_iter
from .mel import hz_to_mel, mel_filterbank, mel_to_hz, melspectrogram
from .mfcc import dct, dct_matrix, delta, mfcc
from .stft import check_nola, istft, magnitude, overlap_add, pad_signal, phase, stft
from .windows import get_window
from .pipeline import LogMelPlan
from .pitch import autocorrelation, periodicity, pitch_detect_acf
from .resample import resample, resample_poly
from . import distributed

__version__ = "0.1.0"


def _scope_entry_points():
    """Every array-taking entry point runs on the device of its input (``_tensor.on_input_device``); the wrapped
    function replaces the plain one in its home module too, so ``from mlx_audio_primitives_b200.stft import stft``
    gets the same behaviour."""
    import sys
    from ._tensor import on_input_device
    names = ["stft", "istft", "magnitude", "phase", "melspectrogram", "power_to_db", "amplitude_to_db", "db_to_power",
             "db_to_amplitude", "dct", "mfcc", "griffinlim", "griffinlim_iter", "frame", "pad_signal", "overlap_add",
             "spectral_centroid", "spectral_bandwidth", "spectral_rolloff", "spectral_flatness", "spectral_contrast",
             "zero_crossing_rate", "rms", "preemphasis", "delta", "pitch_detect_acf", "autocorrelation", "periodicity",
             "deemphasis", "resample", "resample_poly"]
    pkg = sys.modules[__name__]
    for name in names:
        fn = getattr(pkg, name)
        wrapped = on_input_device(fn)
        setattr(pkg, name, wrapped)
        home = sys.modules.get(fn.__module__)
        if home is not None and getattr(home, name, None) is fn:
            setattr(home, name, wrapped)


_scope_entry_points()
del _scope_entry_points

__all__ = [
    "stft", "istft", "magnitude", "phase", "check_nola", "get_window",
    "mel_filterbank", "melspectrogram", "hz_to_mel", "mel_to_hz",
    "power_to_db", "amplitude_to_db", "db_to_power", "db_to_amplitude",
    "dct", "dct_matrix", "mfcc", "griffinlim", "griffinlim_iter", "frame",
    "linear_filterbank", "bark_filterbank", "hz_to_bark", "bark_to_hz",
    "pad_signal", "overlap_add", "distributed", "LogMelPlan",
    "spectral_centroid", "spectral_bandwidth", "spectral_rolloff", "spectral_flatness", "spectral_contrast",
    "zero_crossing_rate", "rms", "preemphasis", "delta", "pitch_detect_acf", "autocorrelation", "periodicity", "deemphasis", "resample", "resample_poly",
]
