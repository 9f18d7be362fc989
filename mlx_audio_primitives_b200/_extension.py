"""Loader of the sm_100a C-ABI library (replaces reference ``_extension.py:30-43``).

The reference silently falls back to Python/MLX when its nanobind module is
missing.  This build has NO fallback: if ``_lib/libmlxaudio_cuda.so`` is absent
or does not export the ABI declared in ``include/mlxa_cuda.h`` the import fails
loudly with the build command.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MLXA_CUDA_LIB", os.path.join(_HERE, "_lib", "libmlxaudio_cuda.so"))
ABI_VERSION = 6

_i64, _i32, _f32, _f64, _p = C.c_int64, C.c_int, C.c_float, C.c_double, C.c_void_p

# name -> argtypes; mirrors include/mlxa_cuda.h one to one
SIGNATURES = {
    "mlxa_abi_version": [],
    "mlxa_has_fast_plan": [_i32],
    "mlxa_has_fused_feature": [_i32],
    "mlxa_pad_signal_f32": [_p, _i64, _i64, _i64, _i32, _p, _p],
    "mlxa_frame_signal_f32": [_p, _i64, _i64, _i32, _i32, _p, _p],
    "mlxa_overlap_add_f32": [_p, _p, _i64, _i64, _i32, _i32, _i64, _p, _p],
    "mlxa_window_sumsquare_f32": [_p, _i32, _i32, _i64, _i64, _p, _p],
    "mlxa_stft_f32": [_p, _i64, _i64, _i64, _p, _i32, _i32, _i32, _i32, _p, _p],
    "mlxa_plan_group": [_i32],
    "mlxa_pack_filterbank": [_p, _i32, _i32, _i32, _p, _i64, _p],
    "mlxa_melspec_f32": [_p, _i64, _i64, _i64, _p, _i32, _i32, _i32, _i32, _f32, _p, _i32, _i64,
                         _p, _p, _i32, _f32, _f32, _f32, _p, _p, _p],
    "mlxa_istft_f32": [_p, _i64, _i64, _i32, _p, _p, _i32, _i32, _i64, _i64, _i64, _p, _i64, _p],
    "mlxa_istft_momentum_f32": [_p, _p, _f32, _p, _i64, _i64, _i32, _p, _p, _i32, _i32, _i64, _i64, _i64, _p, _i64, _p],
    "mlxa_griffinlim_project_f32": [_p, _i64, _i64, _i64, _p, _i32, _i32, _i32, _i32, _i64, _i64, _p, _p, _p],
    "mlxa_polar_f32": [_p, _p, _i64, _p, _p],
    "mlxa_momentum_f32": [_p, _p, _f32, _i64, _p, _p],
    "mlxa_pcg64_uniform_f32": [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_double, C.c_double, _i64, _p, _p],
    "mlxa_pcg64_polar_f32": [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_double, C.c_double, _p, _i64, _i64, _i64, _p, _p],
    "mlxa_magnitude_f32": [_p, _i64, _p, _p],
    "mlxa_phase_f32": [_p, _i64, _p, _p],
    "mlxa_transpose_f32": [_p, _i64, _i64, _i64, _p, _p],
    "mlxa_transpose_c64": [_p, _i64, _i64, _i64, _p, _p],
    "mlxa_max_f32": [_p, _i64, _p, _p],
    "mlxa_fill_f32": [_p, _i64, _f32, _p],
    "mlxa_to_db_f32": [_p, _i64, _f32, _f32, _f32, _p, _i32, _f32, _p, _p, _p, _p, _p],
    "mlxa_db_floor_f32": [_p, _i64, _f32, _f32, _f32, _f32, _p, _p, _p],
    "mlxa_db_floor_blocks_f32": [_p, _i64, _i32, _i64, _f32, _f32, _f32, _f32, _p, _p, _p, _p, _p, _p],
    "mlxa_spectral_stats_f32": [_p, _i32, _i64, _i64, _i32, _p, _i32, _f32, _f32, _i32, _p, _p, _p],
    "mlxa_pitch_acf_f32": [_p, _i64, _i64, _i64, _i32, _i32, _i32, _f32, _f32, _f32, _f32, _p, _p, _p],
    "mlxa_resample_poly_f32": [_p, _i64, _i64, _p, _i32, _i32, _i32, _i64, _i64, _p, _p],
    "mlxa_resample_linear_f32": [_p, _i64, _i64, _i64, C.c_double, _i32, _p, _p],
    "mlxa_autocorrelation_fft_work_bytes": [_i64, _i64],
    "mlxa_autocorrelation_fft_f32": [_p, _i64, _i64, _i64, _i32, _i32, _i32, _p, _p, _p, _i64, _p],
    "mlxa_resample_fft_work_bytes": [_i64, _i64, _i64],
    "mlxa_resample_fft_f32": [_p, _i64, _i64, _i64, _i64, _f32, _p, _i64, _p, _i64, _p],
    "mlxa_periodicity_f32": [_p, _i64, _i64, _i64, _i32, _i32, _i32, _f32, _f32, _f32, _p, _p],
    "mlxa_deemphasis_f32": [_p, _i64, _i64, _i64, _f64, _p, _i32, _p, _i64, _p, _p],
    "mlxa_autocorrelation_f32": [_p, _i64, _i64, _i64, _i32, _i32, _i32, _p, _p, _p],
    "mlxa_savgol_f32": [_p, _i64, _i64, _p, _i32, _i32, _f32, _p, _p, _p, _p],
    "mlxa_spectral_contrast_f32": [_p, _i32, _i64, _i64, _i32, _p, _i32, _i32, _p, _p],
    "mlxa_spectral_feature_f32": [_p, _i64, _i64, _i64, _p, _i32, _i32, _i32, _i32, _p, _f32, _i32, _f32, _f32, _i32, _p, _p, _p],
    "mlxa_frame_stats_f32": [_p, _i64, _i64, _i64, _i32, _i32, _i32, _i32, _i32, _p, _p],
    "mlxa_preemphasis_f32": [_p, _i64, _i64, _i64, _f32, _p, _p, _p, _p],
    "mlxa_from_db_f32": [_p, _i64, _f32, _f32, _p, _p],
    "mlxa_dct_f32": [_p, _i64, _i32, _p, _i32, _p, _p],
    "mlxa_mfcc_tail_f32": [_p, _i64, _i32, _i64, _p, _i32, _p, _i32, _f32, _f32, _i32, _f32, _p, _p, _p],
    "mlxa_logmel_host_f32": [_p, _i64, _i64, _p, _i32, _i32, _i32, _i32, _f32, _p, _i32, _i64,
                             _i32, _i32, _f32, _f32, _i32, _f32, _p, _p],
    "mlxa_ffma_probe": [_p, _i32, _i32, _i32, _p],
}


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"CUDA extension not found at {LIB_PATH}. Build it with "
            "`python mlx_audio_primitives_b200/build.py` (needs nvcc, targets sm_100a). "
            "There is no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:  # pragma: no cover - stale build
            raise ImportError(f"{LIB_PATH} does not export {name}; rebuild the extension") from e
        fn.argtypes = args
        fn.restype = C.c_int
    lib.mlxa_last_error.argtypes = []
    lib.mlxa_last_error.restype = C.c_char_p
    lib.mlxa_packed_bank_words.argtypes = [_i32, _i64, _i32]
    lib.mlxa_packed_bank_words.restype = _i64
    lib.mlxa_resample_fft_work_bytes.restype = _i64
    lib.mlxa_autocorrelation_fft_work_bytes.restype = _i64
    if lib.mlxa_abi_version() != ABI_VERSION:
        raise ImportError(f"ABI mismatch: library {lib.mlxa_abi_version()} != host layer {ABI_VERSION}; rebuild")
    return lib


_ext = _load()
HAS_CPP_EXT: bool = True  # kept for API compatibility with the reference


def check(rc: int, what: str) -> None:
    """Translate a C-ABI status into the exception the reference would raise."""
    if rc == 0:
        return
    msg = _ext.mlxa_last_error().decode("utf-8", "replace")
    if rc < 0:
        raise ValueError(f"{what}: {msg}")
    raise RuntimeError(f"{what}: CUDA error {rc}: {msg}")


__all__ = ["_ext", "HAS_CPP_EXT", "check", "LIB_PATH", "SIGNATURES"]
