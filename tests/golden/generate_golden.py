"""Generate tests/golden/*.npz by running the REFERENCE'S OWN Python code.

Run in the build container only (needs /root/reference):

    python tests/golden/generate_golden.py

MLX is not installable here, so the reference package is imported on top of
``tests/golden/mlx_shim`` -- a float32 NumPy stand-in for the ``mlx.core``
calls the hot path makes (FFT = scipy pocketfft in float32, like the MLX CPU
backend's pocketfft).  Everything above that substrate -- padding, framing,
window caches, trims, dB order of operations, MFCC, the Griffin-Lim update --
is the reference's unmodified code (its no-extension path, `_extension.py:40`).
The only restated piece is the inline Metal overlap-add kernel
(stft.py:548-596), re-expressed in the shim because Metal cannot run here.

The fixtures are small on purpose; they pin the oracle (tests/test_oracle_golden.py)
and are compared directly with the CUDA path (tests/test_gpu_parity.py).
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "mlx_shim"))
sys.path.insert(0, "/root/reference")

import mlx.core as mx  # noqa: E402  (the shim)
import mlx_audio_primitives as ref  # noqa: E402  (the reference)

assert not ref._HAS_CPP_EXT


def A(x):
    return np.asarray(x)


def signal(n, seed=42):
    return np.random.default_rng(seed).standard_normal(n).astype(np.float32)


def chirp(n, sr=22050, seed=42):
    # reference benchmarks/utils.py:92-115 : chirp + 0.1 * noise
    t = np.arange(n) / sr
    x = np.sin(2 * np.pi * (100 + 1000 * t) * t)
    return (x + 0.1 * np.random.default_rng(seed).standard_normal(n)).astype(np.float32)


out = {}

# ---- constants -----------------------------------------------------------
for name in ["hann", "hamming", "blackman", "bartlett", "rectangular"]:
    for n in [16, 400, 512, 2048]:
        for fb in [True, False]:
            out[f"window/{name}/{n}/{int(fb)}"] = A(ref.get_window(name, n, fftbins=fb))
for (sr, n_fft, n_mels, fmin, fmax, htk, norm) in [
    (22050, 2048, 128, 0.0, None, False, "slaney"),
    (16000, 400, 80, 0.0, None, False, "slaney"),
    (44100, 4096, 128, 0.0, None, False, "slaney"),
    (22050, 1024, 40, 50.0, 8000.0, True, None),
    (16000, 512, 64, 20.0, 7600.0, True, "slaney"),
]:
    out[f"melfb/{sr}/{n_fft}/{n_mels}/{fmin}/{fmax}/{int(htk)}/{norm}"] = A(
        ref.mel_filterbank(sr, n_fft, n_mels, fmin, fmax, htk, norm))
out["linfb/22050/1024/32"] = A(ref.linear_filterbank(22050, 1024, 32))
for (sr, n_fft, n_bands, fmin, fmax, formula, norm) in [
    (22050, 1024, 24, 0.0, None, "zwicker", "slaney"),
    (16000, 512, 20, 50.0, 7000.0, "traunmuller", None),
    (44100, 2048, 32, 0.0, None, "zwicker", None),
]:
    out[f"barkfb/{sr}/{n_fft}/{n_bands}/{fmin}/{fmax}/{formula}/{norm}"] = A(
        ref.bark_filterbank(sr, n_fft, n_bands, fmin, fmax, formula, norm))
hz = np.array([0.0, 55.0, 440.0, 999.0, 1000.0, 4000.0, 11025.0])
out["hz"] = hz
out["hz_to_mel/slaney"] = ref.hz_to_mel(hz)
out["hz_to_mel/htk"] = ref.hz_to_mel(hz, htk=True)
out["mel_to_hz/slaney"] = ref.mel_to_hz(ref.hz_to_mel(hz))
out["mel_to_hz/htk"] = ref.mel_to_hz(ref.hz_to_mel(hz, htk=True), htk=True)
from mlx_audio_primitives.mfcc import _compute_dct_matrix_np  # noqa: E402
for (k, n, nm) in [(40, 128, "ortho"), (13, 40, "ortho"), (20, 80, None)]:
    b, shp = _compute_dct_matrix_np(k, n, nm)
    out[f"dctmat/{k}/{n}/{nm}"] = np.frombuffer(b, dtype=np.float32).reshape(shp)

# ---- pad / frame (indices: ramps make values == source index) --------------
from mlx_audio_primitives.stft import _pad_signal  # noqa: E402
ramp = np.arange(1, 41, dtype=np.float32).reshape(2, 20)  # 1-based so zero padding is visible
for mode in ["constant", "reflect", "edge"]:
    for pad in [0, 3, 8, 19]:
        out[f"pad/{mode}/{pad}"] = A(_pad_signal(mx.array(ramp), pad, mode))
out["pad/input"] = ramp
fr_in = np.arange(100, dtype=np.float32).reshape(2, 50)
out["frame/input"] = fr_in
for (fl, hop) in [(8, 2), (16, 16), (10, 3), (50, 1)]:
    out[f"frame/{fl}/{hop}"] = A(ref.frame(mx.array(fr_in), fl, hop))

# ---- STFT / ISTFT ---------------------------------------------------------
y2 = np.stack([signal(6000, 1), chirp(6000)])
out["stft/input"] = y2
stft_cases = [
    dict(n_fft=2048, hop_length=512),
    dict(n_fft=1024, hop_length=256, pad_mode="reflect"),
    dict(n_fft=512, hop_length=128, pad_mode="edge"),
    dict(n_fft=512, hop_length=128, center=False),
    dict(n_fft=400, hop_length=160),
    dict(n_fft=400, hop_length=160, window="hamming", pad_mode="reflect"),
    dict(n_fft=1024, hop_length=300, win_length=800, window="blackman"),
    dict(n_fft=256, hop_length=256, window="bartlett"),
    dict(n_fft=64, hop_length=1, center=False),
    dict(n_fft=4096, hop_length=1024),
    dict(n_fft=600, hop_length=150),
    dict(n_fft=1000, hop_length=250),
]
for i, kw in enumerate(stft_cases):
    yin = y2[:, :300] if kw.get("hop_length") == 1 else y2
    S = ref.stft(mx.array(yin), **kw)
    out[f"stft/{i}"] = A(S)
    ikw = {k: v for k, v in kw.items() if k != "pad_mode"}
    out[f"istft/{i}"] = A(ref.istft(S, **ikw))
    if kw.get("center", True):
        out[f"istft_len/{i}"] = A(ref.istft(S, length=yin.shape[1], **ikw))
out["stft/ncases"] = np.array(len(stft_cases))
# 1-D input, length shorter / longer than natural, center=False with length
S1 = ref.stft(mx.array(y2[0]), n_fft=512, hop_length=128)
out["stft1d"] = A(S1)
for L in [5000, 6000, 7000]:
    out[f"istft1d_len/{L}"] = A(ref.istft(S1, hop_length=128, length=L))
Snc = ref.stft(mx.array(y2), n_fft=512, hop_length=128, center=False)
for L in [4000, 6500]:
    out[f"istft_nc_len/{L}"] = A(ref.istft(Snc, hop_length=128, center=False, length=L))
out["magnitude"] = A(ref.magnitude(S1))
out["phase"] = A(ref.phase(S1))

# ---- mel / dB / MFCC ------------------------------------------------------
mel_cases = [
    dict(sr=22050, n_fft=2048, hop_length=512, n_mels=128),
    dict(sr=16000, n_fft=400, hop_length=160, n_mels=80),
    dict(sr=22050, n_fft=1024, hop_length=256, n_mels=40, power=1.0),
    dict(sr=44100, n_fft=4096, hop_length=1024, n_mels=128),
    dict(sr=16000, n_fft=512, hop_length=128, n_mels=64, htk=True, fmin=20.0, fmax=7600.0, power=1.5),
]
for i, kw in enumerate(mel_cases):
    M = ref.melspectrogram(mx.array(y2), **kw)
    out[f"mel/{i}"] = A(M)
    out[f"db_default/{i}"] = A(ref.power_to_db(M))
    out[f"db_refmax/{i}"] = A(ref.power_to_db(M, ref=mx.max))
    out[f"db_refmax_notop/{i}"] = A(ref.power_to_db(M, ref=mx.max, top_db=None))
    out[f"db_ref05_amin/{i}"] = A(ref.power_to_db(M, ref=0.5, amin=1e-5, top_db=60.0))
out["mel/ncases"] = np.array(len(mel_cases))
amp = np.abs(A(S1))
out["ampdb"] = A(ref.amplitude_to_db(mx.array(amp)))
out["ampdb_refmax"] = A(ref.amplitude_to_db(mx.array(amp), ref=mx.max, top_db=None))
dbv = np.linspace(-80, 10, 37).astype(np.float32)
out["dbv"] = dbv
out["db_to_power"] = A(ref.db_to_power(mx.array(dbv), ref=2.0))
out["db_to_amplitude"] = A(ref.db_to_amplitude(mx.array(dbv)))
mfcc_cases = [
    dict(sr=22050, n_mfcc=20),
    dict(sr=22050, n_mfcc=40, n_fft=2048, hop_length=512, n_mels=128),
    dict(sr=44100, n_mfcc=40, n_fft=4096, hop_length=1024),
    dict(sr=16000, n_mfcc=13, n_fft=400, hop_length=160, n_mels=40, lifter=22),
    dict(sr=22050, n_mfcc=13, n_fft=1024, hop_length=256, n_mels=80, norm=None),
]
for i, kw in enumerate(mfcc_cases):
    out[f"mfcc/{i}"] = A(ref.mfcc(mx.array(y2), **kw))
out["mfcc/ncases"] = np.array(len(mfcc_cases))
xd = signal(3 * 7 * 64, 5).reshape(3, 7, 64)
out["dct/input"] = xd
out["dct/ortho"] = A(ref.dct(mx.array(xd)))
out["dct/n20_none"] = A(ref.dct(mx.array(xd), n=20, norm=None))
out["dct/axis1"] = A(ref.dct(mx.array(xd), axis=1, n=5))

# ---- Griffin-Lim (values: only this shim-executed reference pins them) -----
yg = np.stack([chirp(4096, seed=3), signal(4096, 7) * 0.3])
Sg = ref.magnitude(ref.stft(mx.array(yg), n_fft=512, hop_length=128))
out["gl/S"] = A(Sg)
out["gl/random8"] = A(ref.griffinlim(Sg, n_iter=8, hop_length=128, random_state=0))
out["gl/zeros4_m0"] = A(ref.griffinlim(Sg, n_iter=4, hop_length=128, init="zeros", momentum=0.0))
out["gl/len"] = A(ref.griffinlim(Sg, n_iter=3, hop_length=128, random_state=1, length=4000))
out["gl/1d"] = A(ref.griffinlim(Sg[0], n_iter=2, hop_length=128, random_state=2))

np.savez_compressed(os.path.join(HERE, "reference_outputs.npz"), **out)
import json  # noqa: E402
with open(os.path.join(HERE, "cases.json"), "w") as f:
    json.dump({"stft": stft_cases, "mel": mel_cases, "mfcc": mfcc_cases}, f, indent=1)
print("wrote", len(out), "arrays,",
      os.path.getsize(os.path.join(HERE, "reference_outputs.npz")) / 1e6, "MB")
