"""Golden vectors for the section-8(f) rows (spectral features, rms, zcr, preemphasis): the reference's
own Python code executed on the MLX stand-in (see generate_golden.py).  Build container only."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "mlx_shim"))
sys.path.insert(0, "/root/reference")
import mlx.core as mx  # noqa: E402
import mlx_audio_primitives as ref  # noqa: E402

A = np.asarray
g = np.load(os.path.join(HERE, "reference_outputs.npz"))
y2 = g["stft/input"]  # (2, 6000)
out = {}
cases = [dict(n_fft=2048, hop_length=512), dict(n_fft=1024, hop_length=256), dict(n_fft=400, hop_length=160, sr=16000),
         dict(n_fft=600, hop_length=150)]
for i, kw in enumerate(cases):
    sr = kw.get("sr", 22050)
    k2 = {k: v for k, v in kw.items() if k != "sr"}
    out[f"centroid/{i}"] = A(ref.spectral_centroid(mx.array(y2), sr=sr, **k2))
    out[f"bandwidth/{i}"] = A(ref.spectral_bandwidth(mx.array(y2), sr=sr, **k2))
    out[f"bandwidth_p3/{i}"] = A(ref.spectral_bandwidth(mx.array(y2), sr=sr, p=3.0, norm=False, **k2))
    out[f"rolloff/{i}"] = A(ref.spectral_rolloff(mx.array(y2), sr=sr, **k2))
    out[f"rolloff50/{i}"] = A(ref.spectral_rolloff(mx.array(y2), sr=sr, roll_percent=0.5, **k2))
    out[f"flatness/{i}"] = A(ref.spectral_flatness(mx.array(y2), **k2))
    out[f"flatness_p1/{i}"] = A(ref.spectral_flatness(mx.array(y2), power=1.0, amin=1e-6, **k2))
    out[f"contrast/{i}"] = A(ref.spectral_contrast(mx.array(y2), sr=sr, **k2))
    out[f"contrast_lin/{i}"] = A(ref.spectral_contrast(mx.array(y2), sr=sr, n_bands=4, fmin=150.0, quantile=0.1, linear=True, **k2))
out["ncases"] = np.array(len(cases))
S = ref.magnitude(ref.stft(mx.array(y2[0]), n_fft=512, hop_length=128))
out["S1d"] = A(S)
out["centroid_S1d"] = A(ref.spectral_centroid(S=S, sr=22050, n_fft=512))
out["rolloff_S1d"] = A(ref.spectral_rolloff(S=S, sr=22050, n_fft=512))
out["contrast_S1d"] = A(ref.spectral_contrast(S=S, sr=22050, n_fft=512))
for (fl, hop, center, mode) in [(2048, 512, True, "constant"), (400, 160, True, "edge"), (256, 64, False, "constant")]:
    out[f"rms/{fl}/{hop}/{int(center)}/{mode}"] = A(ref.rms(mx.array(y2), fl, hop, center=center, pad_mode=mode))
    zm = "edge" if mode == "edge" else "constant"
    out[f"zcr/{fl}/{hop}/{int(center)}/{zm}"] = A(ref.zero_crossing_rate(mx.array(y2), fl, hop, center=center, pad_mode=zm))
out["rms1d"] = A(ref.rms(mx.array(y2[1]), 1024, 256))
out["pre/default"] = A(ref.preemphasis(mx.array(y2)))
o2, zf = ref.preemphasis(mx.array(y2), coef=0.9, zi=mx.array(np.array([0.5, -0.25], np.float32)), return_zf=True)
out["pre/zi"], out["pre/zf"] = A(o2), A(zf)
out["pre/1d"] = A(ref.preemphasis(mx.array(y2[0]), coef=0.5))
tt = np.arange(9000) / 22050.0
yp = np.stack([np.sin(2 * np.pi * 220.0 * tt) + 0.3 * np.sin(2 * np.pi * 440.0 * tt),
               np.sin(2 * np.pi * (150.0 + 400.0 * tt) * tt) * (tt < 0.3) + 0.01 * np.random.default_rng(3).standard_normal(9000)]).astype(np.float32)
out["pitch/input"] = yp
f0, vo = ref.pitch_detect_acf(mx.array(yp), sr=22050)
out["pitch/f0"], out["pitch/voiced"] = A(f0), A(vo)
f0, vo = ref.pitch_detect_acf(mx.array(yp[0]), sr=22050, fmin=80.0, fmax=800.0, frame_length=1024, hop_length=256, threshold=0.3, center=False)
out["pitch/f0_b"], out["pitch/voiced_b"] = A(f0), A(vo)
out["acf/default"] = A(ref.autocorrelation(mx.array(yp), max_lag=600))
out["acf/raw_1d"] = A(ref.autocorrelation(mx.array(yp[0, :3000]), normalize=False, center=False))
out["per/default"] = A(ref.periodicity(mx.array(yp), sr=22050))
out["per/b"] = A(ref.periodicity(mx.array(yp[0]), sr=22050, fmin=80.0, fmax=800.0, frame_length=1024, hop_length=256, center=False))
_de = ref.deemphasis(mx.array(y2), coef=0.97, return_zf=True)
out["de/default"], out["de/default_zf"] = A(_de[0]), A(_de[1])
_de = ref.deemphasis(mx.array(y2[0]), coef=0.9, zi=mx.array([0.25]), return_zf=True)
out["de/zi"], out["de/zi_zf"] = A(_de[0]), A(_de[1])
out["rs/fft_down"] = A(ref.resample(mx.array(y2), 22050, 16000))
out["rs/fft_up"] = A(ref.resample(mx.array(y2[0, :4001]), 16000, 22050, scale=True))
out["rs/fft_half"] = A(ref.resample(mx.array(y2[:, :5000]), 44100, 22050, fix=False))
out["rs/poly_1_2"] = A(ref.resample_poly(mx.array(y2), 1, 2))
out["rs/poly_3_2"] = A(ref.resample_poly(mx.array(y2[0]), 3, 2))
out["rs/poly_160_147"] = A(ref.resample_poly(mx.array(y2[:, :2000]), 160, 147))
out["rs/lin_down"] = A(ref.resample(mx.array(y2), 22050, 16000, res_type="linear"))
out["rs/lin_up_scale"] = A(ref.resample(mx.array(y2[1]), 16000, 44100, res_type="linear", fix=False, scale=True))
M = np.asarray(g["mfcc/0"]) if "mfcc/0" in g.files else np.random.default_rng(5).standard_normal((2, 13, 40)).astype(np.float32)
out["delta/input"] = M.astype(np.float32)
out["delta/w9o1"] = A(ref.delta(mx.array(M)))
out["delta/w9o2"] = A(ref.delta(mx.array(M), order=2))
out["delta/w5o1_mirror"] = A(ref.delta(mx.array(M), width=5, mode="mirror"))
out["delta/w7o1_nearest_axis1"] = A(ref.delta(mx.array(M), width=7, mode="nearest", axis=1))
out["delta/w3o1_wrap"] = A(ref.delta(mx.array(M), width=3, mode="wrap"))
out["delta/w9o1_constant"] = A(ref.delta(mx.array(M), mode="constant"))
np.savez_compressed(os.path.join(HERE, "reference_features.npz"), **out)
import json  # noqa: E402
json.dump(cases, open(os.path.join(HERE, "feature_cases.json"), "w"))
print("wrote", len(out), "arrays")
