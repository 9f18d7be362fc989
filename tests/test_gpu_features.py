"""GPU parity for the section-8(f) rows: spectral centroid / bandwidth / rolloff / flatness, RMS, zero-crossing
rate and pre-emphasis -- the CUDA path (Python host layer -> C ABI) against oracle/features.py (float64) and
the committed outputs of the reference's own code (tests/golden/reference_features.npz).  Tolerances follow
the reference's feature tests (rtol 1e-4 .. 1e-5 against librosa)."""
import json
import os

import numpy as np
import pytest

from oracle import features as of
from oracle import spectral as o

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ap():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import mlx_audio_primitives_b200 as ap
    return ap


def H(t):
    return t.detach().cpu().numpy()


def close(a, b, rtol):
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.abs(a - b).max() <= rtol * max(np.abs(b).max(), 1e-30), (np.abs(a - b).max(), np.abs(b).max())


def rolloff_close(got, ref, freq_step):
    """Rolloff is an index decision on a cumulative sum: at most one bin away, and only where the threshold falls
    within rounding of a cumulative value (a handful of frames)."""
    assert got.shape == ref.shape
    d = np.abs(got - ref)
    assert d.max() <= freq_step * 1.001 and (d > 0).mean() <= 0.01, (d.max(), (d > 0).mean())


def test_features_match_reference_fixtures(ap):
    g = np.load(os.path.join(GOLDEN, "reference_features.npz"))
    y2 = np.load(os.path.join(GOLDEN, "reference_outputs.npz"))["stft/input"]
    cases = json.load(open(os.path.join(GOLDEN, "feature_cases.json")))
    for i, kw in enumerate(cases):
        sr = kw.get("sr", 22050)
        k2 = {k: v for k, v in kw.items() if k != "sr"}
        step = sr / 2.0 / (kw["n_fft"] // 2)
        close(H(ap.spectral_centroid(y2, sr=sr, **k2)), g[f"centroid/{i}"], 2e-5)
        close(H(ap.spectral_bandwidth(y2, sr=sr, **k2)), g[f"bandwidth/{i}"], 2e-5)
        close(H(ap.spectral_bandwidth(y2, sr=sr, p=3.0, norm=False, **k2)), g[f"bandwidth_p3/{i}"], 1e-4)
        rolloff_close(H(ap.spectral_rolloff(y2, sr=sr, **k2)), g[f"rolloff/{i}"], step)
        rolloff_close(H(ap.spectral_rolloff(y2, sr=sr, roll_percent=0.5, **k2)), g[f"rolloff50/{i}"], step)
        close(H(ap.spectral_flatness(y2, **k2)), g[f"flatness/{i}"], 1e-4)
        close(H(ap.spectral_flatness(y2, power=1.0, amin=1e-6, **k2)), g[f"flatness_p1/{i}"], 1e-4)
        # dB of a mean of the SMALLEST magnitudes: a float32 FFT's absolute error (1e-7 of the peak) is a relative
        # error of up to ~1e-3 on those bins in either implementation, i.e. a few 1e-3 dB between two of them
        assert np.abs(H(ap.spectral_contrast(y2, sr=sr, **k2)) - g[f"contrast/{i}"]).max() <= 5e-3
        close(H(ap.spectral_contrast(y2, sr=sr, n_bands=4, fmin=150.0, quantile=0.1, linear=True, **k2)), g[f"contrast_lin/{i}"], 1e-5)
    assert np.abs(H(ap.spectral_contrast(S=g["S1d"], sr=22050, n_fft=512)) - g["contrast_S1d"]).max() <= 1e-4  # same S: selection is exact
    close(H(ap.spectral_centroid(S=g["S1d"], sr=22050, n_fft=512)), g["centroid_S1d"], 2e-5)
    rolloff_close(H(ap.spectral_rolloff(S=g["S1d"], sr=22050, n_fft=512)), g["rolloff_S1d"], 22050 / 2 / 256)
    for key in g.files:
        parts = key.split("/")
        if parts[0] in ("rms", "zcr") and len(parts) == 5:
            fl, hop, center, mode = int(parts[1]), int(parts[2]), bool(int(parts[3])), parts[4]
            fn = ap.rms if parts[0] == "rms" else ap.zero_crossing_rate
            close(H(fn(y2, fl, hop, center=center, pad_mode=mode)), g[key], 2e-6 if parts[0] == "rms" else 1e-7)
    close(H(ap.rms(y2[1], 1024, 256)), g["rms1d"], 2e-6)
    assert np.array_equal(H(ap.preemphasis(y2)), g["pre/default"])
    o2, zf = ap.preemphasis(y2, coef=0.9, zi=np.array([0.5, -0.25], np.float32), return_zf=True)
    assert np.array_equal(H(o2), g["pre/zi"]) and np.array_equal(H(zf), g["pre/zf"])
    assert np.array_equal(H(ap.preemphasis(y2[0], coef=0.5)), g["pre/1d"])


@pytest.mark.parametrize("n_fft,hop,sr", [(2048, 512, 22050), (400, 160, 16000), (1024, 256, 44100), (96, 24, 8000)])
def test_features_match_float64_oracle(ap, n_fft, hop, sr):
    rng = np.random.default_rng(n_fft)
    B, L = 5, 20000 + n_fft
    t = np.arange(L) / sr
    y = (np.sin(2 * np.pi * (200 + 3000 * t) * t)[None] * rng.uniform(0.1, 2.0, (B, 1)) + 0.05 * rng.standard_normal((B, L))).astype(np.float32)
    y[3, : L // 2] = 0.0  # silent frames: the 1e-10 guards decide the value
    kw = dict(n_fft=n_fft, hop_length=hop)
    f64 = np.float64
    close(H(ap.spectral_centroid(y, sr=sr, **kw)), of.spectral_centroid(y, sr=sr, dtype=f64, **kw), 1e-5)
    close(H(ap.spectral_bandwidth(y, sr=sr, **kw)), of.spectral_bandwidth(y, sr=sr, dtype=f64, **kw), 1e-5)
    close(H(ap.spectral_bandwidth(y, sr=sr, p=1.5, **kw)), of.spectral_bandwidth(y, sr=sr, p=1.5, dtype=f64, **kw), 1e-4)
    step = sr / 2.0 / (n_fft // 2)
    for rp in (0.85, 0.1, 1.0, 0.0):
        ref = of.spectral_rolloff(y, sr=sr, roll_percent=rp, dtype=f64, **kw)
        got = H(ap.spectral_rolloff(y, sr=sr, roll_percent=rp, **kw))
        live = np.abs(o.stft(y, n_fft, hop)).sum(1, keepdims=True) > 0  # all-zero frames: every bin ties
        rolloff_close(np.where(live, got, 0), np.where(live, ref, 0).astype(np.float32), step)
    close(H(ap.spectral_flatness(y, **kw)), of.spectral_flatness(y, dtype=f64, **kw), 1e-4)
    close(H(ap.spectral_flatness(y, power=1.0, amin=1e-6, **kw)), of.spectral_flatness(y, power=1.0, amin=1e-6, dtype=f64, **kw), 1e-4)
    # contrast: peaks / valleys are means of order statistics -- exact selection, so only the magnitudes' rounding
    # shows (silent frames give 0 dB in every band: peak == valley == amin-clamped)
    for ckw in (dict(), dict(n_bands=4, fmin=100.0, quantile=0.25, linear=True), dict(n_bands=3, quantile=1.0), dict(quantile=0.0)):
        if n_fft < 256 and ckw.get("n_bands", 6) > 4:
            ckw = dict(ckw, fmin=sr / 2.0 / 64)  # keep the octave ladder inside Nyquist for the small transforms
        ref, _, valley, smax = of.spectral_contrast(y, sr=sr, dtype=f64, parts=True, **kw, **ckw)
        got = H(ap.spectral_contrast(y, sr=sr, **kw, **ckw))
        assert got.shape == ref.shape
        if ckw.get("linear"):
            close(got, ref, 1e-5)
        else:
            # dB of a valley is only as good as the valley's RELATIVE accuracy: a float32 transform is exact to
            # ~1e-7 of the frame's largest magnitude, so entries whose valley is above 1e-4 of it must agree to
            # 5e-3 dB; the rest (bins at the float32 noise floor) to the reference's own test tolerance, 0.5 dB
            well = (valley >= 1e-4 * smax) | (valley == 0)
            assert np.abs(got - ref)[well].max() <= 5e-3, np.abs(got - ref)[well].max()
            assert (np.abs(got - ref) <= 0.5).mean() >= 0.999
    # a pre-computed spectrogram, in both layouts a caller can hand over
    S = ap.magnitude(ap.stft(y, **kw))                      # (B, F, T) view of the physical (B, T, F) buffer
    for Sx in (S, S.contiguous(), S[1]):
        ref = of.spectral_centroid(S=H(Sx), sr=sr, n_fft=n_fft, dtype=f64)
        close(H(ap.spectral_centroid(S=Sx, sr=sr, n_fft=n_fft)), ref, 1e-5)
    cen = ap.spectral_centroid(S=S, sr=sr, n_fft=n_fft)
    close(H(ap.spectral_bandwidth(S=S, sr=sr, n_fft=n_fft, centroid=cen * 0 + 1000.0)),
          of.spectral_bandwidth(S=H(S), sr=sr, n_fft=n_fft, centroid=np.full(H(cen).shape, 1000.0), dtype=f64), 1e-5)
    close(H(ap.spectral_flatness(S=S ** 2)), of.spectral_flatness(S=H(S) ** 2, dtype=f64), 1e-4)


@pytest.mark.parametrize("fl,hop,center,mode", [(2048, 512, True, "constant"), (400, 160, True, "edge"), (256, 64, False, "constant"),
                                                (33, 7, True, "edge")])
def test_rms_zcr_preemphasis_match_oracle(ap, fl, hop, center, mode):
    rng = np.random.default_rng(fl)
    y = rng.standard_normal((3, 9000)).astype(np.float32)
    y[1, 100:400] = 0.0
    close(H(ap.rms(y, fl, hop, center=center, pad_mode=mode)), of.rms(y, fl, hop, center=center, pad_mode=mode), 2e-6)
    assert np.array_equal(H(ap.zero_crossing_rate(y, fl, hop, center=center, pad_mode=mode)),
                          of.zero_crossing_rate(y, fl, hop, center=center, pad_mode=mode))
    for coef in (0.97, 0.0, 1.0):
        assert np.array_equal(H(ap.preemphasis(y, coef=coef)), of.preemphasis(y, coef=coef))


def test_delta_matches_reference_and_scipy(ap):
    """delta(): reference fixtures (its own code over scipy.signal.savgol_filter) and the float64 filter, every
    boundary mode, both orders, odd shapes and a non-last axis."""
    g = np.load(os.path.join(GOLDEN, "reference_features.npz"))
    M = g["delta/input"]
    for key, kw in (("w9o1", {}), ("w9o2", dict(order=2)), ("w5o1_mirror", dict(width=5, mode="mirror")),
                    ("w7o1_nearest_axis1", dict(width=7, mode="nearest", axis=1)), ("w3o1_wrap", dict(width=3, mode="wrap")),
                    ("w9o1_constant", dict(mode="constant"))):
        close(H(ap.delta(M, **kw)), g["delta/" + key], 2e-6)
    rng = np.random.default_rng(9)
    X = rng.standard_normal((3, 40, 257)).astype(np.float32)
    for kw in (dict(), dict(order=2), dict(width=21, order=2, polyorder=3), dict(width=3, mode="mirror"), dict(width=15, mode="wrap"),
               dict(width=9, mode="constant", cval=2.5), dict(width=5, mode="nearest", axis=0), dict(width=11, axis=1), dict(delta=0.5)):
        close(H(ap.delta(X, **kw)), of.delta(X.astype(np.float64), **kw), 5e-6)
    close(H(ap.delta(X[0, 0])), of.delta(X[0, 0].astype(np.float64)), 5e-6)  # 1-D
    short = X[:, :, :5]
    close(H(ap.delta(short, width=9, mode="mirror")), of.delta(short.astype(np.float64), width=9, mode="mirror"), 5e-6)  # window > axis
    with pytest.raises(ValueError, match="width must be odd"):
        ap.delta(X, width=8)
    with pytest.raises(ValueError, match="width must be >= 3"):
        ap.delta(X, width=1)
    with pytest.raises(ValueError, match="cannot exceed"):
        ap.delta(short, width=9)


def test_pitch_detect_acf(ap):
    """Fixtures of the reference's own code, then tones / chirps / noise / silence against the float64 oracle.  f0 is
    sr / (an integer lag picked by comparisons on a float32 autocorrelation): identical lags are required on at
    least 98 % of the frames, and on all frames of clean tones; a differing frame must be a neighbouring-peak
    decision (never garbage): its autocorrelation value at our lag is within 1e-3 of the oracle's at its lag."""
    g = np.load(os.path.join(GOLDEN, "reference_features.npz"))
    yp = g["pitch/input"]
    f0, vo = ap.pitch_detect_acf(yp, sr=22050)
    assert np.array_equal(H(vo), g["pitch/voiced"]) and (H(f0) == g["pitch/f0"]).mean() >= 0.97
    f0, vo = ap.pitch_detect_acf(yp[0], sr=22050, fmin=80.0, fmax=800.0, frame_length=1024, hop_length=256, threshold=0.3, center=False)
    assert np.array_equal(H(vo), g["pitch/voiced_b"]) and (H(f0) == g["pitch/f0_b"]).mean() >= 0.97
    sr = 16000
    t = np.arange(3 * sr) / sr
    rng = np.random.default_rng(12)
    y = np.stack([np.sin(2 * np.pi * 200.0 * t), np.sin(2 * np.pi * 110.0 * t) + 0.5 * np.sin(2 * np.pi * 220.0 * t + 1.0),
                  np.sin(2 * np.pi * (100.0 + 150.0 * t) * t), 0.3 * rng.standard_normal(t.size), np.zeros_like(t),
                  np.sin(2 * np.pi * 440.0 * t) * (t > 1.0) + 1e-3 * rng.standard_normal(t.size)]).astype(np.float32)
    for kw in (dict(), dict(frame_length=1024, hop_length=256, fmin=60.0, fmax=1000.0), dict(frame_length=512, hop_length=128, fmin=100.0, center=False),
               dict(frame_length=700, hop_length=300, fmin=80.0, threshold=0.5)):
        ref_f0, ref_vo, acf = of.pitch_detect_acf(y, sr=sr, return_acf=True, **kw)
        f0, vo = ap.pitch_detect_acf(y, sr=sr, **kw)
        f0, vo = H(f0), H(vo)
        assert f0.shape == ref_f0.shape and vo.dtype == bool
        assert (vo == ref_vo).mean() >= 0.99
        assert np.array_equal(f0[:2], ref_f0[:2]) and not vo[4].any()  # clean tones exact, silence unvoiced
        both = vo & ref_vo
        assert (f0[both] == ref_f0[both]).mean() >= 0.98
        bad = np.argwhere(both & (f0 != ref_f0))
        for b, k in bad:
            ours, theirs = int(round(sr / f0[b, k])), int(round(sr / ref_f0[b, k]))
            assert abs(acf[b, k, ours] - acf[b, k, theirs]) <= 1e-3, (b, k, ours, theirs)
    with pytest.raises(ValueError, match="must be less than fmax"):
        ap.pitch_detect_acf(y, fmin=500.0, fmax=100.0)
    with pytest.raises(ValueError, match="frame_length must be within"):
        ap.pitch_detect_acf(y, frame_length=4096)


def test_resample(ap):
    """resample_poly / resample(linear): reference fixtures, then SciPy / the NumPy restatement on odd shapes and axes."""
    g = np.load(os.path.join(GOLDEN, "reference_features.npz"))
    y2 = np.load(os.path.join(GOLDEN, "reference_outputs.npz"))["stft/input"]
    close(H(ap.resample_poly(y2, 1, 2)), g["rs/poly_1_2"], 2e-6)
    close(H(ap.resample_poly(y2[0], 3, 2)), g["rs/poly_3_2"], 2e-6)
    close(H(ap.resample_poly(y2[:, :2000], 160, 147)), g["rs/poly_160_147"], 2e-6)
    assert np.array_equal(H(ap.resample(y2, 22050, 16000, res_type="linear")), g["rs/lin_down"])
    assert np.array_equal(H(ap.resample(y2[1], 16000, 44100, res_type="linear", fix=False, scale=True)), g["rs/lin_up_scale"])
    rng = np.random.default_rng(4)
    X = rng.standard_normal((3, 5, 1234)).astype(np.float32)
    for up, down, axis in [(2, 1, -1), (4, 6, -1), (7, 3, 2), (1, 5, -1), (3, 2, 1), (147, 160, -1), (6, 6, -1)]:
        close(H(ap.resample_poly(X, up, down, axis=axis)), of.resample_poly(X.astype(np.float64), up, down, axis=axis), 5e-6)
    for a, b, kw in [(44100, 22050, {}), (16000, 22050, dict(fix=False)), (8000, 8001, dict(scale=True)), (48000, 16000, dict(axis=1))]:
        assert np.array_equal(H(ap.resample(X, a, b, res_type="linear", **kw)), of.resample_linear(X, a, b, **kw))
    assert ap.resample(X, 16000, 16000).shape == X.shape
    assert ap.resample(X, 16000, 8000).shape == (3, 5, 617)     # the default is the Fourier method (tested below)
    with pytest.raises(ValueError, match="Unknown res_type"):
        ap.resample(X, 16000, 8000, res_type="sinc")
    with pytest.raises(ValueError, match="up must be positive"):
        ap.resample_poly(X, 0, 2)


def test_feature_errors(ap):
    y = np.zeros(4000, np.float32)
    with pytest.raises(ValueError, match="Either y"):
        ap.spectral_centroid()
    with pytest.raises(ValueError, match="roll_percent must be <= 1.0"):
        ap.spectral_rolloff(y, roll_percent=1.5)
    with pytest.raises(ValueError, match="roll_percent must be >= 0.0"):
        ap.spectral_rolloff(y, roll_percent=-0.1)
    with pytest.raises(ValueError, match="coef must be in"):
        ap.preemphasis(y, coef=1.5)
    with pytest.raises(ValueError, match="frame_length must be positive"):
        ap.rms(y, 0, 10)
    with pytest.raises(ValueError, match="Unknown pad_mode"):
        ap.zero_crossing_rate(y, pad_mode="reflect")
    with pytest.raises(ValueError, match="n_bands must be positive"):
        ap.spectral_contrast(y, n_bands=0)
    with pytest.raises(ValueError, match="quantile"):
        ap.spectral_contrast(y, quantile=1.5)
    assert tuple(ap.spectral_contrast(y).shape) == (7, 8)
    assert tuple(ap.spectral_centroid(y).shape) == (1, 8) and float(ap.spectral_centroid(y).abs().max()) == 0.0


def test_autocorrelation_matches_reference_and_oracle():
    import mlx_audio_primitives_b200 as mb
    from oracle import features as of
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_features.npz"))
    yp = g["pitch/input"]
    got = mb.autocorrelation(torch.from_numpy(yp).cuda(), max_lag=600).cpu().numpy()
    assert got.shape == g["acf/default"].shape
    assert np.abs(got - g["acf/default"]).max() < 1e-5          # normalised: |r| <= 1 (reference tolerance 1e-4)
    raw = mb.autocorrelation(torch.from_numpy(yp[0, :3000]).cuda(), normalize=False, center=False).cpu().numpy()
    assert raw.shape == (3000,)
    assert np.abs(raw - g["acf/raw_1d"]).max() < 1e-5 * np.abs(g["acf/raw_1d"]).max()
    rng = np.random.default_rng(5)
    for B, n, lag in ((1, 1, None), (3, 129, None), (2, 5000, 257), (5, 70001, 1000), (2, 5000, None), (3, 70001, 20000),
                      (2, 300000, 2048)):
        y = (rng.standard_normal((B, n)) + 0.3).astype(np.float32)
        for normalize, center in ((True, True), (False, True), (True, False)):
            if n == 1 and center:
                continue                                       # r[0] = 0 -> 0 / 1e-10: nothing to compare
            want = of.autocorrelation(y, max_lag=lag, normalize=normalize, center=center)
            got = mb.autocorrelation(torch.from_numpy(y).cuda(), max_lag=lag, normalize=normalize, center=center).cpu().numpy()
            assert got.shape == want.shape
            # up to 1024 lags: the direct sum (float64 across chunks); beyond: two float32 transforms
            tol = 2e-6 if got.shape[1] <= 1024 else 1e-5
            assert np.abs(got - want).max() <= tol * np.abs(want).max() + 1e-12, (B, n, lag, normalize, center)
    # a strided view is a legal input; bits equal the contiguous call
    big = torch.from_numpy(rng.standard_normal((4, 9000)).astype(np.float32)).cuda()
    assert torch.equal(mb.autocorrelation(big[:, :8000], max_lag=300), mb.autocorrelation(big[:, :8000].contiguous(), max_lag=300))


def test_periodicity_matches_reference_and_oracle():
    import mlx_audio_primitives_b200 as mb
    g = np.load(os.path.join(GOLDEN, "reference_features.npz"))
    yp = g["pitch/input"]
    got = mb.periodicity(torch.from_numpy(yp).cuda(), sr=22050).cpu().numpy()
    assert got.shape == g["per/default"].shape and np.abs(got - g["per/default"]).max() < 2e-5
    got = mb.periodicity(torch.from_numpy(yp[0]).cuda(), sr=22050, fmin=80.0, fmax=800.0, frame_length=1024, hop_length=256,
                         center=False).cpu().numpy()
    assert got.shape == g["per/b"].shape and np.abs(got - g["per/b"]).max() < 2e-5
    rng = np.random.default_rng(9)
    y = rng.standard_normal((3, 20000)).astype(np.float32)
    y[1, 4000:12000] = 0.0                                     # silent frames -> exactly 0
    y[2] = np.sin(2 * np.pi * 220.0 * np.arange(20000) / 16000.0).astype(np.float32)
    for fl, hop, center in ((512, 128, True), (2048, 512, True), (400, 160, False)):
        want = of.periodicity(y, sr=16000, fmin=60.0, fmax=1000.0, frame_length=fl, hop_length=hop, center=center)
        got = mb.periodicity(torch.from_numpy(y).cuda(), sr=16000, fmin=60.0, fmax=1000.0, frame_length=fl, hop_length=hop,
                             center=center).cpu().numpy()
        assert got.shape == want.shape and np.abs(got - want).max() < 5e-5, (fl, hop, center)
        assert np.array_equal(got == 0, want == 0)
    # an empty lag range is all zeros, as in the reference
    assert not mb.periodicity(torch.from_numpy(y).cuda(), sr=16000, fmin=900.0, fmax=100.0, frame_length=512).any()
    with pytest.raises(ValueError):
        mb.periodicity(torch.from_numpy(y).cuda(), frame_length=4096)


def test_deemphasis_matches_reference_and_inverts_preemphasis():
    import mlx_audio_primitives_b200 as mb
    g = np.load(os.path.join(GOLDEN, "reference_features.npz"))
    y2 = np.load(os.path.join(GOLDEN, "reference_outputs.npz"))["stft/input"]
    out, zf = mb.deemphasis(torch.from_numpy(y2).cuda(), coef=0.97, return_zf=True)
    scale = np.abs(g["de/default"]).max()
    assert np.abs(out.cpu().numpy() - g["de/default"]).max() < 2e-6 * scale
    assert zf.shape == g["de/default_zf"].shape and np.abs(zf.cpu().numpy() - g["de/default_zf"]).max() < 2e-6 * scale
    out, zf = mb.deemphasis(torch.from_numpy(y2[0]).cuda(), coef=0.9, zi=[0.25], return_zf=True)
    assert out.shape == g["de/zi"].shape and np.abs(out.cpu().numpy() - g["de/zi"]).max() < 2e-6 * np.abs(g["de/zi"]).max()
    assert zf.shape == g["de/zi_zf"].shape and np.abs(zf.cpu().numpy() - g["de/zi_zf"]).max() < 2e-6 * np.abs(g["de/zi"]).max()
    rng = np.random.default_rng(10)
    for B, L, coef in ((1, 2, 0.97), (3, 8191, 0.5), (2, 8192, 0.0), (2, 8193, 1.0), (4, 100001, 0.97), (1, 480000, 0.95)):
        y = rng.standard_normal((B, L)).astype(np.float32)
        for zi in (None, rng.standard_normal(B).astype(np.float32)):
            want, want_zf = of.deemphasis(y, coef, zi=zi)
            got, got_zf = mb.deemphasis(torch.from_numpy(y).cuda(), coef=coef, zi=zi, return_zf=True)
            # float32 lfilter accumulates rounding along the recurrence (gain 1 / (1 - coef); a running sum at coef = 1)
            tol = 3e-7 * np.abs(want).max() * (np.sqrt(L) if coef == 1.0 else 1.0 / (1.0 - coef))
            assert np.abs(got.cpu().numpy() - want).max() <= tol, (B, L, coef, zi is None)
            assert np.abs(got_zf.cpu().numpy() - want_zf).max() <= tol, (B, L, coef)
    # round trip with preemphasis (the pair is exact in real arithmetic)
    y = torch.from_numpy(rng.standard_normal((2, 50000)).astype(np.float32)).cuda()
    back = mb.deemphasis(mb.preemphasis(y, 0.97), 0.97)
    assert (back - y).abs().max().item() < 2e-5
    with pytest.raises(ValueError):
        mb.deemphasis(y, coef=1.5)
    with pytest.raises(ValueError):
        mb.deemphasis(y[:, :1])


def test_resample_fft_matches_reference_and_oracle():
    import mlx_audio_primitives_b200 as mb
    g = np.load(os.path.join(GOLDEN, "reference_features.npz"))
    y2 = np.load(os.path.join(GOLDEN, "reference_outputs.npz"))["stft/input"]
    for name, args, kw in (("rs/fft_down", (y2, 22050, 16000), {}), ("rs/fft_up", (y2[0, :4001], 16000, 22050), dict(scale=True)),
                           ("rs/fft_half", (y2[:, :5000], 44100, 22050), dict(fix=False))):
        got = mb.resample(torch.from_numpy(np.ascontiguousarray(args[0])).cuda(), args[1], args[2], **kw).cpu().numpy()
        assert got.shape == g[name].shape, name
        assert np.abs(got - g[name]).max() < 1e-5 * np.abs(g[name]).max(), name
    rng = np.random.default_rng(21)
    # odd / even / prime lengths, both directions, lengths that cross the transform sizes, a long clip
    for B, n, a, b in ((1, 2, 1, 2), (3, 1000, 2, 1), (2, 1001, 3, 2), (2, 4099, 147, 160), (1, 65537, 1, 2), (2, 32768, 3, 1),
                       (1, 480000, 16000, 22050), (4, 220500, 22050, 16000)):
        y = rng.standard_normal((B, n)).astype(np.float32)
        for scale in (False, True):
            want = of.resample_fft(y, a, b, scale=scale)
            got = mb.resample(torch.from_numpy(y).cuda(), a, b, scale=scale).cpu().numpy()
            assert got.shape == want.shape, (B, n, a, b)
            assert np.abs(got - want).max() < 2e-5 * np.abs(want).max(), (B, n, a, b, np.abs(got - want).max())
    # many distinct lengths: the per-length table cache is flushed and rebuilt along the way
    for i in range(10):
        y = rng.standard_normal((2, 3000 + 37 * i)).astype(np.float32)
        want = of.resample_fft(y, 3, 2)
        got = mb.resample(torch.from_numpy(y).cuda(), 3, 2).cpu().numpy()
        assert got.shape == want.shape and np.abs(got - want).max() < 2e-5 * np.abs(want).max(), i
    # same length -> the input itself; other axes; deterministic
    y = torch.from_numpy(rng.standard_normal((3, 5, 700)).astype(np.float32)).cuda()
    assert mb.resample(y, 8000, 8000) is y
    r1 = mb.resample(y, 8000, 12000, axis=-1)
    assert r1.shape == (3, 5, 1050) and torch.equal(r1, mb.resample(y, 8000, 12000))
    r2 = mb.resample(y.transpose(1, 2), 8000, 12000, axis=1)
    assert torch.equal(r2, r1.transpose(1, 2))
    with pytest.raises(ValueError):
        mb.resample(y, 8000, 12000, res_type="sinc")
