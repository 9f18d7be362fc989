"""Pin the CPU oracle (oracle/spectral.py) before anything trusts it:
  1. the reference's own golden vectors (reference tests/test_cpp_extension.py:525-546),
  2. outputs of the reference's own Python code (tests/golden/reference_outputs.npz,
     made by tests/golden/generate_golden.py),
  3. the libraries the reference's tests use as oracle that exist here
     (torch.stft / istft, torchaudio mel, scipy windows and DCT) at the reference's tolerances.
"""
import os

import numpy as np
import pytest

from oracle import spectral as o

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# ---- 1. reference golden vectors -------------------------------------------
def test_ref_golden_pad_constant():
    x = np.ones((1, 10), np.float32)
    p = o.pad_signal(x, 3, "constant")
    assert p.shape == (1, 16)
    assert np.array_equal(p[0, :3], [0, 0, 0]) and np.array_equal(p[0, -3:], [0, 0, 0])


def test_ref_golden_pad_reflect():
    x = np.arange(10, dtype=np.float32)[None]
    p = o.pad_signal(x, 3, "reflect")
    assert np.array_equal(p[0, :3], [3, 2, 1])
    assert np.array_equal(p[0, -3:], [8, 7, 6])


# ---- 2. reference-code outputs ---------------------------------------------
def test_constants_bit_exact(golden):
    for key in golden.files:
        parts = key.split("/")
        if parts[0] == "window":
            w = o.get_window(parts[1], int(parts[2]), bool(int(parts[3])))
            assert np.array_equal(w, golden[key]), key
        elif parts[0] == "melfb":
            sr, n_fft, n_mels = int(parts[1]), int(parts[2]), int(parts[3])
            fmax = None if parts[5] == "None" else float(parts[5])
            norm = None if parts[7] == "None" else parts[7]
            fb = o.mel_filterbank(sr, n_fft, n_mels, float(parts[4]), fmax, bool(int(parts[6])), norm)
            assert np.array_equal(fb, golden[key]), key
        elif parts[0] == "dctmat":
            norm = None if parts[3] == "None" else parts[3]
            assert np.array_equal(o.dct_matrix(int(parts[1]), int(parts[2]), norm), golden[key]), key
    assert np.array_equal(o.linear_filterbank(22050, 1024, 32), golden["linfb/22050/1024/32"])
    hz = golden["hz"]
    assert np.array_equal(o.hz_to_mel(hz), golden["hz_to_mel/slaney"])
    assert np.array_equal(o.hz_to_mel(hz, True), golden["hz_to_mel/htk"])
    assert np.array_equal(o.mel_to_hz(o.hz_to_mel(hz)), golden["mel_to_hz/slaney"])
    assert np.array_equal(o.mel_to_hz(o.hz_to_mel(hz, True), True), golden["mel_to_hz/htk"])


def test_pad_frame_bit_exact(golden):
    x = golden["pad/input"]
    for mode in o.PAD_MODES:
        for pad in [0, 3, 8, 19]:
            assert np.array_equal(o.pad_signal(x, pad, mode), golden[f"pad/{mode}/{pad}"])
    f = golden["frame/input"]
    for fl, hop in [(8, 2), (16, 16), (10, 3), (50, 1)]:
        assert np.array_equal(o.frame_signal(f, fl, hop), golden[f"frame/{fl}/{hop}"])


def test_stft_istft_match_reference_code(golden, cases):
    y = golden["stft/input"]
    for i, kw in enumerate(cases["stft"]):
        yin = y[:, :300] if kw.get("hop_length") == 1 else y
        S = o.stft(yin, **kw)
        ref = golden[f"stft/{i}"]
        assert S.shape == ref.shape
        assert np.abs(S - ref).max() <= 1e-6 * np.abs(ref).max(), (i, kw)
        ikw = {k: v for k, v in kw.items() if k != "pad_mode"}
        r = o.istft(ref, **ikw)
        assert r.shape == golden[f"istft/{i}"].shape
        np.testing.assert_allclose(r, golden[f"istft/{i}"], atol=2e-6)
        if kw.get("center", True):
            r = o.istft(ref, length=yin.shape[1], **ikw)
            np.testing.assert_allclose(r, golden[f"istft_len/{i}"], atol=2e-6)
    S1 = golden["stft1d"]
    assert np.abs(o.stft(y[0], 512, 128) - S1).max() <= 1e-6 * np.abs(S1).max()
    for L in [5000, 6000, 7000]:
        np.testing.assert_allclose(o.istft(S1, 128, length=L), golden[f"istft1d_len/{L}"], atol=2e-6)
    Snc = o.stft(y, 512, 128, center=False)
    for L in [4000, 6500]:
        np.testing.assert_allclose(o.istft(Snc, 128, center=False, length=L),
                                   golden[f"istft_nc_len/{L}"], atol=2e-6)
    np.testing.assert_allclose(o.magnitude(S1), golden["magnitude"], rtol=1e-6)
    np.testing.assert_allclose(o.phase(S1), golden["phase"], atol=1e-6)


def test_mel_db_mfcc_match_reference_code(golden, cases):
    y = golden["stft/input"]
    for i, kw in enumerate(cases["mel"]):
        M = o.melspectrogram(y, **kw)
        ref = golden[f"mel/{i}"]
        assert np.abs(M - ref).max() <= 2e-6 * np.abs(ref).max(), (i, kw)
        np.testing.assert_allclose(o.power_to_db(ref), golden[f"db_default/{i}"], atol=1e-4)
        np.testing.assert_allclose(o.power_to_db(ref, ref=np.max), golden[f"db_refmax/{i}"], atol=1e-4)
        np.testing.assert_allclose(o.power_to_db(ref, ref=np.max, top_db=None),
                                   golden[f"db_refmax_notop/{i}"], atol=1e-4)
        np.testing.assert_allclose(o.power_to_db(ref, ref=0.5, amin=1e-5, top_db=60.0),
                                   golden[f"db_ref05_amin/{i}"], atol=1e-4)
    amp = np.abs(golden["stft1d"])
    np.testing.assert_allclose(o.amplitude_to_db(amp), golden["ampdb"], atol=1e-4)
    np.testing.assert_allclose(o.amplitude_to_db(amp, ref=np.max, top_db=None), golden["ampdb_refmax"], atol=1e-4)
    np.testing.assert_allclose(o.db_to_power(golden["dbv"], 2.0), golden["db_to_power"], rtol=1e-6)
    np.testing.assert_allclose(o.db_to_amplitude(golden["dbv"]), golden["db_to_amplitude"], rtol=1e-6)
    for i, kw in enumerate(cases["mfcc"]):
        np.testing.assert_allclose(o.mfcc(y, **kw), golden[f"mfcc/{i}"], atol=2e-4, rtol=1e-5)
    x = golden["dct/input"]
    np.testing.assert_allclose(o.dct(x), golden["dct/ortho"], atol=1e-5)
    np.testing.assert_allclose(o.dct(x, n=20, norm=None), golden["dct/n20_none"], atol=1e-4)
    np.testing.assert_allclose(o.dct(x, axis=1, n=5), golden["dct/axis1"], atol=1e-5)


def test_griffinlim_matches_reference_code(golden):
    S = golden["gl/S"]
    np.testing.assert_allclose(o.griffinlim(S, 8, 128, random_state=0), golden["gl/random8"], atol=1e-4)
    np.testing.assert_allclose(o.griffinlim(S, 4, 128, init="zeros", momentum=0.0),
                               golden["gl/zeros4_m0"], atol=1e-4)
    np.testing.assert_allclose(o.griffinlim(S, 3, 128, random_state=1, length=4000), golden["gl/len"], atol=1e-4)
    np.testing.assert_allclose(o.griffinlim(S[0], 2, 128, random_state=2), golden["gl/1d"], atol=1e-4)


# ---- 3. the reference tests' third-party oracles that exist here ------------
@pytest.mark.parametrize("n_fft,hop", [(512, 128), (1024, 256), (2048, 512), (400, 160)])
@pytest.mark.parametrize("pad_mode", ["constant", "reflect"])
def test_stft_vs_torch(random_signal, n_fft, hop, pad_mode):
    torch = pytest.importorskip("torch")
    y = random_signal
    ref = torch.stft(torch.from_numpy(y), n_fft, hop, window=torch.hann_window(n_fft),
                     center=True, pad_mode=pad_mode, return_complex=True).numpy()
    S = o.stft(y, n_fft, hop, pad_mode=pad_mode)
    # reference tolerance: tests/test_torchaudio_crossval.py:26-78 (rtol=atol=1e-4)
    np.testing.assert_allclose(S, ref, rtol=1e-4, atol=1e-4)
    S64 = o.stft(y, n_fft, hop, pad_mode=pad_mode, dtype=np.float64)
    assert np.abs(S64 - ref).max() <= 1e-5 * np.abs(ref).max()


def test_istft_round_trip_and_torch(random_signal):
    torch = pytest.importorskip("torch")
    y = random_signal
    S = o.stft(y, 2048, 512)
    r = o.istft(S, 512, length=len(y))
    assert np.abs(r[1:] - y[1:]).max() < 1e-5  # sample 0 has zero window sum (hann) -> 0
    rt = torch.istft(torch.from_numpy(S), 2048, 512, window=torch.hann_window(2048), length=len(y)).numpy()
    assert np.abs(r[1024:-1024] - rt[1024:-1024]).max() < 1e-5


def test_mel_vs_torchaudio(random_signal):
    torch = pytest.importorskip("torch")
    ta = pytest.importorskip("torchaudio")
    y = random_signal
    tr = ta.transforms.MelSpectrogram(22050, n_fft=2048, hop_length=512, n_mels=128, center=True,
                                      pad_mode="constant", norm="slaney", mel_scale="slaney", power=2.0)
    ref = tr(torch.from_numpy(y)).numpy()
    M = o.melspectrogram(y)
    np.testing.assert_allclose(M, ref, rtol=1e-3, atol=1e-4 * ref.max())


@pytest.mark.parametrize("name", ["hann", "hamming", "blackman", "bartlett"])
@pytest.mark.parametrize("n", [256, 400, 1024, 2048])
@pytest.mark.parametrize("fftbins", [True, False])
def test_windows_vs_scipy(name, n, fftbins):
    sig = pytest.importorskip("scipy.signal")
    np.testing.assert_allclose(o.get_window(name, n, fftbins), sig.get_window(name, n, fftbins=fftbins),
                               rtol=1e-5, atol=1e-5)
    w = o.get_window(name, n, False)
    assert np.array_equal(w, w[::-1])  # exact symmetry (test_torchaudio_crossval.py:199-224)


def test_dct_vs_scipy():
    sf = pytest.importorskip("scipy.fft")
    x = np.random.default_rng(0).standard_normal((5, 128)).astype(np.float32)
    np.testing.assert_allclose(o.dct(x, n=40), sf.dct(x.astype(np.float64), type=2, norm="ortho")[:, :40],
                               rtol=1e-4, atol=1e-4)


def test_error_messages():
    y = np.zeros(100, np.float32)
    with pytest.raises(ValueError, match="hop_length must be positive"):
        o.stft(y, 64, 0)
    with pytest.raises(ValueError, match="must be <= n_fft"):
        o.stft(y, 64, 16, 128)
    with pytest.raises(ValueError, match="must be >= frame_length"):
        o.stft(y, 256, 64, center=False)
    with pytest.raises(ValueError, match="Unknown pad_mode"):
        o.stft(y, 64, 16, pad_mode="wrap")
    with pytest.raises(ValueError, match="Unknown window type"):
        o.get_window("kaiser", 64)
    with pytest.raises(ValueError, match="cannot exceed Nyquist"):
        o.mel_filterbank(16000, 512, fmax=9000.0)
    with pytest.raises(ValueError, match="top_db must be positive"):
        o.power_to_db(np.ones(4, np.float32), top_db=-1)
    with pytest.raises(ValueError, match="Only DCT type 2"):
        o.dct(np.ones(4, np.float32), type=3)
    with pytest.raises(ValueError, match="Unknown init"):
        o.griffinlim(np.ones((5, 3), np.float32), init="ones")


# ---------------------------------------------------------------- section 8(f): features, rms, zcr, preemphasis
def test_feature_oracle_matches_reference_code():
    """oracle/features.py against the reference's own features.py / framing.py run on the MLX stand-in
    (tests/golden/generate_golden_features.py -> reference_features.npz)."""
    import json
    from oracle import features as of
    g = np.load(os.path.join(GOLDEN, "reference_features.npz"))
    y2 = np.load(os.path.join(GOLDEN, "reference_outputs.npz"))["stft/input"]
    cases = json.load(open(os.path.join(GOLDEN, "feature_cases.json")))
    assert int(g["ncases"]) == len(cases)

    def close(a, b, rtol):
        assert a.shape == b.shape, (a.shape, b.shape)
        assert np.abs(a - b).max() <= rtol * max(np.abs(b).max(), 1e-30), np.abs(a - b).max()

    for i, kw in enumerate(cases):
        sr = kw.get("sr", 22050)
        k2 = {k: v for k, v in kw.items() if k != "sr"}
        close(of.spectral_centroid(y2, sr=sr, **k2), g[f"centroid/{i}"], 2e-6)
        close(of.spectral_bandwidth(y2, sr=sr, **k2), g[f"bandwidth/{i}"], 5e-6)
        close(of.spectral_bandwidth(y2, sr=sr, p=3.0, norm=False, **k2), g[f"bandwidth_p3/{i}"], 2e-5)
        for key, rp in (("rolloff", 0.85), ("rolloff50", 0.5)):
            got, ref = of.spectral_rolloff(y2, sr=sr, roll_percent=rp, **k2), g[f"{key}/{i}"]
            assert got.shape == ref.shape and (got != ref).mean() <= 0.01  # a float32 cumsum tie may move one bin
        close(of.spectral_flatness(y2, **k2), g[f"flatness/{i}"], 2e-5)
        close(of.spectral_flatness(y2, power=1.0, amin=1e-6, **k2), g[f"flatness_p1/{i}"], 2e-5)
        assert np.abs(of.spectral_contrast(y2, sr=sr, **k2) - g[f"contrast/{i}"]).max() <= 1e-5  # dB
        close(of.spectral_contrast(y2, sr=sr, n_bands=4, fmin=150.0, quantile=0.1, linear=True, **k2), g[f"contrast_lin/{i}"], 1e-6)
    assert np.abs(of.spectral_contrast(S=g["S1d"], sr=22050, n_fft=512) - g["contrast_S1d"]).max() <= 1e-5
    close(of.spectral_centroid(S=g["S1d"], sr=22050, n_fft=512), g["centroid_S1d"], 2e-6)
    assert np.array_equal(of.spectral_rolloff(S=g["S1d"], sr=22050, n_fft=512), g["rolloff_S1d"])
    for key in g.files:
        parts = key.split("/")
        if parts[0] in ("rms", "zcr") and len(parts) == 5:
            fl, hop, center, mode = int(parts[1]), int(parts[2]), bool(int(parts[3])), parts[4]
            fn = of.rms if parts[0] == "rms" else of.zero_crossing_rate
            close(fn(y2, fl, hop, center=center, pad_mode=mode), g[key], 2e-6)
    close(of.rms(y2[1], 1024, 256), g["rms1d"], 2e-6)
    assert np.array_equal(of.preemphasis(y2), g["pre/default"])
    o2, zf = of.preemphasis(y2, coef=0.9, zi=np.array([0.5, -0.25], np.float32), return_zf=True)
    assert np.array_equal(o2, g["pre/zi"]) and np.array_equal(zf, g["pre/zf"])
    assert np.array_equal(of.preemphasis(y2[0], coef=0.5), g["pre/1d"])
    f0, vo = of.pitch_detect_acf(g["pitch/input"], sr=22050)
    assert np.array_equal(f0, g["pitch/f0"]) and np.array_equal(vo, g["pitch/voiced"])
    f0, vo = of.pitch_detect_acf(g["pitch/input"][0], sr=22050, fmin=80.0, fmax=800.0, frame_length=1024, hop_length=256,
                                 threshold=0.3, center=False)
    assert np.array_equal(f0, g["pitch/f0_b"]) and np.array_equal(vo, g["pitch/voiced_b"])
    assert np.abs(of.autocorrelation(g["pitch/input"], max_lag=600) - g["acf/default"]).max() < 1e-6
    raw = of.autocorrelation(g["pitch/input"][0, :3000], normalize=False, center=False)
    assert np.abs(raw - g["acf/raw_1d"]).max() < 1e-6 * np.abs(raw).max()
    assert np.array_equal(of.periodicity(g["pitch/input"]), g["per/default"])
    assert np.array_equal(of.periodicity(g["pitch/input"][0], fmin=80.0, fmax=800.0, frame_length=1024, hop_length=256, center=False),
                          g["per/b"])
    out, zf = of.deemphasis(y2, 0.97)
    assert np.array_equal(out, g["de/default"]) and np.array_equal(zf, g["de/default_zf"])
    out, zf = of.deemphasis(y2[0], 0.9, zi=[0.25])
    assert np.array_equal(out, g["de/zi"]) and np.array_equal(zf, g["de/zi_zf"])
    for name, args, kw in (("rs/fft_down", (y2, 22050, 16000), {}), ("rs/fft_up", (y2[0, :4001], 16000, 22050), dict(scale=True)),
                           ("rs/fft_half", (y2[:, :5000], 44100, 22050), dict(fix=False))):
        got = of.resample_fft(*args, **kw)                      # float64 restatement vs scipy's float32 transforms
        assert got.shape == g[name].shape and np.abs(got - g[name]).max() < 2e-6 * np.abs(got).max(), name
    assert np.array_equal(of.resample_poly(y2, 1, 2), g["rs/poly_1_2"]) and np.array_equal(of.resample_poly(y2[0], 3, 2), g["rs/poly_3_2"])
    assert np.array_equal(of.resample_poly(y2[:, :2000], 160, 147), g["rs/poly_160_147"])
    assert np.array_equal(of.resample_linear(y2, 22050, 16000), g["rs/lin_down"])
    assert np.array_equal(of.resample_linear(y2[1], 16000, 44100, fix=False, scale=True), g["rs/lin_up_scale"])
    M = g["delta/input"]
    for key, kw in (("w9o1", {}), ("w9o2", dict(order=2)), ("w5o1_mirror", dict(width=5, mode="mirror")),
                    ("w7o1_nearest_axis1", dict(width=7, mode="nearest", axis=1)), ("w3o1_wrap", dict(width=3, mode="wrap")),
                    ("w9o1_constant", dict(mode="constant"))):
        assert np.array_equal(of.delta(M, **kw), g["delta/" + key]), key
    with pytest.raises(ValueError, match="roll_percent must be <= 1.0"):
        of.spectral_rolloff(y2, roll_percent=1.5)
    with pytest.raises(ValueError, match="Either y"):
        of.spectral_centroid()
