"""GPU parity tests: the CUDA path (through the Python host layer -> C ABI) against the CPU oracle
and the committed reference fixtures.  Tolerances are the ones BASELINE.json's north_star states:
pad/frame indices bit-exact; STFT and mel within 1e-5 of peak magnitude; ISTFT round trip <= 1e-5;
dB within 1e-3 dB; MFCC at the reference tests' rtol=atol=1e-4 scale."""
import os

import numpy as np
import pytest

from oracle import spectral as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def ap():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import mlx_audio_primitives_b200 as ap
    return ap


def H(t):
    return t.detach().cpu().numpy()


def rel_peak(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


# ---------------------------------------------------------------- indices: bit exact
@pytest.mark.parametrize("mode", ["constant", "reflect", "edge"])
@pytest.mark.parametrize("pad", [0, 1, 3, 19, 512])
def test_pad_signal_bit_exact(ap, mode, pad):
    L = 20 if pad < 20 else 1000
    x = np.arange(1, 3 * L + 1, dtype=np.float32).reshape(3, L)
    got = H(ap.pad_signal(x, pad, mode))
    assert np.array_equal(got, o.pad_signal(x, pad, mode))


def test_pad_reference_golden_vectors(ap):
    # reference tests/test_cpp_extension.py:525-546
    got = H(ap.pad_signal(np.arange(10, dtype=np.float32)[None], 3, "reflect"))[0]
    assert list(got[:3]) == [3, 2, 1] and list(got[-3:]) == [8, 7, 6]
    got = H(ap.pad_signal(np.ones((1, 10), np.float32), 3, "constant"))[0]
    assert list(got[:3]) == [0, 0, 0] and list(got[-3:]) == [0, 0, 0] and got.shape == (16,)


@pytest.mark.parametrize("fl,hop", [(8, 2), (16, 16), (10, 3), (50, 1), (2048, 512)])
def test_frame_bit_exact(ap, fl, hop):
    L = 50 if fl <= 50 else 9000
    x = np.arange(2 * L, dtype=np.float32).reshape(2, L)
    assert np.array_equal(H(ap.frame(x, fl, hop)), o.frame_signal(x, fl, hop))
    assert np.array_equal(H(ap.frame(x[0], fl, hop)), o.frame_signal(x[:1], fl, hop)[0])


@pytest.mark.parametrize("n_fft,hop", [(64, 16), (400, 160), (512, 128), (2048, 512), (600, 150)])
@pytest.mark.parametrize("mode", ["constant", "reflect", "edge"])
@pytest.mark.parametrize("center", [True, False])
def test_fused_stft_indices_via_ramp(ap, n_fft, hop, mode, center):
    """Rectangular window + index ramp: the DC bin of frame t is the exact integer sum of the source
    indices the fused loader picked, so a single wrong pad/frame index shows up."""
    L = 5 * n_fft + 37
    x = np.arange(L, dtype=np.float32)[None] % 97  # small integers: sums stay exact in float32
    S = H(ap.stft(x, n_fft, hop, window="rectangular", center=center, pad_mode=mode))
    xp = o.pad_signal(x, n_fft // 2, mode) if center else x
    fr = o.frame_signal(xp, n_fft, hop)
    assert S.shape == (1, n_fft // 2 + 1, fr.shape[1])
    np.testing.assert_allclose(S[0, 0].real, fr[0].sum(-1), rtol=0, atol=0.51)
    ref = o.stft(x, n_fft, hop, window="rectangular", center=center, pad_mode=mode, dtype=np.float64)
    assert rel_peak(S, ref) <= 1e-5


# ---------------------------------------------------------------- STFT values
STFT_CASES = [
    dict(n_fft=2048, hop_length=512), dict(n_fft=1024, hop_length=256, pad_mode="reflect"),
    dict(n_fft=512, hop_length=128, pad_mode="edge"), dict(n_fft=512, hop_length=128, center=False),
    dict(n_fft=400, hop_length=160), dict(n_fft=400, hop_length=160, window="hamming", pad_mode="reflect"),
    dict(n_fft=1024, hop_length=300, win_length=800, window="blackman"),
    dict(n_fft=256, hop_length=256, window="bartlett"), dict(n_fft=4096, hop_length=1024),
    dict(n_fft=128, hop_length=32), dict(n_fft=64, hop_length=1, center=False),
    dict(n_fft=2048, hop_length=333), dict(n_fft=400, hop_length=77),
    # no compiled plan -> O(n^2) DFT kernel
    dict(n_fft=600, hop_length=150), dict(n_fft=1000, hop_length=250), dict(n_fft=255, hop_length=64),
    dict(n_fft=32, hop_length=8), dict(n_fft=8192, hop_length=2048),
]


@pytest.mark.parametrize("kw", STFT_CASES, ids=lambda k: "-".join(f"{a}{b}" for a, b in k.items()))
def test_stft_matches_oracle(ap, kw):
    rng = np.random.default_rng(42)
    L = 600 if kw.get("hop_length") == 1 else 22050
    y = rng.standard_normal((3, L)).astype(np.float32)
    S = H(ap.stft(y, **kw))
    ref = o.stft(y, dtype=np.float64, **kw)
    assert S.shape == ref.shape and S.dtype == np.complex64
    assert rel_peak(S, ref) <= 1e-5, rel_peak(S, ref)
    S1 = H(ap.stft(y[1], **kw))  # 1-D input
    assert np.array_equal(S1, S[1])


def test_stft_golden_fixtures(ap, golden, cases):
    y = golden["stft/input"]
    for i, kw in enumerate(cases["stft"]):
        yin = y[:, :300] if kw.get("hop_length") == 1 else y
        ref = golden[f"stft/{i}"]
        S = H(ap.stft(yin, **kw))
        assert S.shape == ref.shape
        assert rel_peak(S, ref) <= 1e-5, (i, kw)


def test_stft_scales_and_special_inputs(ap):
    rng = np.random.default_rng(0)
    y = rng.standard_normal(8000).astype(np.float32)
    for scale in [1e-7, 1e4]:  # reference test_mathematical_properties.py:430-451
        S = H(ap.stft(y * np.float32(scale), 1024, 256))
        ref = o.stft(y * np.float32(scale), 1024, 256, dtype=np.float64)
        assert rel_peak(S, ref) <= 1e-5
    assert np.all(H(ap.stft(np.zeros(4000, np.float32), 512, 128)) == 0)
    # pure tone lands in its bin; DC signal lands in bin 0
    sr, f0 = 22050, 440.0
    tone = np.sin(2 * np.pi * f0 * np.arange(sr) / sr).astype(np.float32)
    mag = np.abs(H(ap.stft(tone, 2048, 512)))
    assert abs(int(mag[:, 10].argmax()) - round(f0 * 2048 / sr)) <= 1
    dc = np.abs(H(ap.stft(np.ones(8192, np.float32), 1024, 256)))
    assert dc[:, 8].argmax() == 0
    # linearity (reference :133-212, 1e-5)
    a, b = rng.standard_normal((2, 6000)).astype(np.float32)
    lhs = H(ap.stft(2.0 * a + 3.0 * b, 512, 128))
    rhs = 2.0 * H(ap.stft(a, 512, 128)) + 3.0 * H(ap.stft(b, 512, 128))
    assert rel_peak(lhs, rhs) <= 1e-5


def test_stft_errors(ap):
    y = np.zeros(1000, np.float32)
    with pytest.raises(ValueError, match="hop_length must be positive"):
        ap.stft(y, 256, 0)
    with pytest.raises(ValueError, match="win_length must be positive"):
        ap.stft(y, 256, 64, 0)
    with pytest.raises(ValueError, match="must be <= n_fft"):
        ap.stft(y, 256, 64, 512)
    with pytest.raises(ValueError, match="should typically be <= n_fft"):
        ap.stft(y, 256, 512)
    with pytest.raises(ValueError, match="must be >= frame_length"):
        ap.stft(y, 2048, 512, center=False)
    with pytest.raises(ValueError, match="Unknown pad_mode"):
        ap.stft(y, 256, 64, pad_mode="wrap")
    with pytest.raises(ValueError, match="Unknown window type"):
        ap.stft(y, 256, 64, window="kaiser")
    with pytest.raises(ValueError, match="must match n_fft"):
        ap.get_window(np.ones(100, np.float32), 256)
    with pytest.raises(TypeError):
        ap.get_window(3.0, 256)


# ---------------------------------------------------------------- ISTFT
@pytest.mark.parametrize("n_fft,hop", [(64, 16), (128, 32), (256, 64), (400, 160), (400, 100), (512, 128),
                                       (1024, 256), (2048, 512), (2048, 1024), (4096, 1024), (600, 150)])
def test_round_trip(ap, n_fft, hop):
    """reference tests/test_stft.py:122-178, NUMERICAL_ACCURACY.md:12,79: error <= 1e-5"""
    rng = np.random.default_rng(42)
    y = rng.standard_normal((2, 22050)).astype(np.float32)
    S = ap.stft(y, n_fft, hop)
    r = H(ap.istft(S, hop, length=y.shape[1]))
    assert r.shape == y.shape
    # sample 0 has zero window sum for hann -> reconstructed as 0, like the reference
    assert np.abs(r[:, 1:] - y[:, 1:]).max() <= 1e-5
    r2 = H(ap.istft(S, hop))  # natural length
    n = r2.shape[1]
    assert n == (S.shape[-1] - 1) * hop
    assert np.abs(r2[:, 1:] - y[:, 1:n]).max() <= 1e-5


def _assert_istft_close(got, ref, kw, T, length=None):
    """1e-5 where the window-sum normaliser is well conditioned.  Where sum(w^2) is tiny (first/last
    samples without centring, frame edges when hop == n_fft) the division amplifies float32 rounding
    of the inverse FFT by ~1/w, so those samples are bounded by that amplification instead."""
    n_fft = kw["n_fft"]
    hop, win = o._resolve(n_fft, kw.get("hop_length"), kw.get("win_length"))
    center = kw.get("center", True)
    n_ola = (length + n_fft if center else length) if length is not None else n_fft + (T - 1) * hop
    wss = o.window_sumsquare(o.padded_window(kw.get("window", "hann"), win, n_fft), T, hop, n_ola)
    if center:
        wss = wss[n_fft // 2:]
    wss = wss[: ref.shape[-1]]
    wss = np.concatenate([wss, np.full(ref.shape[-1] - wss.shape[0], wss.max(), wss.dtype)])
    ok = wss >= 1e-2 * wss.max()
    err = np.abs(got - ref)
    if not ok.all():
        amp = np.sqrt(wss[~ok]) / np.maximum(wss[~ok], 1e-8)  # d(y)/d(frame value), one covering frame
        assert np.all(err[..., ~ok] <= 2e-6 * amp + 1e-5), kw
    assert err[..., ok].max() <= 1e-5, (kw, err[..., ok].max())


def test_istft_matches_oracle_on_golden(ap, golden, cases):
    for i, kw in enumerate(cases["stft"]):
        S = golden[f"stft/{i}"]
        ikw = {k: v for k, v in kw.items() if k != "pad_mode"}
        ref = golden[f"istft/{i}"]
        got = H(ap.istft(S, **ikw))  # (B, F, T)-contiguous input: exercises the transposing path
        assert got.shape == ref.shape, (i, kw)
        _assert_istft_close(got, ref, kw, S.shape[-1])
        if kw.get("center", True):
            L = 300 if kw.get("hop_length") == 1 else 6000
            got = H(ap.istft(S, length=L, **ikw))
            _assert_istft_close(got, golden[f"istft_len/{i}"], kw, S.shape[-1], length=L)
    S1 = golden["stft1d"]
    for L in [5000, 6000, 7000]:
        got = H(ap.istft(S1, 128, length=L))
        assert got.shape == (L,)
        _assert_istft_close(got, golden[f"istft1d_len/{L}"], dict(n_fft=512, hop_length=128), S1.shape[-1], length=L)
    Snc = o.stft(golden["stft/input"], 512, 128, center=False)
    for L in [4000, 6500]:
        got = H(ap.istft(Snc, 128, center=False, length=L))
        _assert_istft_close(got, golden[f"istft_nc_len/{L}"], dict(n_fft=512, hop_length=128, center=False),
                            Snc.shape[-1], length=L)


def test_istft_n_fft_mismatch_and_short(ap):
    rng = np.random.default_rng(3)
    S = (rng.standard_normal((2, 200, 9)) + 1j * rng.standard_normal((2, 200, 9))).astype(np.complex64)
    for n_fft in [512, 256]:  # spectrum zero-padded / cropped by irfft(n=n_fft) (reference stft.py:295)
        got = H(ap.istft(S, hop_length=64, n_fft=n_fft))
        ref = o.istft(S, hop_length=64, n_fft=n_fft, dtype=np.float64)
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max())
    one = ap.istft(S[:, :129, :1], hop_length=64)  # T = 1 -> empty (reference stft.py:322-329)
    assert tuple(one.shape) == (2, 0)
    with pytest.raises(ValueError, match="2D or 3D"):
        ap.istft(np.zeros(5, np.complex64))


def test_overlap_add_primitive(ap):
    rng = np.random.default_rng(5)
    fr = rng.standard_normal((2, 9, 64)).astype(np.float32)
    w = o.get_window("hann", 64)
    got = H(ap.overlap_add(fr, w, 16, 64 + 8 * 16))
    np.testing.assert_allclose(got, o.overlap_add(fr, w, 16, 64 + 8 * 16), atol=1e-5)


def test_magnitude_phase(ap, golden):
    S1 = golden["stft1d"]
    np.testing.assert_allclose(H(ap.magnitude(S1)), np.abs(S1), rtol=1e-6, atol=1e-6)  # reference tol 1e-6
    np.testing.assert_allclose(H(ap.phase(S1)), np.angle(S1), atol=1e-5)                # reference tol 1e-5
    y = np.random.default_rng(1).standard_normal(4000).astype(np.float32)
    S = ap.stft(y, 512, 128)  # transposed view in: same logical layout out
    m = ap.magnitude(S)
    assert m.shape == S.shape
    np.testing.assert_allclose(H(m), np.abs(H(S)), rtol=1e-6, atol=1e-6)
    assert ap.check_nola("hann", 512, 2048) is True


# ---------------------------------------------------------------- windows / filterbanks on device
def test_constants_bit_exact_on_device(ap, golden):
    for key in golden.files:
        parts = key.split("/")
        if parts[0] == "window":
            assert np.array_equal(H(ap.get_window(parts[1], int(parts[2]), bool(int(parts[3])))), golden[key]), key
        elif parts[0] == "melfb":
            fmax = None if parts[5] == "None" else float(parts[5])
            norm = None if parts[7] == "None" else parts[7]
            fb = ap.mel_filterbank(int(parts[1]), int(parts[2]), int(parts[3]), float(parts[4]), fmax,
                                   bool(int(parts[6])), norm)
            assert np.array_equal(H(fb), golden[key]), key
        elif parts[0] == "dctmat":
            norm = None if parts[3] == "None" else parts[3]
            assert np.array_equal(H(ap.dct_matrix(int(parts[1]), int(parts[2]), norm)), golden[key]), key
    assert np.array_equal(H(ap.linear_filterbank(22050, 1024, 32)), golden["linfb/22050/1024/32"])
    for key in [k for k in golden.files if k.startswith("barkfb/")]:  # reference filterbanks.py:159-231, bit for bit
        _, sr, n_fft, nb, fmin, fmax, formula, norm = key.split("/")
        fb = ap.bark_filterbank(int(sr), int(n_fft), int(nb), float(fmin), None if fmax == "None" else float(fmax), formula,
                                None if norm == "None" else norm)
        assert np.array_equal(H(fb), golden[key]), key
    assert ap.get_window("hann", 512) is ap.get_window("HANN", 512)  # device-resident cache hit
    with pytest.raises(ValueError, match="cannot exceed Nyquist"):
        ap.mel_filterbank(16000, 512, fmax=9000.0)
    with pytest.raises(ValueError, match="n_mels must be positive"):
        ap.mel_filterbank(16000, 512, n_mels=0)


# ---------------------------------------------------------------- mel / dB / MFCC
def test_mel_db_mfcc_golden(ap, golden, cases):
    y = golden["stft/input"]
    for i, kw in enumerate(cases["mel"]):
        ref = golden[f"mel/{i}"]
        M = ap.melspectrogram(y, **kw)
        assert tuple(M.shape) == ref.shape
        ref64 = o.melspectrogram(y, dtype=np.float64, **kw)
        tol = 1e-5 if kw.get("power", 2.0) in (1.0, 2.0) else 3e-5  # powf adds a couple of ulp
        assert rel_peak(H(M), ref64) <= tol, (i, kw, rel_peak(H(M), ref64))
        # dB within 1e-3 dB of the float64 restatement applied to OUR mel (isolates the dB step) ...
        M64 = H(M).astype(np.float64)
        for name, args in [("db_default", {}), ("db_refmax", dict(ref=np.max)),
                           ("db_refmax_notop", dict(ref=np.max, top_db=None)),
                           ("db_ref05_amin", dict(ref=0.5, amin=1e-5, top_db=60.0))]:
            targs = dict(args)
            if "ref" in targs and callable(targs["ref"]):
                targs["ref"] = torch.max
            got = H(ap.power_to_db(M, **targs))
            want = o.power_to_db(M64, dtype=np.float64, **args)
            assert np.abs(got - want).max() <= 1e-3, (i, name)
            # ... and end to end against the reference-code fixture where mel is well conditioned
            mask = golden[f"mel/{i}"] > 1e-6 * golden[f"mel/{i}"].max()
            assert np.abs(got - golden[f"{name}/{i}"])[mask].max() <= 2e-3, (i, name)
    for i, kw in enumerate(cases["mfcc"]):
        got = H(ap.mfcc(y, **kw))
        ref = o.mfcc(y, dtype=np.float64, **kw)
        assert got.shape == ref.shape
        np.testing.assert_allclose(got, ref, rtol=1e-4, atol=2e-3, err_msg=str(kw))
        np.testing.assert_allclose(got, golden[f"mfcc/{i}"], rtol=1e-3, atol=5e-3, err_msg=str(kw))


def test_mel_1d_powers_and_batch_consistency(ap):
    rng = np.random.default_rng(7)
    y = rng.standard_normal((5, 16000)).astype(np.float32)
    M = H(ap.melspectrogram(y, sr=16000, n_fft=400, hop_length=160, n_mels=80))
    for b in range(5):  # clip alone == clip in batch (bit exact: same kernel, same order)
        assert np.array_equal(H(ap.melspectrogram(y[b], sr=16000, n_fft=400, hop_length=160, n_mels=80)), M[b])
    # odd frame count / tile edges: T not a multiple of the tile
    for L in [400, 401, 559, 560, 5119, 5120, 5121]:
        got = H(ap.melspectrogram(y[0, :L], sr=16000, n_fft=400, hop_length=160, n_mels=80))
        ref = o.melspectrogram(y[0, :L], sr=16000, n_fft=400, hop_length=160, n_mels=80, dtype=np.float64)
        assert got.shape == ref.shape and rel_peak(got, ref) <= 1e-5, L


def test_db_family(ap, golden):
    amp = np.abs(golden["stft1d"])
    assert np.abs(H(ap.amplitude_to_db(amp)) - golden["ampdb"]).max() <= 1e-3
    assert np.abs(H(ap.amplitude_to_db(amp, ref=torch.max, top_db=None)) - golden["ampdb_refmax"]).max() <= 1e-3
    np.testing.assert_allclose(H(ap.db_to_power(golden["dbv"], 2.0)), golden["db_to_power"], rtol=1e-5)
    np.testing.assert_allclose(H(ap.db_to_amplitude(golden["dbv"])), golden["db_to_amplitude"], rtol=1e-5)
    # round trip and a non-max callable ref
    x = np.random.default_rng(2).random((3, 40, 50)).astype(np.float32) + 1e-3
    rt = H(ap.db_to_power(ap.power_to_db(x, top_db=None)))
    np.testing.assert_allclose(rt, x, rtol=1e-5)
    got = H(ap.power_to_db(x, ref=torch.median, top_db=None))
    med = float(torch.median(torch.from_numpy(x)))
    np.testing.assert_allclose(got, o.power_to_db(x, ref=med, top_db=None, dtype=np.float64), atol=1e-3)
    zeros = H(ap.power_to_db(np.zeros((4, 4), np.float32)))
    assert np.all(zeros == zeros[0, 0]) and np.isfinite(zeros).all()
    with pytest.raises(ValueError, match="top_db must be positive"):
        ap.power_to_db(x, top_db=0)


@pytest.mark.parametrize("n_fft,hop,n_mels,sr", [(400, 160, 80, 16000), (1024, 256, 64, 22050), (600, 150, 40, 16000),
                                                 (2048, 512, 128, 22050), (4096, 1024, 96, 44100), (64, 16, 12, 8000), (256, 64, 33, 8000)])
@pytest.mark.parametrize("ref", [1.0, 0.37, "max"])
def test_logmel_plan_floor_paths_same_bits(ap, n_fft, hop, n_mels, sr, ref):
    """LogMelPlan fuses dB into the mel kernel and applies top_db block-wise (only 64-frame blocks whose
    minimum is below the floor are rewritten); the host-buffer entry copies results back speculatively and
    re-copies rewritten chunks.  Both must give the bits of melspectrogram -> power_to_db, on material
    where the floor bites in some blocks (silence, fades) and not in others."""
    rng = np.random.default_rng(n_fft + n_mels)
    B, L = 11, 9 * 64 * hop + 123
    y = rng.standard_normal((B, L)).astype(np.float32)
    y[0, : L // 2] = 0.0                      # digital silence: amin-clamped, far below the floor
    y[3] *= np.linspace(1.0, 1e-6, L, dtype=np.float32)  # fade: the floor bites only in the tail blocks
    y[5] = 0.0
    y[7] *= 1e-5
    y[9, 2000:2000 + 64 * hop] = 0.0
    yt = torch.from_numpy(y).cuda()
    kw = dict(sr=sr, n_fft=n_fft, hop_length=hop, n_mels=n_mels)
    for top_db in (80.0, 25.0, None):
        want = ap.power_to_db(ap.melspectrogram(yt, **kw), ref=(torch.max if ref == "max" else ref), top_db=top_db)
        plan = ap.LogMelPlan(B, L, ref=ref, top_db=top_db, **kw)
        for _ in range(2):  # twice: the peak slots and block minima must re-arm themselves
            assert torch.equal(plan(yt), want)
        out_h = torch.empty(tuple(want.shape), dtype=torch.float32).pin_memory()
        plan.run_host(torch.from_numpy(y).pin_memory(), out_h)  # pinned: raised values are patched in place over PCIe
        assert torch.equal(out_h, want.cpu())
        out_p = torch.empty(tuple(want.shape), dtype=torch.float32)  # pageable: rewritten blocks are re-copied
        plan.run_host(torch.from_numpy(y), out_p)
        assert torch.equal(out_p, want.cpu())
        if top_db is not None:
            assert float(want.max()) - float(want.min()) <= top_db + 1e-3
            assert float((want == want.min()).float().mean()) > 0.05  # the floor really bit


def test_dct(ap, golden):
    x = golden["dct/input"]
    np.testing.assert_allclose(H(ap.dct(x)), golden["dct/ortho"], atol=1e-4)
    np.testing.assert_allclose(H(ap.dct(x, n=20, norm=None)), golden["dct/n20_none"], atol=1e-3)
    np.testing.assert_allclose(H(ap.dct(x, axis=1, n=5)), golden["dct/axis1"], atol=1e-4)
    with pytest.raises(ValueError, match="Only DCT type 2"):
        ap.dct(x, type=3)
    with pytest.raises(ValueError, match="n_mfcc must be positive"):
        ap.mfcc(np.zeros(4096, np.float32), n_mfcc=0)
    # S= path: caller-supplied log-mel, no dB step (reference tests/test_mfcc.py:72)
    logmel = np.random.default_rng(4).standard_normal((2, 40, 30)).astype(np.float32)
    np.testing.assert_allclose(H(ap.mfcc(S=logmel, n_mfcc=13)), o.mfcc(S=logmel, n_mfcc=13, dtype=np.float64),
                               rtol=1e-4, atol=1e-4)


# ---------------------------------------------------------------- Griffin-Lim
def test_griffinlim_matches_reference_code(ap, golden):
    S = golden["gl/S"]
    got = H(ap.griffinlim(S, n_iter=8, hop_length=128, random_state=0))
    assert got.shape == golden["gl/random8"].shape
    assert np.abs(got - golden["gl/random8"]).max() <= 2e-3
    got = H(ap.griffinlim(S, n_iter=4, hop_length=128, init="zeros", momentum=0.0))
    assert np.abs(got - golden["gl/zeros4_m0"]).max() <= 1e-3
    got = H(ap.griffinlim(S, n_iter=3, hop_length=128, random_state=1, length=4000))
    assert got.shape == (2, 4000) and np.abs(got - golden["gl/len"]).max() <= 1e-3
    got = H(ap.griffinlim(S[0], n_iter=2, hop_length=128, random_state=2))
    assert np.abs(got - golden["gl/1d"]).max() <= 1e-3
    a = H(ap.griffinlim(S, n_iter=4, hop_length=128, random_state=5))
    b = H(ap.griffinlim(S, n_iter=4, hop_length=128, random_state=5))
    assert np.array_equal(a, b)  # deterministic kernels: same seed -> same bits


def test_device_rng_is_numpy_pcg64(ap):
    """The on-device phase init reproduces np.random.default_rng(seed).uniform(-pi, pi) bit for bit."""
    from mlx_audio_primitives_b200.griffinlim import _uniform_phase
    for seed, shape in [(0, (2, 257, 33)), (123, (1, 513, 7)), (7, (3, 5, 1000))]:
        got = H(_uniform_phase(np.random.default_rng(seed), shape, torch.device("cuda")))
        want = np.random.default_rng(seed).uniform(-np.pi, np.pi, shape).astype(np.float32)
        assert np.array_equal(got, want)
    g = np.random.default_rng(5)
    a = H(_uniform_phase(g, (1000,), torch.device("cuda")))
    b = g.uniform(-np.pi, np.pi, 10)  # the caller's generator was advanced past what the device consumed
    ref = np.random.default_rng(5).uniform(-np.pi, np.pi, 1010)
    assert np.array_equal(a, ref[:1000].astype(np.float32)) and np.array_equal(b, ref[1000:])


def test_fused_random_start_matches_draw_transpose_polar(ap):
    """S * exp(i * uniform) in one kernel (phases drawn in logical (B, F, T) order, written at their physical
    (B, T, F) positions) == NumPy's stream, transposed, through cos / sin: bit for bit against the unfused kernels,
    and against NumPy to float32 rounding of cos / sin."""
    from mlx_audio_primitives_b200._extension import _ext, check
    from mlx_audio_primitives_b200._tensor import ptr, stream_ptr
    from mlx_audio_primitives_b200.griffinlim import _polar, _uniform_phase
    for seed, (B, F, T) in [(0, (2, 257, 33)), (9, (1, 513, 300)), (3, (3, 5, 1000)), (4, (2, 33, 1))]:
        mag = torch.rand((B, T, F), device="cuda") + 0.1  # physical layout
        st = np.random.default_rng(seed).bit_generator.state["state"]
        m64 = (1 << 64) - 1
        out = torch.empty((B, T, F, 2), device="cuda")
        check(_ext.mlxa_pcg64_polar_f32(st["state"] >> 64, st["state"] & m64, st["inc"] >> 64, st["inc"] & m64, -np.pi, np.pi,
                                        ptr(mag), B, F, T, ptr(out), stream_ptr(mag)), "pcg64_polar")
        ang = _uniform_phase(np.random.default_rng(seed), (B, F, T), torch.device("cuda"))
        want = torch.view_as_real(_polar(mag.transpose(1, 2).contiguous(), ang)).permute(0, 2, 1, 3)  # -> (B, T, F, 2)
        assert torch.equal(out, want.contiguous())
        a = np.random.default_rng(seed).uniform(-np.pi, np.pi, (B, F, T)).transpose(0, 2, 1)
        ref = H(mag).astype(np.float64)[..., None] * np.stack([np.cos(a), np.sin(a)], -1)
        assert np.abs(H(out) - ref).max() <= 1e-6
    # the public entry point takes the fused path and leaves a caller's generator where the reference would
    S = torch.rand((2, 65, 40), device="cuda") + 0.1
    g = np.random.default_rng(11)
    ap.griffinlim(S, n_iter=1, hop_length=32, random_state=g)
    assert g.uniform() == np.random.default_rng(11).uniform(size=2 * 65 * 40 + 1)[-1]


def test_griffinlim_quality_and_errors(ap):
    """reference tests/test_griffinlim.py:31,100-121: spectral MSE thresholds per iteration count"""
    t = np.arange(22050) / 22050.0  # chirp + noise, the reference's benchmark signal (benchmarks/utils.py:92-115)
    y = (np.sin(2 * np.pi * (100 + 1000 * t) * t) + 0.1 * np.random.default_rng(42).standard_normal(22050)).astype(np.float32)
    S = np.abs(o.stft(y, 1024, 256))
    for n_iter, thr in [(16, 10.0), (32, 5.0)]:
        r = H(ap.griffinlim(S, n_iter=n_iter, hop_length=256, random_state=0))
        S2 = np.abs(o.stft(r, 1024, 256))
        T = min(S.shape[1], S2.shape[1])
        assert np.mean((S[:, :T] - S2[:, :T]) ** 2) < thr
    with pytest.raises(ValueError, match="n_iter must be positive"):
        ap.griffinlim(S, n_iter=0)
    with pytest.raises(ValueError, match="momentum must be < 1.0"):
        ap.griffinlim(S, momentum=1.0)
    with pytest.raises(ValueError, match="momentum must be >= 0.0"):
        ap.griffinlim(S, momentum=-0.1)
    with pytest.raises(ValueError, match="Unknown init"):
        ap.griffinlim(S, init="ones")
    ang, reb, err = ap.griffinlim_iter(S, np.zeros_like(S), 256, 1024, 1024)
    assert tuple(ang.shape) == S.shape and float(err) >= 0
    # one iteration with momentum against the same step written out (reference griffinlim.py:199-284)
    St = torch.from_numpy(np.asarray(S)).cuda() if not torch.is_tensor(S) else S
    tprev = torch.polar(St, torch.zeros_like(St))
    ang2, reb2, _ = ap.griffinlim_iter(S, np.zeros_like(H(St)), 256, 1024, 1024, tprev=tprev)
    new = torch.polar(St, ang2)
    want = new + 0.99 * (new - tprev)
    assert float((reb2 - want).abs().max()) <= 1e-5 * float(want.abs().max())


def test_array_windows_are_not_aliased(ap):
    """Two different host windows of equal length, back to back: each upload is freed on return and the allocator
    hands the same block to the next one, so a (pointer, version) cache key would serve the first window's
    contents to the second call.  Also a caller-owned CUDA window modified in place."""
    rng = np.random.default_rng(7)
    y = rng.standard_normal(6000).astype(np.float32)
    w1 = np.hanning(512).astype(np.float32)
    w2 = np.hamming(512).astype(np.float32)
    for w in (w1, w2, w1, w2):
        got = H(ap.stft(y, 512, 128, window=w))
        ref = o.stft(y, 512, 128, window=w, dtype=np.float64)
        assert np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max()
        M = H(ap.melspectrogram(y, sr=16000, n_fft=512, hop_length=128, n_mels=40, window=w))
        Mr = o.melspectrogram(y, sr=16000, n_fft=512, hop_length=128, n_mels=40, window=w, dtype=np.float64)
        assert np.abs(M - Mr).max() <= 1e-5 * Mr.max()
    wd = torch.from_numpy(w1).cuda()
    a = H(ap.stft(y, 512, 128, window=wd))
    wd.copy_(torch.from_numpy(w2).cuda())  # same storage, new version
    b = H(ap.stft(y, 512, 128, window=wd))
    ref = o.stft(y, 512, 128, window=w2, dtype=np.float64)
    assert np.abs(b - ref).max() <= 1e-5 * np.abs(ref).max() and np.abs(a - b).max() > 1e-3


def test_empty_batch_returns_empty(ap):
    y = torch.zeros((0, 4000), device="cuda")
    assert tuple(ap.stft(y, 512, 128).shape) == (0, 257, 32)
    assert tuple(ap.melspectrogram(y, sr=16000, n_fft=400, hop_length=160, n_mels=80).shape) == (0, 80, 26)
    assert tuple(ap.istft(torch.zeros((0, 257, 32), dtype=torch.complex64, device="cuda"), 128).shape) == (0, 31 * 128)


def test_griffinlim_momentum_is_signal_domain_linear(ap):
    """The inverse kernel takes Griffin-Lim's momentum step through the (linear) inverse transform: with u_prev
    given it must return u + m (u - u_prev) where u is the plain inverse of the same spectrum, and hand u back."""
    from mlx_audio_primitives_b200.stft import _istft_physical, _spectrum_physical
    from mlx_audio_primitives_b200.windows import padded_window
    rng = np.random.default_rng(11)
    for n_fft, hop, L in [(1024, 256, 30000), (400, 160, 16000), (600, 150, 9000)]:  # planned pack / pair, O(n^2) path
        y = torch.from_numpy(rng.standard_normal((3, L)).astype(np.float32)).cuda()
        P = _spectrum_physical(ap.stft(y, n_fft, hop))
        win = padded_window("hann", n_fft, n_fft)
        u = _istft_physical(P, n_fft, hop, win, True, None)
        up = torch.from_numpy(rng.standard_normal(tuple(u.shape)).astype(np.float32)).cuda()
        uo = torch.empty_like(u)
        got = _istft_physical(P, n_fft, hop, win, True, None, u_prev=up, momentum=0.99, u_out=uo)
        assert torch.equal(uo, u)
        want = u + 0.99 * (u - up)
        assert float((got - want).abs().max()) <= 2e-6 * float(want.abs().max())


def test_warp_specialised_mel_kernel_matches(ap):
    """MLXA_MEL_WS=1 selects the warp-specialised n_fft = 400 kernel (fwd_mel_ws.cuh); it is an A/B variant of the
    default kernel and must produce the same log-mel (the environment switch is read once per process)."""
    import subprocess
    import sys
    code = (
        "import numpy as np, torch, sys; sys.path.insert(0, %r)\n"
        "import mlx_audio_primitives_b200 as ap\n"
        "from oracle import spectral as o\n"
        "rng = np.random.default_rng(3)\n"
        "for B, L in [(3, 16000), (2, 160 * 64 * 3 + 77), (70, 48000)]:\n"
        "    y = rng.standard_normal((B, L)).astype(np.float32) * np.logspace(-3, 0, B)[:, None].astype(np.float32)\n"
        "    M = ap.melspectrogram(torch.from_numpy(y).cuda(), sr=16000, n_fft=400, hop_length=160, n_mels=80)\n"
        "    for b in (0, B - 1):\n"
        "        ref = o.melspectrogram(y[b], sr=16000, n_fft=400, hop_length=160, n_mels=80, dtype=np.float64)\n"
        "        assert np.abs(M[b].cpu().numpy() - ref).max() <= 1e-5 * ref.max()\n"
        "    plan = ap.LogMelPlan(B, L, sr=16000, n_fft=400, hop_length=160, n_mels=80)\n"
        "    D = plan(torch.from_numpy(y).cuda())\n"
        "    assert torch.equal(D, ap.power_to_db(M))\n"
        "print('ws ok')\n" % ROOT)
    env = dict(os.environ, MLXA_MEL_WS="1")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and "ws ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
