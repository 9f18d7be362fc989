"""Seeded randomised sweep over the parameter space of the fused kernels: plan sizes and sizes
without a plan, odd hops, short windows, centre on/off, all pad modes, clip lengths around tile
boundaries, batch sizes, and inputs whose storage is NOT 16-byte aligned or whose rows are not
16-byte multiples (the TMA bulk path needs an alignment lead; edge tiles use the index path)."""
import numpy as np
import pytest

from oracle import spectral as o

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

PLANNED = [32, 64, 128, 256, 400, 480, 512, 600, 800, 1000, 1024, 1200, 1600, 2000, 2048, 3072, 4096, 8192]
N_FFTS = PLANNED + [96, 250, 777]
WINDOWS = ["hann", "hamming", "blackman", "bartlett", "rectangular"]


@pytest.fixture(scope="module")
def ap():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import mlx_audio_primitives_b200 as ap
    return ap


def H(t):
    return t.detach().cpu().numpy()


def draw(rng):
    n_fft = int(rng.choice(N_FFTS))
    hop = int(rng.choice([n_fft // 4, n_fft // 2, n_fft, max(1, n_fft // 3), max(1, n_fft // 4 + 1), 160 if n_fft >= 160 else n_fft // 2]))
    hop = max(1, min(hop, n_fft))
    win = int(rng.choice([n_fft, n_fft, max(2, n_fft - int(rng.integers(1, max(2, n_fft // 2))))]))
    center = bool(rng.integers(0, 2))
    mode = str(rng.choice(["constant", "reflect", "edge"]))
    B = int(rng.choice([1, 2, 3, 5]))
    frames = int(rng.choice([1, 2, 7, 31, 32, 33, 63, 64, 65, 70, 130]))
    L = max(n_fft if not center else n_fft // 2 + 2, (frames - 1) * hop + (0 if center else n_fft) + int(rng.integers(0, hop)))
    return dict(n_fft=n_fft, hop_length=hop, win_length=win, window=str(rng.choice(WINDOWS)), center=center, pad_mode=mode), B, L


def misaligned(rng, B, L):
    """float32 CUDA tensor (B, L) whose rows start 4*k bytes off a 16-byte boundary."""
    off = int(rng.integers(0, 4))
    base = torch.from_numpy(rng.standard_normal(B * L + off).astype(np.float32)).cuda()
    return base[off:off + B * L].view(B, L)


@pytest.mark.parametrize("seed", range(24))
def test_stft_istft_mel_sweep(ap, seed):
    rng = np.random.default_rng(1000 + seed)
    for _ in range(3):
        kw, B, L = draw(rng)
        y = misaligned(rng, B, L)
        yh = H(y)
        S = ap.stft(y, **kw)
        ref = o.stft(yh, dtype=np.float64, **kw)
        assert tuple(S.shape) == ref.shape, kw
        peak = max(np.abs(ref).max(), 1e-20)
        assert np.abs(H(S) - ref).max() <= 1e-5 * peak, (kw, B, L)
        # inverse on the reference spectrum (float64 oracle as yardstick), natural length
        ikw = {k: v for k, v in kw.items() if k != "pad_mode"}
        if S.shape[-1] > 1 or not kw["center"]:
            got = H(ap.istft(S, **ikw))
            want = o.istft(ref, dtype=np.float64, **ikw)
            assert got.shape == want.shape, kw
            if got.size:
                hop, win = kw["hop_length"], kw["win_length"]
                n_ola = kw["n_fft"] + (S.shape[-1] - 1) * hop
                wss = o.window_sumsquare(o.padded_window(kw["window"], win, kw["n_fft"]), S.shape[-1], hop, n_ola)
                if kw["center"]:
                    wss = wss[kw["n_fft"] // 2:]
                wss = wss[: got.shape[-1]]
                ok = wss >= 1e-2 * wss.max()
                assert np.abs(got - want)[:, ok].max(initial=0.0) <= 2e-5 * max(1.0, np.abs(want).max()), (kw, B, L)
        if kw["n_fft"] >= 128:
            sr = 16000
            mk = dict(sr=sr, n_mels=int(rng.choice([20, 40, 80])), power=float(rng.choice([1.0, 2.0])), **kw)
            M = ap.melspectrogram(y, **mk)
            mref = o.melspectrogram(yh, dtype=np.float64, **mk)
            assert np.abs(H(M) - mref).max() <= 1e-5 * max(mref.max(), 1e-20), mk
            peak_dev = float(ap.power_to_db(M, ref=torch.max, top_db=None).max())
            assert abs(peak_dev) <= 1e-4  # fused running max == true max


def test_batch_rows_not_multiple_of_16_bytes(ap):
    """L odd: consecutive clips start at different 16-byte phases (every TMA lead value occurs)."""
    rng = np.random.default_rng(7)
    for L in [48001, 48002, 48003]:
        y = torch.from_numpy(rng.standard_normal((6, L)).astype(np.float32)).cuda()
        for n_fft, hop in [(400, 160), (2048, 512), (1024, 256)]:
            S = H(ap.stft(y, n_fft, hop))
            ref = o.stft(H(y), n_fft, hop, dtype=np.float64)
            assert np.abs(S - ref).max() <= 1e-5 * np.abs(ref).max(), (L, n_fft)
            M = H(ap.melspectrogram(y, sr=16000, n_fft=n_fft, hop_length=hop, n_mels=64))
            mref = o.melspectrogram(H(y), sr=16000, n_fft=n_fft, hop_length=hop, n_mels=64, dtype=np.float64)
            assert np.abs(M - mref).max() <= 1e-5 * mref.max(), (L, n_fft)


def test_many_clips_persistent_schedule(ap):
    """More work items than resident CTAs, tiles of several clips interleaved in one CTA's loop."""
    rng = np.random.default_rng(11)
    y = torch.from_numpy(rng.standard_normal((300, 4000)).astype(np.float32)).cuda()
    M = H(ap.melspectrogram(y, sr=16000, n_fft=400, hop_length=160, n_mels=80))
    for b in [0, 1, 147, 148, 149, 299]:
        ref = o.melspectrogram(H(y[b]), sr=16000, n_fft=400, hop_length=160, n_mels=80, dtype=np.float64)
        assert np.abs(M[b] - ref).max() <= 1e-5 * ref.max(), b
    S = ap.stft(y, 512, 128)
    r = ap.istft(S, 128, length=4000)
    assert float((r[:, 1:] - y[:, 1:]).abs().max()) <= 1e-5


def test_many_clips_barrier_free_loops(ap):
    """The tile loops without a CTA barrier per tile (staging buffers handed over by full / empty mbarriers, split
    arrive / wait barriers around the power tile, neighbour-warp overlap-add rounds): many more (clip, tile) items
    than resident CTAs, every tile an edge tile, against torch.stft in float64 and the oracle -- and run three
    times: a race would show up as results that change from run to run."""
    rng = np.random.default_rng(23)
    y = torch.from_numpy(rng.standard_normal((600, 9000)).astype(np.float32)).cuda()
    win = torch.hann_window(2048, periodic=True, dtype=torch.float64, device="cuda")
    ref = torch.stft(y.double(), 2048, 512, window=win, center=True, pad_mode="constant", return_complex=True)
    S = ap.stft(y, 2048, 512)
    assert float((S - ref).abs().max()) <= 1e-5 * float(ref.abs().max())
    M = ap.melspectrogram(y, sr=22050, n_fft=2048, hop_length=512, n_mels=128)
    for b in [0, 147, 148, 295, 296, 599]:
        mref = o.melspectrogram(H(y[b]), sr=22050, n_fft=2048, hop_length=512, n_mels=128, dtype=np.float64)
        assert np.abs(H(M[b]) - mref).max() <= 1e-5 * mref.max(), b
    mag = ap.magnitude(S[:96])
    g = ap.griffinlim(mag, n_iter=2, hop_length=512, random_state=0)
    r = ap.istft(S, 512, length=9000)
    assert float((r[:, 1:] - y[:, 1:]).abs().max()) <= 1e-5
    for _ in range(2):
        assert torch.equal(ap.stft(y, 2048, 512), S)
        assert torch.equal(ap.melspectrogram(y, sr=22050, n_fft=2048, hop_length=512, n_mels=128), M)
        assert torch.equal(ap.griffinlim(mag, n_iter=2, hop_length=512, random_state=0), g)
        assert torch.equal(ap.istft(S, 512, length=9000), r)
    # long clips: interior tiles, both staging buffers in rotation
    y2 = torch.from_numpy(rng.standard_normal((5, 300000)).astype(np.float32)).cuda()
    for n_fft, hop in [(1024, 256), (512, 128), (4096, 1024), (64, 16), (256, 64), (2048, 512)]:
        w = torch.hann_window(n_fft, periodic=True, dtype=torch.float64, device="cuda")
        ref2 = torch.stft(y2.double(), n_fft, hop, window=w, center=True, pad_mode="constant", return_complex=True)
        S2 = ap.stft(y2, n_fft, hop)
        assert float((S2 - ref2).abs().max()) <= 1e-5 * float(ref2.abs().max()), n_fft
        r2 = ap.istft(S2, hop, length=300000)
        assert float((r2[:, 1:] - y2[:, 1:]).abs().max()) <= 1e-5, n_fft
        assert torch.equal(ap.stft(y2, n_fft, hop), S2) and torch.equal(ap.istft(S2, hop, length=300000), r2)


@pytest.mark.parametrize("n_fft", PLANNED)
def test_every_planned_size(ap, n_fft):
    """Every compiled plan (powers of two 32..8192 and the 2^a 3^b 5^c sizes) through the fused kernels: forward
    transform, mel epilogue, inverse + overlap-add round trip, Griffin-Lim projection chain, against the float64
    oracle at the north-star tolerances.  mlxa_has_fast_plan says the O(n^2) path is not what is being tested."""
    from mlx_audio_primitives_b200._extension import _ext
    assert _ext.mlxa_has_fast_plan(n_fft) == 1
    rng = np.random.default_rng(n_fft)
    for hop, center, B, frames in [(n_fft // 4, True, 3, 70), (max(1, n_fft // 3 + 1), False, 2, 37)]:
        L = (frames - 1) * hop + (0 if center else n_fft) + int(rng.integers(0, hop))
        y = misaligned(rng, B, L)
        yh = H(y)
        kw = dict(n_fft=n_fft, hop_length=hop, center=center)
        S = ap.stft(y, **kw)
        ref = o.stft(yh, dtype=np.float64, **kw)
        assert tuple(S.shape) == ref.shape
        assert np.abs(H(S) - ref).max() <= 1e-5 * np.abs(ref).max(), kw
        r = H(ap.istft(S, hop, center=center, length=L if center else None))
        want = o.istft(ref, hop, center=center, length=L if center else None, dtype=np.float64)
        n_ola = n_fft + (S.shape[-1] - 1) * hop
        wss = o.window_sumsquare(o.padded_window("hann", n_fft, n_fft), S.shape[-1], hop, n_ola)
        if center:
            wss = wss[n_fft // 2:]
        ok = np.zeros(r.shape[-1], bool)
        m = min(ok.size, wss.size)
        ok[:m] = wss[:m] >= 1e-2 * wss.max()
        assert np.abs(r - want)[:, ok].max(initial=0.0) <= 2e-5 * max(1.0, np.abs(want).max()), kw
        if n_fft >= 128:
            mk = dict(sr=16000, n_mels=40, **kw)
            M = ap.melspectrogram(y, **mk)
            mref = o.melspectrogram(yh, dtype=np.float64, **mk)
            assert np.abs(H(M) - mref).max() <= 1e-5 * mref.max(), mk
            D = ap.power_to_db(M)
            dref = o.power_to_db(mref, dtype=np.float64)
            assert np.abs(H(D) - dref).max() <= 1e-3
    if n_fft >= 64:
        y = torch.from_numpy(rng.standard_normal((2, 20 * (n_fft // 4))).astype(np.float32)).cuda()
        mag = ap.magnitude(ap.stft(y, n_fft))
        got = H(ap.griffinlim(mag, n_iter=2, n_fft=n_fft, random_state=3))
        ref = o.griffinlim(H(mag), 2, n_fft // 4, n_fft=n_fft, random_state=3)
        assert np.abs(got - ref).max() <= 2e-3 * max(1.0, np.abs(ref).max())
