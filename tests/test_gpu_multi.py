"""Multi-GPU parity (needs >= 2 visible GPUs; skipped on a one-GPU box): the peer-memory peak exchange between
the mel kernel and the dB kernel must give every rank the bits of the unsharded computation."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_peak_exchange_two_ranks():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29531", os.path.join(ROOT, "tools", "check_peak_exchange.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "bits equal: False" not in r.stdout


def test_inputs_on_another_device_than_the_current_one():
    """the C ABI launches on the CURRENT device and the constant caches are keyed on it, so an input on
    cuda:1 while cuda:0 is current used to mix a device-1 stream with device-0 launches and constants.  Every public
    entry point now runs on the device of its input."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import numpy as np
    import mlx_audio_primitives_b200 as ap
    from oracle import spectral as o
    torch.cuda.set_device(0)
    rng = np.random.default_rng(0)
    y = rng.standard_normal((3, 16000)).astype(np.float32)
    y1 = torch.from_numpy(y).to("cuda:1")
    assert torch.cuda.current_device() == 0
    M = ap.melspectrogram(y1, sr=16000, n_fft=400, hop_length=160, n_mels=80)
    D = ap.power_to_db(M)
    S = ap.stft(y1, 1024, 256)
    r = ap.istft(S, 256, length=y.shape[1])
    g = ap.griffinlim(ap.magnitude(S), n_iter=2, hop_length=256, random_state=0)
    c = ap.spectral_centroid(y1, sr=16000, n_fft=1024, hop_length=256)
    assert all(t.device == y1.device for t in (M, D, S, r, g, c)) and torch.cuda.current_device() == 0
    ref = o.melspectrogram(y, sr=16000, n_fft=400, hop_length=160, n_mels=80, dtype=np.float64)
    assert np.abs(M.cpu().numpy() - ref).max() <= 1e-5 * ref.max()
    assert np.abs(r.cpu().numpy()[:, 1:] - y[:, 1:]).max() <= 1e-5
    M0 = ap.melspectrogram(torch.from_numpy(y).to("cuda:0"), sr=16000, n_fft=400, hop_length=160, n_mels=80)
    assert torch.equal(M0.cpu(), M.cpu())
