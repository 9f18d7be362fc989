"""Dry run of the whole Python host layer on CPU with the C ABI stubbed out.

There is no GPU in the build container, so this test swaps the ctypes library for a recorder
that type-checks every call against the declared signatures and returns success, and lets
tensors stay on the CPU.  It catches host-side bugs (wrong argument lists, shape logic, cache
dead-locks) before a GPU box is involved.  Numerical results are NOT checked here -- the stub
computes nothing; parity lives in the ``-m gpu`` tests."""
import ctypes
import threading

import numpy as np
import pytest
import torch


class _Stream:
    cuda_stream = 0


class _StubLib:
    def __init__(self, signatures):
        self.calls = []
        for name, argtypes in signatures.items():
            setattr(self, name, self._make(name, argtypes))

    def _make(self, name, argtypes):
        def fn(*args):
            assert len(args) == len(argtypes), f"{name}: {len(args)} args, ABI declares {len(argtypes)}"
            for a, t in zip(args, argtypes):
                if t is ctypes.c_void_p:
                    assert a is None or isinstance(a, (int, ctypes.c_void_p)) or hasattr(a, "_obj"), \
                        f"{name}: pointer argument got {type(a)}"
                else:
                    t(a)  # raises TypeError on a wrong Python type
            self.calls.append(name)
            if name == "mlxa_pack_filterbank":  # out-parameter: pretend every band has 4 weights
                args[6]._obj.value = 4 * args[1]
            if name == "mlxa_plan_group":
                return 16
            return 0
        return fn

    def mlxa_last_error(self):
        return b""

    def mlxa_packed_bank_words(self, n_bands, n_wt, group):
        return (n_wt + 2 * n_bands + 2 * (-(-n_bands // group)) + 3) & ~3


@pytest.fixture
def ap(monkeypatch):
    import mlx_audio_primitives_b200 as ap
    import importlib
    _extension, _tensor, convert, framing, griffinlim, mel, mfcc, stft, windows = (
        importlib.import_module('mlx_audio_primitives_b200.' + m) for m in
        ('_extension', '_tensor', 'convert', 'framing', 'griffinlim', 'mel', 'mfcc', 'stft', 'windows'))
    stub = _StubLib(_extension.SIGNATURES)
    for mod in (_extension, convert, framing, griffinlim, mel, mfcc, stft):
        monkeypatch.setattr(mod, "_ext", stub, raising=True)
    for mod in (_tensor, windows, mel, mfcc):
        monkeypatch.setattr(mod, "require_cuda", lambda: None, raising=True)
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda device=None: _Stream())
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))
    real_to = torch.Tensor.to

    def fake_to(self, *a, **k):
        a = tuple(x for x in a if not (isinstance(x, torch.device) or (isinstance(x, str) and x.startswith("cuda"))))
        k = {kk: v for kk, v in k.items() if kk != "device"}
        return real_to(self, *a, **k) if (a or k) else self
    monkeypatch.setattr(torch.Tensor, "to", fake_to)
    windows.clear_caches(); mel.clear_caches(); mfcc._dct_device.clear(); stft._wss_cache.clear()
    ap._stub = stub
    yield ap
    windows.clear_caches(); mel.clear_caches(); mfcc._dct_device.clear(); stft._wss_cache.clear()


def _run_with_deadline(fn, seconds=20):
    box = {}
    t = threading.Thread(target=lambda: box.setdefault("r", fn()), daemon=True)
    t.start()
    t.join(seconds)
    assert not t.is_alive(), "host layer dead-locked"
    return box.get("r")


def test_every_public_entry_point_runs(ap):
    y = np.random.default_rng(0).standard_normal((2, 8000)).astype(np.float32)

    def body():
        S = ap.stft(y, 512, 128)
        assert tuple(S.shape) == (2, 257, 63) and S.dtype == torch.complex64
        assert tuple(ap.stft(y[0], 400, 160, window="hamming", pad_mode="reflect").shape) == (201, 51)
        assert tuple(ap.stft(y, 1024, 300, win_length=800, center=False).shape) == (2, 513, 24)
        assert tuple(ap.istft(S, 128).shape) == (2, 62 * 128)
        assert tuple(ap.istft(S, 128, length=8000).shape) == (2, 8000)
        assert tuple(ap.istft(np.zeros((257, 63), np.complex64), 128, center=False, length=9000).shape) == (9000,)
        assert tuple(ap.magnitude(S).shape) == tuple(S.shape) and ap.phase(S).dtype == torch.float32
        M = ap.melspectrogram(y, sr=16000, n_fft=400, hop_length=160, n_mels=80)
        assert tuple(M.shape) == (2, 80, 51)
        for kw in [{}, dict(ref=torch.max), dict(ref=torch.max, top_db=None), dict(ref=0.5, top_db=None),
                   dict(ref=torch.median)]:
            assert tuple(ap.power_to_db(M, **kw).shape) == (2, 80, 51)
        ap.amplitude_to_db(np.abs(y)); ap.db_to_power(y); ap.db_to_amplitude(y)
        assert tuple(ap.mfcc(y, n_mfcc=13, lifter=22).shape) == (2, 13, 16)
        assert tuple(ap.mfcc(S=np.zeros((40, 30), np.float32), n_mfcc=13).shape) == (13, 30)
        assert tuple(ap.dct(np.zeros((3, 7, 64), np.float32), axis=1, n=5).shape) == (3, 5, 64)
        assert tuple(ap.frame(y, 512, 128).shape) == (2, 59, 512)
        assert tuple(ap.pad_signal(y, 100, "reflect").shape) == (2, 8200)
        assert tuple(ap.overlap_add(np.zeros((2, 9, 64), np.float32), np.ones(64, np.float32), 16, 192).shape) == (2, 192)
        mag = np.abs(np.random.default_rng(1).standard_normal((2, 257, 20))).astype(np.float32)
        assert tuple(ap.griffinlim(mag, n_iter=3, hop_length=128, random_state=0).shape) == (2, 19 * 128)
        assert tuple(ap.griffinlim(mag[0], n_iter=2, hop_length=128, momentum=0.0, init="zeros", length=2000).shape) == (2000,)
        ap.griffinlim_iter(mag, np.zeros_like(mag), 128, 512, 512)
        assert tuple(ap.mel_filterbank(22050, 2048).shape) == (128, 1025)
        assert tuple(ap.linear_filterbank(22050, 1024, 32).shape) == (32, 513)
        assert tuple(ap.bark_filterbank(22050, 1024).shape) == (24, 513)
        assert ap.check_nola("hann", 512, 2048)
        return True

    assert _run_with_deadline(body)
    calls = set(ap._stub.calls)
    for must in ["mlxa_stft_f32", "mlxa_istft_f32", "mlxa_melspec_f32", "mlxa_to_db_f32", "mlxa_mfcc_tail_f32",
                 "mlxa_griffinlim_project_f32", "mlxa_polar_f32", "mlxa_window_sumsquare_f32", "mlxa_max_f32",
                 "mlxa_dct_f32", "mlxa_frame_signal_f32", "mlxa_pad_signal_f32", "mlxa_overlap_add_f32",
                 "mlxa_magnitude_f32", "mlxa_phase_f32", "mlxa_transpose_c64", "mlxa_transpose_f32", "mlxa_from_db_f32"]:
        assert must in calls, must


def test_fused_peak_is_reused_only_for_untouched_producer_output(ap):
    y = np.random.default_rng(0).standard_normal(4000).astype(np.float32)
    M = ap.melspectrogram(y, sr=16000, n_fft=400, hop_length=160, n_mels=40)
    ap._stub.calls.clear()
    ap.power_to_db(M, ref=torch.max)
    assert "mlxa_max_f32" not in ap._stub.calls          # peak came from the mel kernel's epilogue
    M.mul_(2.0)                                           # in-place edit bumps the version counter
    ap._stub.calls.clear()
    ap.power_to_db(M, ref=torch.max)
    assert "mlxa_max_f32" in ap._stub.calls               # stale peak must not be used
    ap._stub.calls.clear()
    ap.power_to_db(M.clone())
    assert "mlxa_max_f32" in ap._stub.calls


def test_error_paths_do_not_reach_the_library(ap):
    y = np.zeros(1000, np.float32)
    ap._stub.calls.clear()
    for fn, match in [(lambda: ap.stft(y, 256, 0), "hop_length must be positive"),
                      (lambda: ap.stft(y, 2048, 512, center=False), "must be >= frame_length"),
                      (lambda: ap.stft(y, 256, 64, pad_mode="wrap"), "Unknown pad_mode"),
                      (lambda: ap.melspectrogram(y, sr=16000, fmax=9000.0), "cannot exceed Nyquist"),
                      (lambda: ap.power_to_db(y, top_db=-1.0), "top_db must be positive"),
                      (lambda: ap.mfcc(y, n_mfcc=0), "n_mfcc must be positive"),
                      (lambda: ap.dct(y, type=1), "Only DCT type 2"),
                      (lambda: ap.griffinlim(np.ones((5, 3), np.float32), init="ones"), "Unknown init"),
                      (lambda: ap.frame(y, 64, 16, axis=0), "axis must be -1"),
                      (lambda: ap.istft(np.zeros(4, np.complex64)), "2D or 3D"),
                      (lambda: ap.stft(np.zeros(100, np.float32), 512, 128, pad_mode="reflect"), "reflect padding")]:
        with pytest.raises(ValueError, match=match):
            fn()
    assert ap._stub.calls == []
