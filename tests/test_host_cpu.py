"""CPU-side tests: host constants of the product package against the reference fixtures, the
C-ABI surface (library loads, exports every symbol include/mlxa_cuda.h declares), the FFT engine
through its host emulation, and the sharding/all-reduce logic on a 2-rank gloo group."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_every_declared_symbol():
    from mlx_audio_primitives_b200 import _extension as ext
    hdr = open(os.path.join(ROOT, "include", "mlxa_cuda.h")).read()
    declared = set(re.findall(r"\b(mlxa_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = ctypes.CDLL(ext.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in mlxa_cuda.h but not exported"
    assert declared - {"mlxa_last_error", "mlxa_packed_bank_words"} == set(ext.SIGNATURES), "host-layer signature table out of sync"
    assert ext._ext.mlxa_plan_group(400) == -2 and ext._ext.mlxa_plan_group(2048) == -4 and ext._ext.mlxa_plan_group(777) == 32
    assert ext._ext.mlxa_abi_version() == ext.ABI_VERSION
    assert ext._ext.mlxa_has_fast_plan(400) == 1 and ext._ext.mlxa_has_fast_plan(2048) == 1
    assert ext._ext.mlxa_has_fast_plan(600) == 1 and ext._ext.mlxa_has_fast_plan(601) == 0 and ext._ext.mlxa_has_fast_plan(8192) == 1
    # argument validation happens before any CUDA call, so it is testable without a GPU
    rc = ext._ext.mlxa_pad_signal_f32(None, 1, 10, 3, 0, None, None)
    assert rc == -1 and b"null" in ext._ext.mlxa_last_error()
    with pytest.raises(ValueError):
        ext.check(rc, "pad_signal")


def test_host_constants_match_reference_fixtures(golden):
    from mlx_audio_primitives_b200.windows import window_host
    from mlx_audio_primitives_b200.mel import mel_filterbank_host, pack_bank_host, hz_to_mel, mel_to_hz
    from mlx_audio_primitives_b200.mfcc import dct_matrix_host
    from mlx_audio_primitives_b200.filterbanks import _linear_host
    for key in golden.files:
        parts = key.split("/")
        if parts[0] == "window":
            assert np.array_equal(window_host(parts[1], int(parts[2]), bool(int(parts[3]))), golden[key]), key
        elif parts[0] == "melfb":
            sr, n_fft, n_mels, fmin = int(parts[1]), int(parts[2]), int(parts[3]), float(parts[4])
            fmax = sr / 2.0 if parts[5] == "None" else float(parts[5])
            norm = None if parts[7] == "None" else parts[7]
            fb = mel_filterbank_host(sr, n_fft, n_mels, fmin, fmax, bool(int(parts[6])), norm)
            assert np.array_equal(fb, golden[key]), key
            # the packed band-sparse forms (what the kernels consume) reproduce the dense matrix exactly
            # row-pair format (group = -GP): per band 1 + nq entries of 4 words ({w0, w1, 0, 0}, then quads), nq shared
            # by GP adjacent bands, descriptors padded to a multiple of 32 bands, every run inside the F + 3 rows of
            # the power tile
            F = fb.shape[1]
            for GP in (1, 2, 4, 8, 20):
                packed, n_wt = pack_bank_host(fb, -GP)
                n_pad = -(-n_mels // 32) * 32
                start, nq, off, ln = packed[n_wt:].view(np.int32).reshape(n_pad, 4).T
                assert packed.size == n_wt + 4 * n_pad and n_wt == 4 * int((1 + nq).sum())
                dense = np.zeros((n_pad, F + 3), np.float32)
                for m in range(n_pad):
                    ent = packed[4 * off[m]:4 * (off[m] + 1 + nq[m])]
                    run = np.concatenate([ent[:2], ent[4:]])
                    assert not ent[2:4].any() and start[m] >= 0 and start[m] + run.size <= F + 3
                    dense[m, start[m]:start[m] + run.size] = run
                    assert (nq[m // GP * GP:(m // GP + 1) * GP] == nq[m]).all()  # warp-uniform trip count
                assert np.array_equal(dense[:n_mels, :F], fb) and not dense[:, F:].any() and not dense[n_mels:].any()
            for group in (16, 32):
                packed, n_wt = pack_bank_host(fb, group)
                n_groups = -(-n_mels // group)
                ints = packed[n_wt:].view(np.int32)
                start, ln = ints[:n_mels], ints[n_mels:2 * n_mels]
                goff, glen = ints[2 * n_mels:2 * n_mels + n_groups], ints[2 * n_mels + n_groups:2 * n_mels + 2 * n_groups]
                dense = np.zeros_like(fb)
                for m in range(n_mels):
                    j_, g_ = divmod(m, group)
                    col = packed[goff[j_] + g_: goff[j_] + glen[j_] * group: group]
                    assert ln[m] <= glen[j_] and not col[ln[m]:].any()
                    dense[m, start[m]:start[m] + ln[m]] = col[:ln[m]]
                assert np.array_equal(dense, fb)
        elif parts[0] == "dctmat":
            norm = None if parts[3] == "None" else parts[3]
            assert np.array_equal(dct_matrix_host(int(parts[1]), int(parts[2]), norm), golden[key]), key
    assert np.array_equal(_linear_host(22050, 1024, 32, 0.0, 11025.0, "slaney"), golden["linfb/22050/1024/32"])
    from mlx_audio_primitives_b200.filterbanks import _bark_host
    for key in [k for k in golden.files if k.startswith("barkfb/")]:  # reference filterbanks.py:159-231, bit for bit
        _, sr, n_fft, nb, fmin, fmax, formula, norm = key.split("/")
        fmax = int(sr) / 2.0 if fmax == "None" else float(fmax)
        fb = _bark_host(int(sr), int(n_fft), int(nb), float(fmin), fmax, formula, None if norm == "None" else norm)
        assert np.array_equal(fb, golden[key]), key
    assert np.array_equal(hz_to_mel(golden["hz"]), golden["hz_to_mel/slaney"])
    assert np.array_equal(hz_to_mel(golden["hz"], True), golden["hz_to_mel/htk"])
    assert np.array_equal(mel_to_hz(hz_to_mel(golden["hz"])), golden["mel_to_hz/slaney"])


def test_savgol_operators_match_scipy():
    """The host-side Savitzky-Golay taps / edge operators of delta() (NumPy least squares) against SciPy's own
    coefficients and against savgol_filter(mode='interp') on random data."""
    from scipy.signal import savgol_coeffs, savgol_filter
    from mlx_audio_primitives_b200.mfcc import savgol_operators_host
    rng = np.random.default_rng(0)
    for width, polyorder, deriv, spacing in [(9, 1, 1, 1.0), (9, 2, 2, 1.0), (5, 1, 1, 1.0), (3, 1, 1, 1.0), (7, 3, 2, 0.5), (11, 2, 1, 2.0)]:
        taps, left, right = savgol_operators_host(width, polyorder, deriv, spacing)
        ref = savgol_coeffs(width, polyorder, deriv=deriv, delta=spacing)[::-1]
        assert np.abs(taps - ref).max() <= 1e-6 * max(np.abs(ref).max(), 1e-30)
        x = rng.standard_normal((4, 40))
        want = savgol_filter(x, width, polyorder, deriv=deriv, delta=spacing, mode="interp")
        h = width // 2
        assert np.abs(x[:, :width] @ left.T.astype(np.float64) - want[:, :h]).max() <= 2e-5 * np.abs(want).max()
        assert np.abs(x[:, -width:] @ right.T.astype(np.float64) - want[:, -h:]).max() <= 2e-5 * np.abs(want).max()


def test_poly_filter_matches_scipy():
    """The host-side design of resample_poly (Kaiser low-pass, padding, trimming) against SciPy: applying the taps by the
    kernel's formula reproduces scipy.signal.resample_poly."""
    from scipy.signal import resample_poly
    from mlx_audio_primitives_b200.resample import poly_filter_host
    x = np.random.default_rng(1).standard_normal((2, 157)).astype(np.float32)
    for up, down in [(1, 2), (3, 1), (2, 3), (160, 147), (5, 4)]:
        taps, pre, n_out = poly_filter_host(up, down, x.shape[1])
        ref = resample_poly(x, up, down, axis=-1)
        assert ref.shape == (2, n_out)
        got = np.zeros((2, n_out))
        for j in range(n_out):
            m = (j + pre) * down
            lo = max(0, -((-(m - taps.size + 1)) // up))
            for i in range(lo, min(x.shape[1] - 1, m // up) + 1):
                got[:, j] += x[:, i].astype(np.float64) * taps[m - i * up]
        assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max()


def test_host_validation_messages():
    from mlx_audio_primitives_b200.mel import _resolve_stft_args, check_band_args, frames_or_raise, pad_mode_code
    from mlx_audio_primitives_b200.windows import window_host
    from mlx_audio_primitives_b200.stft import _istft_geometry
    with pytest.raises(ValueError, match="hop_length must be positive"):
        _resolve_stft_args(512, 0, None)
    with pytest.raises(ValueError, match=r"win_length \(1024\) must be <= n_fft \(512\)"):
        _resolve_stft_args(512, 128, 1024)
    with pytest.raises(ValueError, match="should typically be <= n_fft"):
        _resolve_stft_args(512, 1024, None)
    assert _resolve_stft_args(2048, None, None) == (512, 2048)
    with pytest.raises(ValueError, match="must be less than fmax"):
        check_band_args(40, "n_mels", 5000.0, 4000.0, 16000)
    with pytest.raises(ValueError, match="fmin must be non-negative"):
        check_band_args(40, "n_mels", -1.0, None, 16000)
    with pytest.raises(ValueError, match="Unknown pad_mode"):
        pad_mode_code("wrap")
    with pytest.raises(ValueError, match="Unknown window type"):
        window_host("kaiser", 64, True)
    with pytest.raises(ValueError, match="must be >= frame_length"):
        frames_or_raise(100, 512, 128, False, "constant")
    with pytest.raises(ValueError, match="reflect padding"):
        frames_or_raise(100, 512, 128, True, "reflect")
    assert frames_or_raise(22050, 2048, 512, True, "constant") == 44       # reference docstring stft.py:181
    assert frames_or_raise(480000, 400, 160, True, "constant") == 3001     # BASELINE config C2
    # ISTFT geometry (reference stft.py:300-338)
    assert _istft_geometry(44, 2048, 512, True, None) == (2048 + 43 * 512, 1024, 43 * 512)
    assert _istft_geometry(44, 2048, 512, True, 22050) == (22050 + 2048, 1024, 22050)
    assert _istft_geometry(44, 2048, 512, False, 5000) == (5000, 0, 5000)
    assert _istft_geometry(1, 2048, 512, True, None)[2] == 0


@pytest.fixture(scope="module")
def emul():
    so = os.path.join(ROOT, "tests", "emul", "libfft_emul.so")
    src = os.path.join(ROOT, "tests", "emul", "fft_emul.cpp")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-I/usr/local/cuda/include",
                        "-I" + os.path.join(ROOT, "mlx_audio_primitives_b200", "csrc", "cuda"), src, "-o", so],
                       check=True)
    return ctypes.CDLL(so)


def _c(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.mark.parametrize("R", [2, 3, 4, 5, 8, 9, 10, 16, 20, 25, 32, 64])
def test_in_register_dft(emul, R):
    rng = np.random.default_rng(R)
    x = (rng.standard_normal(R) + 1j * rng.standard_normal(R)).astype(np.complex64)
    out = np.zeros(R, np.complex64)
    assert emul.emul_radix(R, _c(x), _c(out)) == 0
    ref = np.fft.fft(x.astype(np.complex128))
    assert np.abs(out - ref).max() <= 5e-7 * np.abs(ref).max()


PLANNED = [32, 64, 128, 256, 400, 480, 512, 600, 800, 1000, 1024, 1200, 1600, 2000, 2048, 3072, 4096, 8192]


def test_planned_sizes_are_one_list():
    """csrc/cuda/fft_sizes.cuh is the one list of compiled plans: the build script, the library and the tests agree."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_b", os.path.join(ROOT, "mlx_audio_primitives_b200", "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    assert list(b.PLANNED_NFFT) == PLANNED
    from mlx_audio_primitives_b200 import _extension as ext
    for n in range(16, 8300):
        assert ext._ext.mlxa_has_fast_plan(n) == int(n in PLANNED), n


@pytest.mark.parametrize("n_fft", PLANNED)
def test_plan_fft_emulated(emul, n_fft):
    """The same __host__ __device__ pass functions the kernels run, lanes executed sequentially."""
    N = emul.emul_plan_length(n_fft)
    rng = np.random.default_rng(n_fft)
    x = (rng.standard_normal(N) + 1j * rng.standard_normal(N)).astype(np.complex64)
    out = np.zeros(N, np.complex64)
    assert emul.emul_plan_fft(n_fft, _c(x), _c(out)) == 0
    ref = np.fft.fft(x.astype(np.complex128))
    assert np.abs(out - ref).max() <= 5e-7 * np.abs(ref).max()


@pytest.mark.parametrize("rot", [0, 1])
def test_mirror_paired_powers_emulated(emul, rot):
    """fft_mirror.cuh (the n_fft = 400 mel kernel's transform): both members of every Hermitian pair come
    out of one lane's registers, so |Xa|^2, |Xb|^2 of a frame pair need no unpack; with the frames fed
    cyclically rotated (odd lane group) the powers are unchanged."""
    rng = np.random.default_rng(400 + rot)
    fa, fb = rng.standard_normal((2, 400)).astype(np.float32)
    win = (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(400) / 400)).astype(np.float32)
    pa, pb = np.full(201, np.nan, np.float32), np.full(201, np.nan, np.float32)
    assert emul.emul_mirror_powers_400(_c(fa), _c(fb), _c(win), rot, _c(pa), _c(pb)) == 0
    for got, fr in ((pa, fa), (pb, fb)):
        ref = 4.0 * np.abs(np.fft.rfft(fr.astype(np.float64) * win)) ** 2
        assert np.isfinite(got).all()  # every bin 0..200 was produced by some lane
        assert np.abs(got - ref).max() <= 2e-6 * ref.max()


@pytest.mark.parametrize("seed", [0, 1, 42, 2**40 + 7])
def test_pcg64_stream_matches_numpy(emul, seed):
    """Device RNG for the Griffin-Lim init (same __host__ __device__ code, run on the host): bit-identical
    to np.random.default_rng(seed).uniform(-pi, pi, n).astype(float32), with per-chunk jump-ahead."""
    n = 10007
    st = np.random.default_rng(seed).bit_generator.state["state"]
    m64 = (1 << 64) - 1
    out = np.zeros(n, np.float32)
    emul.emul_pcg64_uniform.argtypes = [ctypes.c_uint64] * 4 + [ctypes.c_double, ctypes.c_double, ctypes.c_longlong,
                                                                ctypes.c_int, ctypes.c_void_p]
    emul.emul_pcg64_uniform.restype = None
    emul.emul_pcg64_uniform(st["state"] >> 64, st["state"] & m64, st["inc"] >> 64, st["inc"] & m64, -np.pi, np.pi, n, 32, _c(out))
    want = np.random.default_rng(seed).uniform(-np.pi, np.pi, n).astype(np.float32)
    assert np.array_equal(out, want)


def test_shard_bounds():
    from mlx_audio_primitives_b200.distributed import shard_bounds
    for n, w in [(1024, 8), (64, 8), (10, 4), (3, 8), (1, 2)]:
        spans = [shard_bounds(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


_GLOO_WORKER = r"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
from mlx_audio_primitives_b200 import distributed as d
from oracle import spectral as o
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
# global batch of mel-like data, sharded by clips; the dB clamp must use the GLOBAL peak
full = np.random.default_rng(0).random((6, 8, 20)).astype(np.float32) * np.arange(1, 7, dtype=np.float32)[:, None, None]
lo, hi = d.shard_bounds(6, rank, 2)
local = full[lo:hi]
d.enable()
peak = torch.tensor([float(local.max())])
d.all_reduce_max_(peak)
assert float(peak) == float(full.max()), (float(peak), float(full.max()))
mine = o.power_to_db(local, ref=float(peak), top_db=None)
mine = np.maximum(mine, o.power_to_db(np.array([float(peak)], np.float32), ref=float(peak), top_db=None)[0] - 80.0)
want = o.power_to_db(full, ref=np.max, top_db=80.0)[lo:hi]
assert np.allclose(mine, want, atol=1e-5)
d.disable()
q = torch.tensor([float(rank)]); d.all_reduce_max_(q); assert float(q) == float(rank)  # disabled: untouched
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_global_peak_allreduce_gloo_world2(tmp_path):
    """N>1 host path on CPU: shard -> local peak -> all_reduce(MAX) -> dB equals the unsharded result."""
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    outs = [p.communicate(timeout=180)[0].decode() for p in procs]
    for p, out in zip(procs, outs):
        assert p.returncode == 0, out


def test_public_signatures_match_the_reference():
    """Every function the reference exports exists here with the same argument names, order and defaults
    (tests/golden/reference_signatures.json, parsed from the reference's sources by generate_reference_signatures.py)."""
    import ast
    import glob
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    want = json.load(open(os.path.join(root, "tests", "golden", "reference_signatures.json")))
    sys.path.insert(0, os.path.join(root, "tests", "golden"))
    from generate_reference_signatures import signatures
    mine = signatures(glob.glob(os.path.join(root, "mlx_audio_primitives_b200", "*.py")))
    exported = None
    for n in ast.walk(ast.parse(open(os.path.join(root, "mlx_audio_primitives_b200", "__init__.py")).read())):
        if isinstance(n, ast.Assign) and getattr(n.targets[0], "id", "") == "__all__":
            exported = {e.value for e in n.value.elts}
    assert len(want) >= 37
    norm = lambda v: None if v is None else v.replace('"', "'")
    for name, sig in want.items():
        assert name in exported and name in mine, name
        got = mine[name]
        assert [a for a, _ in got["args"]] == [a for a, _ in sig["args"]], name
        assert [norm(d) for _, d in got["args"]] == [norm(d) for _, d in sig["args"]], name
        assert [(a, norm(d)) for a, d in got["kwonly"]] == [(a, norm(d)) for a, d in sig["kwonly"]], name
