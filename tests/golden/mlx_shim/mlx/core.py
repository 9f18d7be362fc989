"""Minimal NumPy-backed `mlx.core` lookalike (see package docstring)."""
from __future__ import annotations

import numpy as np

float32 = np.float32
float64 = np.float64
complex64 = np.complex64
int32 = np.int32
int64 = np.int64
bool_ = np.bool_
pi = np.pi


class array(np.ndarray):
    """ndarray subclass so `isinstance(x, mx.array)` and `mx.array(data, dtype=)` work."""

    def __new__(cls, data, dtype=None):
        a = np.asarray(data)
        if dtype is None:
            if a.dtype == np.float64:
                dtype = np.float32  # MLX defaults python/NumPy doubles to float32
            elif a.dtype == np.complex128:
                dtype = np.complex64
            elif a.dtype == np.int64:
                dtype = np.int32
        if dtype is not None:
            a = a.astype(dtype)
        return np.array(a, copy=True).view(cls)

    def astype(self, dtype, *a, **k):  # keep the subclass
        return np.ndarray.astype(self, dtype, *a, **k).view(array)

    def item(self):
        return np.ndarray.item(np.asarray(self))


def _w(x):
    if isinstance(x, np.ndarray):
        return x.view(array)
    return array(x)


def eval(*a, **k):
    return None


def synchronize(*a, **k):
    return None


def compile(fn, *a, **k):
    return fn


def zeros(shape, dtype=float32):
    return _w(np.zeros(shape, dtype=dtype))


def ones(shape, dtype=float32):
    return _w(np.ones(shape, dtype=dtype))


def arange(*a, dtype=None):
    r = np.arange(*a)
    if dtype is not None:
        r = r.astype(dtype)
    elif r.dtype == np.int64:
        r = r.astype(np.int32)
    elif r.dtype == np.float64:
        r = r.astype(np.float32)
    return _w(r)


def linspace(a, b, n, dtype=float32):
    return _w(np.linspace(a, b, n).astype(dtype))


def pad(x, pad_width, mode="constant", constant_values=0):
    if mode == "constant":
        return _w(np.pad(np.asarray(x), pad_width, mode="constant", constant_values=constant_values))
    return _w(np.pad(np.asarray(x), pad_width, mode=mode))


def concatenate(xs, axis=0):
    return _w(np.concatenate([np.asarray(v) for v in xs], axis=axis))


def transpose(x, axes=None):
    return _w(np.transpose(np.asarray(x), axes))


def moveaxis(x, s, d):
    return _w(np.moveaxis(np.asarray(x), s, d))


def broadcast_to(x, shape):
    return _w(np.broadcast_to(np.asarray(x), shape))


def take(x, idx, axis=None):
    return _w(np.take(np.asarray(x), np.asarray(idx), axis=axis))


def as_strided(x, shape, strides, offset=0):
    a = np.ascontiguousarray(np.asarray(x)).reshape(-1)[offset:]
    it = a.dtype.itemsize
    v = np.lib.stride_tricks.as_strided(a, shape=shape, strides=[s * it for s in strides])
    return _w(np.array(v))


def _f32(fn):
    def g(*a, **k):
        r = fn(*[np.asarray(v) for v in a], **k)
        if isinstance(r, np.ndarray):
            if r.dtype == np.float64:
                r = r.astype(np.float32)
            elif r.dtype == np.complex128:
                r = r.astype(np.complex64)
            return _w(r)
        if isinstance(r, np.floating):
            return _w(np.float32(r))
        return _w(r)
    return g


abs = _f32(np.abs)
power = _f32(np.power)
maximum = _f32(np.maximum)
minimum = _f32(np.minimum)
log10 = _f32(np.log10)
log = _f32(np.log)
exp = _f32(np.exp)
sqrt = _f32(np.sqrt)
sin = _f32(np.sin)
cos = _f32(np.cos)
arctan2 = _f32(np.arctan2)
matmul = _f32(np.matmul)
cumsum = _f32(np.cumsum)


def _red(fn):
    def g(x, axis=None, keepdims=False):
        return _w(np.asarray(fn(np.asarray(x), axis=axis, keepdims=keepdims)))
    return g


max = _red(np.max)
min = _red(np.min)
sum = _red(np.sum)
mean = _red(np.mean)


class _FFT:
    @staticmethod
    def rfft(x, n=None, axis=-1):
        x = np.asarray(x)
        import scipy.fft as sf
        return _w(sf.rfft(x, n=n, axis=axis).astype(np.complex64 if x.dtype == np.float32 else np.complex128))

    @staticmethod
    def irfft(x, n=None, axis=-1):
        x = np.asarray(x)
        import scipy.fft as sf
        return _w(sf.irfft(x, n=n, axis=axis).astype(np.float32 if x.dtype == np.complex64 else np.float64))


fft = _FFT()


def _fused_overlap_add(inputs, output_shapes, **kw):
    """Stand-in for the reference's inline Metal kernel source (stft.py:548-596),
    executed thread-by-thread semantics in float32 (ascending-frame accumulation)."""
    frames, window = (np.asarray(v, dtype=np.float32) for v in inputs)
    t = dict(kw["template"])
    hop, out_len, B, T, n_fft = (t[k] for k in ("hop_length", "output_length", "batch_size", "n_frames", "n_fft"))
    i = np.arange(out_len)
    first = np.maximum(-(-(i - n_fft + 1) // hop), 0)
    last = np.minimum(i // hop, T - 1)
    s = np.zeros((B, out_len), np.float32)
    ws = np.zeros(out_len, np.float32)
    span = int((last - first).max()) + 1 if out_len else 0
    for d in range(span):
        f = first + d
        ok = f <= last
        fi = np.where(ok, f, 0)
        k = np.where(ok, i - fi * hop, 0)
        wv = np.where(ok, window[k], np.float32(0))
        s += (wv[None, :] * frames[:, fi, k]).astype(np.float32)
        ws += (wv * wv).astype(np.float32)
    return [_w((s / np.maximum(ws, np.float32(1e-8))).astype(np.float32))]


class _Fast:
    @staticmethod
    def metal_kernel(name, input_names, output_names, source, **kw):
        if name != "fused_overlap_add":
            raise NotImplementedError(name)
        def run(inputs, output_shapes, output_dtypes=None, grid=None, threadgroup=None, template=None, init_value=0):
            return _fused_overlap_add(inputs, output_shapes, template=template)
        return run


fast = _Fast()
