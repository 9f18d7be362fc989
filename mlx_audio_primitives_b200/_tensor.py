"""PyTorch / DLPack plumbing: device tensors in, raw pointers to the C ABI, current stream."""
from __future__ import annotations

import functools

import numpy as np
import torch


def require_cuda() -> None:
    if not torch.cuda.is_available():
        raise RuntimeError("mlx_audio_primitives_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def to_tensor(x, dtype=None, device=None) -> torch.Tensor:
    """Accept a torch tensor, any ``__dlpack__`` producer or a NumPy array; return a CUDA tensor."""
    require_cuda()
    if isinstance(x, torch.Tensor):
        t = x
    elif isinstance(x, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(x))
    elif hasattr(x, "__dlpack__"):
        t = torch.from_dlpack(x)
    else:
        t = torch.as_tensor(x)
    if not t.is_cuda:
        t = t.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t


def on_input_device(fn):
    """Run a public entry point with the CUDA device of its first CUDA-tensor argument current.  The C ABI launches
    on the current device and the constant caches (windows, filterbanks, twiddles) are keyed on it, so an input on
    cuda:1 while cuda:0 is current would otherwise mix a device-1 stream with device-0 launches and constants
    (PyTorch ops handle that case transparently; so do these).  Host inputs go to the current device."""
    @functools.wraps(fn)
    def scoped(*args, **kwargs):
        for a in (*args, *kwargs.values()):
            if isinstance(a, torch.Tensor) and a.device.type == "cuda":
                if a.device.index != torch.cuda.current_device():
                    with torch.cuda.device(a.device):
                        return fn(*args, **kwargs)
                break
        return fn(*args, **kwargs)
    return scoped


def publish(t):
    """Call before a freshly built device constant goes into a cache: the stream that filled it is synchronised, so a
    later hit from ANOTHER stream (or thread) never reads a table whose fill is still in flight.  Once per entry."""
    if torch.cuda.is_available():
        torch.cuda.current_stream().synchronize()
    return t


def f32c(x) -> torch.Tensor:
    """float32, contiguous, on the GPU."""
    return to_tensor(x, torch.float32).contiguous()


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def stream_ptr(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _is_dense_permutation(x: torch.Tensor) -> bool:
    """True when x covers one contiguous block exactly once (any permutation of a contiguous tensor)."""
    dims = sorted((st, sz) for st, sz in zip(x.stride(), x.shape) if sz != 1)
    expect = 1
    for st, sz in dims:
        if st != expect:
            return False
        expect *= sz
    return True


def dense_like(x: torch.Tensor, dtype=None) -> tuple[torch.Tensor, torch.Tensor]:
    """(x_dense, out): a densely laid-out version of x (any permutation of a contiguous block is
    kept as is) and an empty output with identical strides, so flat elementwise kernels apply."""
    if not _is_dense_permutation(x):
        x = x.contiguous()
    out = torch.empty_like(x, dtype=dtype if dtype is not None else x.dtype, memory_format=torch.preserve_format)
    if out.stride() != x.stride():  # pragma: no cover - defensive
        x = x.contiguous()
        out = torch.empty_like(x, dtype=dtype if dtype is not None else x.dtype)
    return x, out
