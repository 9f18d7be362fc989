"""Griffin-Lim phase reconstruction (reference ``griffinlim.py``): the loop chains the fused ISTFT
kernel and the fused STFT+projection kernel on one stream, with no host round-trip inside."""
from __future__ import annotations

import numpy as np
import torch

from ._extension import _ext, check
from ._tensor import f32c, ptr, stream_ptr, to_tensor
from ._validation import validate_positive, validate_range
from .mel import frames_or_raise, pad_mode_code
from .stft import _istft_geometry, _istft_physical, istft, magnitude, phase, stft
from .windows import padded_window


def _to_physical_f32(x: torch.Tensor) -> torch.Tensor:
    """logical (B, F, T) float32 -> contiguous (B, T, F).  A spectrogram that already is a transposed view of the
    physical layout -- what magnitude(stft(y)) returns -- is used as it lies: no copy."""
    P = x.transpose(1, 2)
    if P.is_contiguous():
        return P
    x = x.contiguous()
    B, F, T = x.shape
    out = torch.empty((B, T, F), dtype=torch.float32, device=x.device)
    check(_ext.mlxa_transpose_f32(ptr(x), B, F, T, ptr(out), stream_ptr(x)), "transpose")
    return out


def _uniform_phase(rng: np.random.Generator, shape, device) -> torch.Tensor:
    """rng.uniform(-pi, pi, shape).astype(float32) in C order (reference griffinlim.py:112-115), drawn
    ON THE DEVICE: NumPy's PCG64 stream is reproduced bit for bit by a jump-ahead kernel, so seeds give
    the reference's phases without a host draw + upload of the whole spectrogram."""
    n = int(np.prod(shape))
    st = rng.bit_generator.state
    if n == 0 or st.get("bit_generator") != "PCG64":
        return torch.from_numpy(rng.uniform(-np.pi, np.pi, shape).astype(np.float32)).to(device)
    state, inc = int(st["state"]["state"]), int(st["state"]["inc"])
    out = torch.empty(shape, dtype=torch.float32, device=device)
    m64 = (1 << 64) - 1
    check(_ext.mlxa_pcg64_uniform_f32(state >> 64, state & m64, inc >> 64, inc & m64, -np.pi, np.pi, n, ptr(out),
                                      torch.cuda.current_stream(device).cuda_stream), "pcg64_uniform")
    rng.bit_generator.advance(n)  # keep a caller-supplied Generator in step with what was consumed
    return out


def griffinlim(S, n_iter: int = 32, hop_length: int | None = None, win_length: int | None = None,
               n_fft: int | None = None, window="hann", center: bool = True, length: int | None = None,
               pad_mode: str = "constant", momentum: float = 0.99, init: str = "random", random_state=None):
    """Reconstruct a signal from a magnitude spectrogram (reference griffinlim.py:17-196).

    The update is the reference's: new = S*exp(j*angle(stft(istft(rebuilt)))),
    rebuilt = new + momentum*(new - tprev), tprev = new.  The random initial phase is NumPy's
    default_rng(random_state).uniform(-pi, pi) stream in (B, F, T) order, like the reference, so seeds
    reproduce -- generated on the device by a bit-identical PCG64 kernel (no host draw, no upload)."""
    validate_positive(n_iter, "n_iter")
    validate_range(momentum, "momentum", min_val=0.0, max_val=1.0, max_inclusive=False)
    S = to_tensor(S, torch.float32)  # (made dense in the physical layout below: a (B, T, F)-backed view costs nothing)
    if S.ndim not in (2, 3):
        raise ValueError(f"S must be (freq_bins, n_frames) or (batch, freq_bins, n_frames), got shape {tuple(S.shape)}")
    batched = S.ndim == 3
    if not batched:
        S = S[None]
    B, F, T = S.shape
    if n_fft is None:
        n_fft = 2 * (F - 1)
    if hop_length is None:
        hop_length = n_fft // 4
    if win_length is None:
        win_length = n_fft
    mode = pad_mode_code(pad_mode)
    rng = np.random.default_rng(random_state)
    if init not in ("random", "zeros"):
        raise ValueError(f"Unknown init: '{init}'. Supported: 'random', 'zeros'")
    win = padded_window(window, win_length, n_fft)
    mag = _to_physical_f32(S)                 # (B, T, F)
    # One projected spectrum lives in HBM.  The momentum extrapolation rebuilt = new + m*(new - tprev) is taken
    # through the (linear) inverse transform: istft(rebuilt) = u + m*(u - u_prev) with u = istft(new), so each
    # iteration keeps the previous inverse (a signal) instead of the previous projection (a spectrum).
    cur = torch.empty((B, T, F, 2), dtype=torch.float32, device=S.device)
    st = rng.bit_generator.state
    if init == "random" and mag.numel() and st.get("bit_generator") == "PCG64":
        # S * exp(i * uniform(-pi, pi)) in one kernel: the phases of the reference's (B, F, T)-ordered draw land
        # at their physical (B, T, F) positions, the angle tensor is never written
        state, inc, m64 = int(st["state"]["state"]), int(st["state"]["inc"]), (1 << 64) - 1
        check(_ext.mlxa_pcg64_polar_f32(state >> 64, state & m64, inc >> 64, inc & m64, -np.pi, np.pi, ptr(mag), B, F, T,
                                        ptr(cur), stream_ptr(S)), "pcg64_polar")
        rng.bit_generator.advance(B * F * T)  # keep a caller-supplied Generator in step with what was consumed
    else:
        angles = (_uniform_phase(rng, (B, F, T), S.device) if init == "random"
                  else torch.zeros((B, F, T), dtype=torch.float32, device=S.device))
        ang = _to_physical_f32(angles)
        if mag.numel():
            check(_ext.mlxa_polar_f32(ptr(mag), ptr(ang), mag.numel(), ptr(cur), stream_ptr(S)), "polar")
        del ang, angles
    spec = torch.view_as_complex(cur)
    y = u = u_prev = None
    for it in range(n_iter + 1):  # n_iter projections, n_iter + 1 inverse transforms (reference griffinlim.py:129-183)
        if momentum > 0:
            if u is None:
                ola_len, trim, out_len = _istft_geometry(T, n_fft, hop_length, center, length)
                u, spare = (torch.empty((B, max(out_len, 0)), dtype=torch.float32, device=S.device) for _ in range(2))
            else:
                u_prev, u = u, (spare if u_prev is None else u_prev)  # tprev of the first update is the initial spectrum
            y = _istft_physical(spec, n_fft, hop_length, win, center, length, out=y, u_prev=u_prev if it > 0 else None,
                                momentum=momentum, u_out=u)
        else:
            y = _istft_physical(spec, n_fft, hop_length, win, center, length, out=y)
        if it == n_iter:
            break
        L = y.shape[1]
        T_new = frames_or_raise(L, n_fft, hop_length, center, pad_mode)
        check(_ext.mlxa_griffinlim_project_f32(ptr(y), B, L, y.stride(0), ptr(win), n_fft, hop_length, int(center),
                                               mode, T, min(T, T_new), ptr(mag), ptr(cur), stream_ptr(S)),
              "griffinlim")
    return y if batched else y[0]


def _polar(S: torch.Tensor, angles: torch.Tensor) -> torch.Tensor:
    """S * exp(j * angles), complex64, by the library's polar kernel (any common dense layout)."""
    S, angles = S.contiguous(), angles.contiguous()
    out = torch.empty(tuple(S.shape) + (2,), dtype=torch.float32, device=S.device)
    check(_ext.mlxa_polar_f32(ptr(S), ptr(angles), S.numel(), ptr(out), stream_ptr(S)), "polar")
    return torch.view_as_complex(out)


def griffinlim_iter(S, angles, hop_length: int, win_length: int, n_fft: int, window="hann", center: bool = True,
                    pad_mode: str = "constant", momentum: float = 0.99, tprev=None):
    """One iteration + reconstruction MSE (reference griffinlim.py:199-284), composed from the public kernels
    (polar, istft, stft, magnitude, phase, momentum step); meant for custom stopping rules, not for speed.  Only
    the scalar MSE diagnostic is a torch reduction."""
    S = f32c(S)
    angles = f32c(angles)
    rebuilt = _polar(S, angles)
    y = istft(rebuilt, hop_length=hop_length, win_length=win_length, n_fft=n_fft, window=window, center=center)
    new = stft(y, n_fft=n_fft, hop_length=hop_length, win_length=win_length, window=window, center=center,
               pad_mode=pad_mode)
    error = torch.mean((S - magnitude(new)) ** 2)
    new_angles = phase(new)
    new = _polar(S, new_angles)
    if momentum > 0 and tprev is not None:
        tp = torch.view_as_real(tprev.to(torch.complex64).contiguous())
        nw = torch.view_as_real(new)
        out = torch.empty_like(nw)
        check(_ext.mlxa_momentum_f32(ptr(nw), ptr(tp), float(momentum), nw.numel(), ptr(out), stream_ptr(nw)), "momentum")
        out = torch.view_as_complex(out)
    else:
        out = new
    return new_angles, out, error
