// Epilogues of the fused forward kernel: what happens to spectrum bin X[b, t, k] the moment it
// exists in registers.  EP_STFT stores it; EP_MEL turns it into |X|^p inside a shared-memory
// [F][tile+1] tile that the band-sparse projection consumes after the tile's transforms are
// done (the spectrum never reaches HBM); EP_GL applies the Griffin-Lim projection + momentum.
#pragma once
#include "common.cuh"
#include "params.cuh"

namespace mlxa {

// padded-coordinate sample fetch: src = padded index - pad  (bit-exact index rules of the
// reference: pad_signal.metal:11-37 reflect, :53 edge (clamp), :92 constant)
MLXA_D float load_padded(const float* __restrict__ yb, int L, int src, int mode) {
    if ((unsigned)src < (unsigned)L) return __ldg(yb + src);
    if (mode == 0) return 0.f;
    if (mode == 1) src = (src < 0) ? -src : 2 * L - 2 - src;
    src = max(0, min(src, L - 1));
    return __ldg(yb + src);
}

MLXA_D float spectral_power(float2 X, int mode, float power) {
    const float sq = fmaf(X.x, X.x, X.y * X.y);
    if (mode == POW_SQUARE) return sq;
    const float a = sqrtf(sq);
    if (mode == POW_ABS) return a;
    return powf(a, power);
}

template <int EP>
MLXA_D void epilogue_bin(const FwdParams& p, int b, int t, int f_local, int k, float2 X,
                         float* s_ep, int ep_stride) {
    if constexpr (EP == EP_STFT) {
        p.spec[((long long)b * p.T + t) * p.F + k] = X;
    } else if constexpr (EP == EP_MEL) {
        s_ep[k * ep_stride + f_local] = spectral_power(X, p.power_mode, p.power);
    } else {
        const long long o = ((long long)b * p.T + t) * p.F + k;
        const float m = __ldg(p.mag + o);
        const float n2 = fmaf(X.x, X.x, X.y * X.y);
        float2 nw;
        if (n2 > 0.f) {
            const float sc = m * rsqrtf(n2);
            nw = make_float2(X.x * sc, X.y * sc);
        } else {
            nw = make_float2(m, 0.f);  // angle(0) = 0 -> mag * exp(0)
        }
        if (p.momentum > 0.f) {
            const float2 tp = p.tprev[o];
            p.rebuilt[o] = make_float2(fmaf(p.momentum, nw.x - tp.x, nw.x), fmaf(p.momentum, nw.y - tp.y, nw.y));
            p.tprev[o] = nw;
        } else {
            p.rebuilt[o] = nw;
        }
    }
}

// Band-sparse projection of the |X|^p tile: lanes run along the frames of the tile (coalesced
// (B, n_bands, T) stores), warps x sub-lanes run along the bands.  Each row of the filterbank
// is its contiguous support only (1.5-2.4 % of the dense matmul of mel.py:344).
template <int THREADS>
MLXA_D void mel_phase(const FwdParams& p, int b, int t0, int nt, const float* s_ep, int TT,
                      float* s_red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int subs = 32 / TT;
    const int t = lane % TT, sub = lane / TT;
    const int stride = TT + 1;
    float vmax = 0.f;
    for (int m = warp * subs + sub; m < p.n_bands; m += (THREADS / 32) * subs) {
        const int len = __ldg(p.band_len + m);
        const float* w = p.band_w + __ldg(p.band_off + m);
        const float* col = s_ep + __ldg(p.band_start + m) * stride + t;
        float a0 = 0.f, a1 = 0.f;
        int j = 0;
        for (; j + 1 < len; j += 2) {
            a0 = fmaf(__ldg(w + j), col[j * stride], a0);
            a1 = fmaf(__ldg(w + j + 1), col[(j + 1) * stride], a1);
        }
        if (j < len) a0 = fmaf(__ldg(w + j), col[j * stride], a0);
        float v = a0 + a1;
        if (t < nt) {
            vmax = fmaxf(vmax, v);
            if (p.db_mode) v = p.db_coef * log10f(fmaxf(v, p.db_amin) / fmaxf(p.db_ref, p.db_amin));
            p.mel[((long long)b * p.n_bands + m) * p.T + t0 + t] = v;
        }
    }
    if (p.gmax != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
        if (lane == 0) s_red[warp] = vmax;
        __syncthreads();
        if (threadIdx.x == 0) {
            float mx = 0.f;
            for (int i = 0; i < THREADS / 32; ++i) mx = fmaxf(mx, s_red[i]);
            atomicMax(reinterpret_cast<int*>(p.gmax), __float_as_int(mx));  // mel >= 0
        }
    }
}

}  // namespace mlxa
