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

template <int PW>
MLXA_D float spectral_power(float2 X, float power) {
    const float sq = fmaf(X.x, X.x, X.y * X.y);
    if constexpr (PW == POW_SQUARE) return sq;
    const float a = sqrtf(sq);
    if constexpr (PW == POW_ABS) return a;
    return powf(a, power);
}

// EP_STFT / EP_GL: one bin straight to global memory
template <int EP>
MLXA_D void epilogue_bin_global(const FwdParams& p, long long o, float2 X) {
    if constexpr (EP == EP_STFT) {
        p.spec[o] = X;
    } else {
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

// ---- band-sparse filterbank, packed (include/mlxa_cuda.h "packed filterbank") -----------------
// words: [w: 4*n_w4 floats][start: n_bands][n4: n_bands][off4: n_bands], padded to a multiple of 4.
// Row m covers bins [start[m], start[m] + 4*n4[m]) with weights w[4*off4[m] ...] zero-padded to a
// multiple of four so the projection loop reads them as float4.  The blob is bulk-copied to smem.
struct MelSmem {
    const float4* w4;
    const int* start;
    const int* n4;
    const int* off4;
};
MLXA_D MelSmem mel_smem_carve(const float* base, int n_bands, long long n_w4) {
    MelSmem m;
    m.w4 = reinterpret_cast<const float4*>(base);
    const int* ip = reinterpret_cast<const int*>(base + 4 * n_w4);
    m.start = ip;
    m.n4 = ip + n_bands;
    m.off4 = ip + 2 * n_bands;
    return m;
}

// Band-sparse projection of the |X|^p tile: lanes run along the frames of the tile (coalesced
// (B, n_bands, T) stores), warps x sub-lanes run along the bands.  Each row of the filterbank
// is its contiguous support only (1.5-2.4 % of the dense matmul of mel.py:344).
template <int THREADS>
MLXA_D void mel_phase(const FwdParams& p, int b, int t0, int nt, const float* s_ep, int TT, const MelSmem ms,
                      float* s_red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int subs = 32 / TT;
    const int t = lane % TT, sub = lane / TT;
    const int stride = TT + 1;
    float vmax = 0.f;
    const float db_ref = fmaxf(p.db_ref, p.db_amin);
    float* outb = p.mel + (long long)b * p.n_bands * p.T + t0 + t;
    for (int m = warp * subs + sub; m < p.n_bands; m += (THREADS / 32) * subs) {
        const int n4 = ms.n4[m];
        const float4* w4 = ms.w4 + ms.off4[m];
        const float* col = s_ep + ms.start[m] * stride + t;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 1
        for (int j = 0; j < n4; ++j) {
            const float4 w = w4[j];
            a0 = fmaf(w.x, col[0], a0);
            a1 = fmaf(w.y, col[stride], a1);
            a2 = fmaf(w.z, col[2 * stride], a2);
            a3 = fmaf(w.w, col[3 * stride], a3);
            col += 4 * stride;
        }
        float v = (a0 + a1) + (a2 + a3);
        if (t < nt) {
            vmax = fmaxf(vmax, v);
            if (p.db_mode) v = p.db_coef * log10f(fmaxf(v, p.db_amin) / db_ref);
            outb[(long long)m * p.T] = v;
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
