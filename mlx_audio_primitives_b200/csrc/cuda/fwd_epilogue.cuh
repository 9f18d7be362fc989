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

// coef * log10(max(x, amin) / refc) (convert.py:48-52) as  c1 * log2(max(x, amin)) + c0  with
// c1 = coef * log10(2), c0 = -c1 * log2(refc): one clamp, one MUFU.LG2 (absolute error 2^-22, i.e. < 1e-5 dB,
// far inside the 1e-3 dB parity bound) and one FFMA per value.  Every kernel that converts to dB uses these
// two functions, so fused and two-pass paths give the same bits (refc = 1 makes c0 an exact zero).
struct DbConst {
    float amin, c1, c0;
};
MLXA_D DbConst db_constants(float coef, float amin, float ref) {
    DbConst c;
    c.amin = amin;
    c.c1 = coef * 0.30102999566398120f;
    c.c0 = -c.c1 * log2f(fmaxf(ref, amin));
    return c;
}
MLXA_D float to_db_one(float x, float coef, float amin, float refc) {
    const DbConst c = db_constants(coef, amin, refc);
    return fmaf(__log2f(fmaxf(x, amin)), c.c1, c.c0);
}

// EP_STFT / EP_GL: one bin straight to global memory
// (m: the target magnitude of the Griffin-Lim projection, loaded by the caller ahead of the bin loop)
template <int EP>
MLXA_D void epilogue_bin_global(const FwdParams& p, long long o, float2 X, float m) {
    if constexpr (EP == EP_STFT) {
        p.spec[o] = X;
    } else {
        // mag * X / |X| with the bare SFU reciprocal square root (rsqrtf() wraps it in a denormal rescue: two
        // multiplies and a compare per bin).  |X|^2 below the smallest normal (|X| < 1.1e-19) counts as zero.
        const float n2 = fmaf(X.x, X.x, X.y * X.y);
        float r;
        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(n2));
        const float sc = m * r;
        float2 nw = pmul(X, sc, sc);
        if (!(n2 >= 1.17549435e-38f)) nw = make_float2(m, 0.f);  // angle(0) = 0 -> mag * exp(0)
        p.rebuilt[o] = nw;  // (the momentum step happens in the signal domain, inside the next inverse transform)
    }
}
template <int EP>
MLXA_D void epilogue_bin_global(const FwdParams& p, long long o, float2 X) {
    epilogue_bin_global<EP>(p, o, X, EP == EP_GL ? __ldg(p.mag + o) : 0.f);
}

// ---- band-sparse filterbank, packed for a lane group of G (include/mlxa_cuda.h) ---------------
// words: [wt: n_wt floats][start: n_bands][len: n_bands][goff: n_groups][glen: n_groups], padded to
// a multiple of 4, n_groups = ceil(n_bands / G).  Bands are taken G at a time (band j*G + g belongs
// to lane g); inside group j the weights are TRANSPOSED, wt[goff[j] + i*G + g] = weight of band
// j*G + g at bin start + i, zero-padded to glen[j] = the longest support in the group, so that the
// G lanes read consecutive words (no bank conflicts) and share one loop bound.
struct MelSmem {
    const float* wt;
    const int* start;
    const int* goff;
    const int* glen;
    int n_groups;
};
template <int G>
MLXA_D MelSmem mel_smem_carve(const float* base, int n_bands, long long n_wt) {
    MelSmem m;
    m.n_groups = (n_bands + G - 1) / G;
    m.wt = base;
    const int* ip = reinterpret_cast<const int*>(base + n_wt);
    m.start = ip;
    m.goff = ip + 2 * n_bands;
    m.glen = m.goff + m.n_groups;
    return m;
}

// Band-sparse projection done by the SAME lane group that produced the spectrum, straight from
// the |X|^p values it just parked in its exchange buffer: lanes run along the bands (row m of
// the filterbank is its contiguous support only -- 1.5-2.4 % of the dense matmul of mel.py:344),
// NF frames (1 for a packed transform, 2 for a frame pair) share every weight load.  Results go
// to the [n_bands][tile+1] staging tile s_out, column f0 (+1).  Reads past a band's own support
// hit zero weights (the values there are finite: stale transform output).
template <int G, int NF>
MLXA_D void mel_project_group(const MelSmem ms, int n_bands, int g, const float* pw, float* s_out, int ostride, int f0) {
    for (int j = 0; j < ms.n_groups; ++j) {
        const int m = j * G + g;
        const int n = ms.glen[j];
        const float* w = ms.wt + ms.goff[j] + g;
        const int st = (m < n_bands) ? ms.start[m] : 0;
        if constexpr (NF == 2) {
            const float2* p2 = reinterpret_cast<const float2*>(pw) + st;
            float a = 0.f, b = 0.f;
#pragma unroll 2
            for (int i = 0; i < n; ++i) {
                const float2 pp = p2[i];
                const float ww = w[i * G];
                a = fmaf(ww, pp.x, a);
                b = fmaf(ww, pp.y, b);
            }
            if (m < n_bands) {
                s_out[m * ostride + f0] = a;
                s_out[m * ostride + f0 + 1] = b;
            }
        } else {
            const float* p1 = pw + st;
            float a = 0.f, b = 0.f;
            int i = 0;
            for (; i + 1 < n; i += 2) {
                a = fmaf(w[i * G], p1[i], a);
                b = fmaf(w[(i + 1) * G], p1[i + 1], b);
            }
            if (i < n) a = fmaf(w[i * G], p1[i], a);
            if (m < n_bands) s_out[m * ostride + f0] = a + b;
        }
    }
}

// min of the raw values of a tile into its 64-frame block's slot (tiles never straddle blocks: the tile
// size is a power of two <= 64).  Values are >= 0, so the int ordering of the bit patterns is the float
// ordering; one atomic per warp.
constexpr int kMinBlockFrames = 64;
MLXA_D void block_min_to_global(const FwdParams& p, int b, int t0, float vmin) {
    // one REDUX instead of five shuffle + min rounds (the bit patterns of values >= 0 order like integers)
    const int m = __reduce_min_sync(0xffffffffu, __float_as_int(vmin));
    if ((threadIdx.x & 31) == 0)
        atomicMin(reinterpret_cast<int*>(p.block_min) + (long long)b * p.blocks_per_clip + t0 / kMinBlockFrames, m);
}

// Tile store: s_out [n_bands][TT+1] -> mel (B, n_bands, T), lanes along the frames (coalesced),
// with the optional fused dB.  TT is a power of two.  Returns the thread's running max of the raw
// values it stored (for power_to_db(ref=max / top_db)); the caller reduces it once per CTA.
template <int THREADS>
MLXA_D float mel_store_tile(const FwdParams& p, int b, int t0, int nt, const float* s_out, int TT, int log2TT, float vmax) {
    const int ostride = TT + 1;
    const float db_ref = fmaxf(p.db_ref, p.db_amin);
    float* outb = p.mel + (long long)b * p.n_bands * p.T + t0;
    const int n = p.n_bands << log2TT;
    float vmin = INFINITY;
    for (int idx = threadIdx.x; idx < n; idx += THREADS) {
        const int m = idx >> log2TT, t = idx & (TT - 1);
        if (t < nt) {
            float v = s_out[m * ostride + t];
            vmax = fmaxf(vmax, v);
            vmin = fminf(vmin, v);
            if (p.db_mode) v = to_db_one(v, p.db_coef, p.db_amin, db_ref);
            outb[(long long)m * p.T + t] = v;
        }
    }
    if (p.block_min != nullptr) block_min_to_global(p, b, t0, vmin);
    return vmax;
}

// ---- EP_FEAT: per-frame spectral statistics by the lane group that produced the spectrum ------------------
// The lane owns bins k = g + q*G (q < NQ) of one frame in sv[q*STRIDE] (0 beyond the last bin): |X| -- or
// |X|^power for flatness.  Reductions are shuffles inside the group (gmask); every lane returns the result.
// Same formulas and guards as feat_kernels.cu / reference features.py:120-442.
template <int KIND, int G, int NQ, int STRIDE>
MLXA_D float group_spectral_stat(const FwdParams& p, const float* sv, int g, unsigned gmask, int n_bins, long long frame) {
    auto gsum = [&](float v) {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o, G);
        return v;
    };
    const float* fq = p.feat_freq + g;
    const float step = p.feat_freq_step;
    auto freq_at = [&](int q_times_g) { return step > 0.f ? float(g + q_times_g) * step : __ldg(fq + q_times_g); };
    if constexpr (KIND == STAT_FLATNESS) {
        float sl = 0.f, sa = 0.f;
        static_for<NQ>([&](auto q) {
            constexpr int Q = decltype(q)::value;
            if (g + Q * G < n_bins) {
                const float v = fmaxf(sv[Q * STRIDE], p.feat_p2);
                sl += logf(v);
                sa += v;
            }
        });
        sl = gsum(sl);
        sa = gsum(sa);
        return expf(sl / float(n_bins)) / (sa / float(n_bins) + 1e-10f);
    }
    if constexpr (KIND == STAT_ROLLOFF) {
        auto gscan = [&](float v) {
#pragma unroll
            for (int o = 1; o < G; o <<= 1) {
                const float t = __shfl_up_sync(gmask, v, o, G);
                if (g >= o) v += t;
            }
            return v;
        };
        float total = 0.f;
        static_for<NQ>([&](auto q) { total += __shfl_sync(gmask, gscan(sv[decltype(q)::value * STRIDE]), G - 1, G); });
        const float thr = p.feat_p1 * total;  // the same blocked scan below reaches exactly `total`
        const int base = (threadIdx.x & 31) & ~(G - 1);
        float carry = 0.f;
        int idx = -1;
        static_for<NQ>([&](auto q) {
            constexpr int Q = decltype(q)::value;
            const float cs = carry + gscan(sv[Q * STRIDE]);
            const unsigned hit = (__ballot_sync(gmask, g + Q * G < n_bins && !(cs < thr)) >> base) & (G == 32 ? 0xffffffffu : ((1u << (G & 31)) - 1u));
            if (idx < 0 && hit) idx = Q * G + __ffs(hit) - 1;
            carry = __shfl_sync(gmask, cs, G - 1, G);
        });
        return __ldg(p.feat_freq + (idx < 0 ? n_bins - 1 : idx));
    }
    if constexpr (KIND == STAT_CENTROID || KIND == STAT_BANDWIDTH) {
        float s0 = 0.f, s1 = 0.f;
        static_for<NQ>([&](auto q) {
            constexpr int Q = decltype(q)::value;
            if (g + Q * G < n_bins) {
                s0 += sv[Q * STRIDE];
                s1 = fmaf(freq_at(Q * G), sv[Q * STRIDE], s1);
            }
        });
        s0 = gsum(s0);
        s1 = gsum(s1);
        const float c = s1 / (s0 + 1e-10f);
        if constexpr (KIND == STAT_CENTROID) {
            return c;
        } else {
            const float cc = p.feat_centroid ? __ldg(p.feat_centroid + frame) : c;
            const bool square = p.feat_p1 == 2.0f;
            float s2 = 0.f;
            static_for<NQ>([&](auto q) {
                constexpr int Q = decltype(q)::value;
                if (g + Q * G < n_bins) {
                    const float d = fabsf(freq_at(Q * G) - cc);
                    s2 = fmaf(sv[Q * STRIDE], square ? d * d : powf(d, p.feat_p1), s2);
                }
            });
            s2 = gsum(s2);
            const float qv = p.feat_norm ? s2 / (s0 + 1e-10f) : s2;
            return square ? sqrtf(qv) : powf(qv, 1.0f / p.feat_p1);
        }
    }
    return 0.f;
}

// ---- peak exchange over peer memory (params.cuh: PeakExchange) ------------------------------------
// The (epoch, peak) word is self-contained -- nothing else has to become visible with it -- so relaxed
// system-scope accesses suffice (a release store made the producer wait ~5 us for its own output writes).
MLXA_D void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
MLXA_D unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// one thread of the last CTA of a producer kernel: this rank's final peak to every rank's slot set
MLXA_D void peak_publish(const PeakExchange& x, float* gmax, unsigned n_ctas) {
    __threadfence();
    if (atomicAdd(x.ticket, 1u) != n_ctas - 1) return;
    *x.ticket = 0u;
    __threadfence();
    const float peak = __int_as_float(atomicMax(reinterpret_cast<int*>(gmax), 0));  // coherent read (values >= 0)
    const unsigned long long v = ((unsigned long long)x.epoch << 32) | (unsigned)__float_as_int(peak);
    for (int r = 0; r < x.world; ++r) st_relaxed_sys_u64(x.peer_slots[r] + (x.epoch & 1u) * x.world + x.rank, v);
}
// any thread of a consumer kernel: the maximum over all ranks.  Waits like a collective would: a slow rank (a
// stalled data loader, a first-call module load) only delays the others.  The poll backs off to ~1 us; a peer
// that never publishes (a dead rank) ends the wait after ~2 minutes with a trap, as a hung NCCL kernel would be
// ended by its watchdog, instead of leaving the GPU spinning for ever.
MLXA_D float peak_collect(const PeakExchange& x, const unsigned long long* my_slots) {
    float m = 0.f;
    for (int r = 0; r < x.world; ++r) {
        const unsigned long long* s = my_slots + (x.epoch & 1u) * x.world + r;
        unsigned long long v = ld_relaxed_sys_u64(s);
        unsigned ns = 32;
        for (unsigned spins = 0; (unsigned)(v >> 32) != x.epoch; ++spins) {
            if (spins > (1u << 27)) __trap();
            __nanosleep(ns);
            if (ns < 1024) ns <<= 1;
            v = ld_relaxed_sys_u64(s);
        }
        m = fmaxf(m, __int_as_float((int)(unsigned)v));
    }
    return m;
}

// all threads of a consumer CTA: the batch-global peak -- the local *gmax, or with an exchange the maximum
// over the ranks (thread 0 collects, one barrier)
MLXA_D float resolve_peak(const float* gmax, const PeakExchange& x, float* s_slot) {
    if (x.peer_slots == nullptr) return gmax ? __ldg(gmax) : 0.f;
    if (threadIdx.x == 0) *s_slot = peak_collect(x, x.peer_slots[x.rank]);
    __syncthreads();
    return *s_slot;
}

// one atomicMax per CTA (mel >= 0, so the int ordering of the bit patterns is the float ordering); with a
// peak exchange the last CTA to get here also publishes the final value to the peers
template <int THREADS>
MLXA_D void block_max_to_global(float vmax, float* gmax, float* s_red, const PeakExchange& xchg) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = vmax;
    __syncthreads();
    if (threadIdx.x == 0) {
        float mx = 0.f;
        for (int i = 0; i < THREADS / 32; ++i) mx = fmaxf(mx, s_red[i]);
        atomicMax(reinterpret_cast<int*>(gmax), __float_as_int(mx));
        if (xchg.peer_slots != nullptr) peak_publish(xchg, gmax, gridDim.x * gridDim.y);
    }
}

}  // namespace mlxa
