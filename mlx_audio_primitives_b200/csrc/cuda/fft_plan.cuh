// Multi-pass Stockham FFT executed by a group of G lanes (G | 32) of one warp.
//
// A plan factors the complex length N = R0*R1(*R2).  In pass p every butterfly b in
// [0, N/Rp) reads elements  b + r*(N/Rp)  (r < Rp), multiplies leg r by
// exp(-2*pi*i*(b mod Ns)*r/(Ns*Rp))  (Ns = product of earlier radices), runs an in-register
// DFT of length Rp and writes  (b/Ns)*Ns*Rp + (b mod Ns) + k*Ns.  After the last pass the
// data is in natural order.  Lanes own butterflies b = g, g+G, ...; between passes the data
// crosses lanes through a padded shared-memory buffer (phys(i) = i + PAD*(i/R0), which keeps
// every access pattern affine in the leg index and bank-conflict free for 64-bit accesses).
//
// All per-lane steps are __host__ __device__ so tests/emul can run the same code on the CPU.
#pragma once
#include "fft_radix.cuh"

namespace mlxa {

template <int N_, int G_, int R0_, int R1_, int R2_ = 1>
struct FftPlan {
    static constexpr int N = N_, G = G_, R0 = R0_, R1 = R1_, R2 = R2_;
    static constexpr int NPASS = (R2_ > 1) ? 3 : 2;
    static_assert(R0_ * R1_ * R2_ == N_, "radices must multiply to N");
    static_assert(32 % G_ == 0 || G_ == 64, "a group is a divisor of a warp, or two warps (exchange through a named barrier)");
    static constexpr int radix(int p) { return p == 0 ? R0 : (p == 1 ? R1 : R2); }
    static constexpr int ns(int p) { return p == 0 ? 1 : (p == 1 ? R0 : R0 * R1); }
    static constexpr int nb(int p) { return N / radix(p); }
    static constexpr int rounds(int p) { return ceil_div(nb(p), G); }
    static constexpr int regs(int p) { return rounds(p) * radix(p); }
    static constexpr int cmax(int a, int b) { return a > b ? a : b; }
    static constexpr int E = cmax(regs(0), cmax(regs(1), NPASS == 3 ? regs(2) : 0));
    static constexpr int PAD = (R0 & 1) ? 2 : 1;  // (R0 + PAD) odd -> conflict-free strides
    static constexpr int BUF = (N + PAD * (N / R0) + 1) & ~1;  // float2 slots per transform (even: 16-byte multiples)
    static MLXA_HD int phys(int i) { return i + PAD * (i / R0); }
    // inter-pass twiddles: pass p >= 1 uses tw[tw_off(p) + (r-1)*ns(p) + (b mod ns(p))]
    static constexpr int tw_off(int p) { return p <= 1 ? 0 : (R1 - 1) * R0; }
    static constexpr int TW = (R1 - 1) * R0 + (NPASS == 3 ? (R2 - 1) * R0 * R1 : 0);
};

// ---- pass steps ---------------------------------------------------------------------------
// load(idx) -> float2 supplies natural-order element idx (used for the first pass)
template <class P, int PASS, class LoadF>
MLXA_HD void pass_load_fn(int g, float2* v, LoadF&& load) {
    constexpr int R = P::radix(PASS), NB = P::nb(PASS), RD = P::rounds(PASS);
    static_for<RD>([&](auto rd) {
        const int b = g + decltype(rd)::value * P::G;
        if ((NB % P::G == 0) || b < NB) {
            static_for<R>([&](auto r) {
                v[decltype(rd)::value * R + decltype(r)::value] = load(b + decltype(r)::value * NB);
            });
        }
    });
}

// passes >= 1 read the padded exchange buffer
template <class P, int PASS>
MLXA_HD void pass_load_buf(int g, float2* v, const float2* buf) {
    constexpr int R = P::radix(PASS), NB = P::nb(PASS), RD = P::rounds(PASS);
    static_assert(PASS >= 1 && NB % P::R0 == 0, "affine read needs R0 | NB");
    constexpr int RS = NB + P::PAD * (NB / P::R0);
    static_for<RD>([&](auto rd) {
        const int b = g + decltype(rd)::value * P::G;
        if ((NB % P::G == 0) || b < NB) {
            const float2* src = buf + P::phys(b);
            static_for<R>([&](auto r) {
                v[decltype(rd)::value * R + decltype(r)::value] = src[decltype(r)::value * RS];
            });
        }
    });
}

template <class P, int PASS>
MLXA_HD void pass_compute(int g, float2* v, const float2* __restrict__ tw) {
    constexpr int R = P::radix(PASS), NB = P::nb(PASS), RD = P::rounds(PASS), NS = P::ns(PASS);
    static_for<RD>([&](auto rd) {
        constexpr int o = decltype(rd)::value * R;
        const int b = g + decltype(rd)::value * P::G;
        if ((NB % P::G == 0) || b < NB) {
            if constexpr (PASS > 0) {
                const float2* t = tw + P::tw_off(PASS) + (b % NS);
#ifdef MLXA_TWIDDLE_POWERS
                if constexpr (R >= 16) {
                    // legs 1, 4, 16 from the table, the others as products of at most three of their powers (phase error
                    // <= 3 roundings of a table entry): R - 4 fewer 64-bit shared-memory reads per butterfly
                    float2 p1[4], p4[4], p16[2];
                    p1[1] = t[0];
                    p1[2] = cmul(p1[1], p1[1]);
                    p1[3] = cmul(p1[2], p1[1]);
                    p4[1] = t[3 * NS];
                    p4[2] = cmul(p4[1], p4[1]);
                    p4[3] = cmul(p4[2], p4[1]);
                    if constexpr (R > 16) p16[1] = t[15 * NS];
                    static_for<R - 1>([&](auto r1) {
                        constexpr int r = decltype(r1)::value + 1;
                        constexpr int c = r & 3, bb = (r >> 2) & 3, a = r >> 4;
                        float2 w;
                        if constexpr (c > 0) {
                            w = p1[c];
                            if constexpr (bb > 0) w = cmul(w, p4[bb]);
                            if constexpr (a > 0) w = cmul(w, p16[a]);
                        } else if constexpr (bb > 0) {
                            w = p4[bb];
                            if constexpr (a > 0) w = cmul(w, p16[a]);
                        } else {
                            w = p16[a];
                        }
                        v[o + r] = cmul(v[o + r], w);
                    });
                } else
#endif
                static_for<R - 1>([&](auto r1) {
                    constexpr int r = decltype(r1)::value + 1;
                    v[o + r] = cmul(v[o + r], t[decltype(r1)::value * NS]);
                });
            }
            DftInplace<R, 1, o>::run(v);
        }
    });
}

// write into the padded exchange buffer (natural index -> phys), affine per leg
template <class P, int PASS>
MLXA_HD void pass_store_buf(int g, const float2* v, float2* buf) {
    constexpr int R = P::radix(PASS), NB = P::nb(PASS), RD = P::rounds(PASS), NS = P::ns(PASS);
    static_for<RD>([&](auto rd) {
        constexpr int o = decltype(rd)::value * R;
        const int b = g + decltype(rd)::value * P::G;
        if ((NB % P::G == 0) || b < NB) {
            if constexpr (PASS == 0) {
                float2* dst = buf + b * (P::R0 + P::PAD);
                static_for<R>([&](auto i) { dst[dft_perm(R, decltype(i)::value)] = v[o + decltype(i)::value]; });
            } else {
                constexpr int WS = NS + P::PAD * (NS / P::R0);
                float2* dst = buf + P::phys((b / NS) * NS * R + (b % NS));
                static_for<R>([&](auto i) { dst[dft_perm(R, decltype(i)::value) * WS] = v[o + decltype(i)::value]; });
            }
        }
    });
}

// last pass: natural-order, UNPADDED store (lanes write consecutive elements, so no padding is
// needed and the unpack step can index Z[k] / Z[N-k] without the phys() division)
template <class P, int PASS>
MLXA_HD void pass_store_natural(int g, const float2* v, float2* buf) {
    constexpr int R = P::radix(PASS), NB = P::nb(PASS), RD = P::rounds(PASS), NS = P::ns(PASS);
    static_assert(PASS == P::NPASS - 1, "natural store is for the last pass");
    static_for<RD>([&](auto rd) {
        constexpr int o = decltype(rd)::value * R;
        const int b = g + decltype(rd)::value * P::G;
        if ((NB % P::G == 0) || b < NB) {
            float2* dst = buf + b;  // b < NS here: index = b + k*NS
            static_for<R>([&](auto i) { dst[dft_perm(R, decltype(i)::value) * NS] = v[o + decltype(i)::value]; });
        }
    });
}

// hand natural-order results to a functor store(idx, value)
template <class P, int PASS, class StoreF>
MLXA_HD void pass_store_fn(int g, const float2* v, StoreF&& store) {
    constexpr int R = P::radix(PASS), NB = P::nb(PASS), RD = P::rounds(PASS), NS = P::ns(PASS);
    static_for<RD>([&](auto rd) {
        constexpr int o = decltype(rd)::value * R;
        const int b = g + decltype(rd)::value * P::G;
        if ((NB % P::G == 0) || b < NB) {
            const int base = (b / NS) * NS * R + (b % NS);
            static_for<R>([&](auto i) { store(base + dft_perm(R, decltype(i)::value) * NS, v[o + decltype(i)::value]); });
        }
    });
}

// ---- host-side twiddle table for a plan (double precision, rounded once) ---------------
template <class P>
inline void fill_plan_twiddles(float2* tw /* P::TW entries */) {
    for (int p = 1; p < P::NPASS; ++p) {
        const int R = P::radix(p), NS = P::ns(p);
        for (int r = 1; r < R; ++r)
            for (int bm = 0; bm < NS; ++bm) {
                const double a = -2.0 * kPi * double(bm) * double(r) / (double(NS) * double(R));
                tw[P::tw_off(p) + (r - 1) * NS + bm] = make_float2(float(__builtin_cos(a)), float(__builtin_sin(a)));
            }
    }
}

}  // namespace mlxa
