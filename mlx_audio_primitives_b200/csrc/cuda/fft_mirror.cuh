// Power spectra of a frame PAIR without a real-spectrum unpack through shared memory.
//
// Two real frames a, b ride one N-point complex transform Z = FFT(a + i b) of a two-pass plan
// N = R0 * R1 (R0 odd).  Pass 0: lane g runs the radix-R0 butterfly g (as fft_plan.cuh).  Last pass
// (radix R1 over R0 butterflies): lane g <= R0/2 runs butterfly g AND its mirror R0 - g.  Output leg k
// of butterfly g is Z[g + R0*k]; its Hermitian partner Z[N - g - R0*k] is leg R1-1-k of the mirror, so
// both members of every (k, N-k) pair sit in one lane and
//     |Z[k] + conj Z[N-k]|^2 = 4 |A[k]|^2,   |Z[k] - conj Z[N-k]|^2 = 4 |B[k]|^2
// come out of registers.  The mirror's inter-pass twiddles are conj(t_r) * W_R1^r; the W_R1^r factor is
// a rotation of the DFT output by one leg (register renaming), so one table read serves both
// butterflies.  Butterfly 0 is its own mirror; running it twice (b2 = 0, t = 1) makes the same formulas
// yield its pairs (k, R1-k).
//
// All per-lane steps are __host__ __device__: tests/emul runs them lane by lane on the CPU.
#pragma once
#include "fft_plan.cuh"
#include "params.cuh"

namespace mlxa {

constexpr int dft_pos(int R, int k) {  // array position that holds frequency k after dft_inplace<R>
    for (int i = 0; i < R; ++i)
        if (dft_perm(R, i) == k) return i;
    return -1;
}

template <class P>
struct Mirror {
    static constexpr int N = P::N, G = P::G, R0 = P::R0, R1 = P::R1;
    static_assert(P::NPASS == 2 && P::nb(0) == G && P::rounds(0) == 1, "one radix-R0 butterfly per lane in pass 0");
    static_assert((R0 & 1) && R0 / 2 + 1 <= G && (R1 % 2 == 0), "lanes 0..R0/2 own the butterfly pairs");
    static constexpr int RS = R0 + P::PAD;  // exchange-buffer stride between the legs of a last-pass butterfly
    static constexpr int OWNERS = R0 / 2 + 1;
    // frequency bin produced by leg k of owner lane g (always <= N/2)
    static MLXA_HD int row(int g, int k) { return k < R1 / 2 ? g + R0 * k : (R0 - g) + R0 * (R1 - 1 - k); }
};

// pass 0: load(r) supplies the lane's leg r (natural element g + G*r, or any cyclic rotation of the
// frame when only powers are wanted), radix-R0 DFT, store into the padded exchange buffer
template <class P, class LoadF>
MLXA_HD void mirror_pass0(int g, LoadF&& load, float2* buf) {
    using M = Mirror<P>;
    float2 v[M::R0];
    static_for<M::R0>([&](auto r) { v[decltype(r)::value] = load(r); });
    DftInplace<M::R0, 1, 0>::run(v);
    float2* dst = buf + g * M::RS;
    static_for<M::R0>([&](auto i) { dst[dft_perm(M::R0, decltype(i)::value)] = v[decltype(i)::value]; });
}

// The last pass' inter-pass twiddles of lane g are t_r = exp(-2 pi i g r / N), r = 1 .. R1-1: powers of ONE per-lane root.
// MirrorTwiddles keeps t_1, t_2, t_3, t_4, t_8, t_12 (the first three from the table, exact to one rounding; t_2, t_3, t_8,
// t_12 read from the table too, so no power carries more than one rounding) in registers for the whole persistent
// kernel and forms t_r = t_(r mod 4) * t_(r - r mod 4) on the fly: at most one complex product per twiddle instead of one
// shared-memory read per twiddle and tile.
template <class P>
struct MirrorTwiddles {
    static constexpr int R1 = Mirror<P>::R1, R0 = Mirror<P>::R0;
    static_assert(R1 == 16, "digits 1..3 and 4, 8, 12");
    float2 lo[4], hi[4];  // lo[c] = t_c (c = 1..3), hi[b] = t_(4b) (b = 1..3)
    MLXA_HD void load(int g, const float2* __restrict__ tw) {
        for (int c = 1; c < 4; ++c) lo[c] = tw[(c - 1) * R0 + g];
        for (int b = 1; b < 4; ++b) hi[b] = tw[(4 * b - 1) * R0 + g];
    }
    template <int r>
    MLXA_HD float2 get() const {
        constexpr int c = r & 3, b = r >> 2;
        if constexpr (b == 0) return lo[c];
        else if constexpr (c == 0) return hi[b];
        else return cmul(lo[c], hi[b]);
    }
};

// last pass + powers: pp[k] = (4^(p/2) |A[bin]|^p, 4^(p/2) |B[bin]|^p), bin = Mirror::row(g, k)
template <class P, int PW, class TW>
MLXA_HD void mirror_last_pass_powers(int g, const float2* buf, const TW& tw, float power, float2* pp) {
    using M = Mirror<P>;
    constexpr int R0 = M::R0, R1 = M::R1, RS = M::RS;
    float2 a[R1], m[R1];
    const float2* pa = buf + g;
    const float2* pm = buf + (g ? R0 - g : 0);
    static_for<R1>([&](auto r) {
        a[decltype(r)::value] = pa[decltype(r)::value * RS];
        m[decltype(r)::value] = pm[decltype(r)::value * RS];
    });
    static_for<R1 - 1>([&](auto r1) {
        constexpr int r = decltype(r1)::value + 1;
        const float2 t = tw.template get<r>();
        a[r] = cmul(a[r], t);
        m[r] = cmul_conj(m[r], t);
    });
    DftInplace<R1, 1, 0>::run(a);
    DftInplace<R1, 1, 0>::run(m);
    static_for<R1>([&](auto k_) {
        constexpr int k = decltype(k_)::value;
        const float2 z1 = a[dft_pos(R1, k)], z2 = m[dft_pos(R1, (R1 - k) % R1)];
        const float2 s = cadd_conj(z1, z2), d = csub_conj(z1, z2);  // 2 A[bin], 2i B[bin]
        float qa = fmaf(s.x, s.x, s.y * s.y), qb = fmaf(d.x, d.x, d.y * d.y);
        if constexpr (PW != POW_SQUARE) { qa = sqrtf(qa); qb = sqrtf(qb); }
        if constexpr (PW == POW_GENERAL) { qa = powf(qa, power); qb = powf(qb, power); }
        pp[k] = make_float2(qa, qb);
    });
}

// the table-read form (one shared-memory read per twiddle): what the last pass used before MirrorTwiddles
template <class P>
struct MirrorTwiddleTable {
    const float2* tw;
    int g;
    template <int r>
    MLXA_HD float2 get() const { return tw[(r - 1) * Mirror<P>::R0 + g]; }
};

}  // namespace mlxa
