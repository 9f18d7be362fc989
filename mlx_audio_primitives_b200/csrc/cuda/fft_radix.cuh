// In-register forward DFTs of compile-time length R on a thread-private float2 array.
// Everything is resolved at compile time: indices are template constants, twiddles are
// immediates.  Composite lengths recurse (Cooley-Tukey, A x B with A the first small factor);
// results are left in a digit-permuted order described by dft_perm<R>(i), which the callers
// fold into their store addresses (register renaming, zero instructions).
#pragma once
#include "common.cuh"

namespace mlxa {

constexpr bool is_base_radix(int R) { return R == 1 || R == 2 || R == 3 || R == 4 || R == 5; }
constexpr int first_factor(int R) {
    return (R % 4 == 0) ? 4 : (R % 2 == 0) ? 2 : (R % 3 == 0) ? 3 : (R % 5 == 0) ? 5 : R;
}
// frequency index held at array position i after dft_inplace<R>
constexpr int dft_perm(int R, int i) {
    if (is_base_radix(R)) return i;
    const int A = first_factor(R), B = R / A;
    return dft_perm(A, i / B) + A * dft_perm(B, i % B);
}

// operates on v[OFF + S*i], i in [0, R)
template <int R, int S, int OFF>
struct DftInplace;

template <int S, int OFF>
struct DftInplace<1, S, OFF> {
    static MLXA_HD void run(float2*) {}
};

template <int S, int OFF>
struct DftInplace<2, S, OFF> {
    static MLXA_HD void run(float2* v) {
        const float2 a = v[OFF], b = v[OFF + S];
        v[OFF] = cadd(a, b);
        v[OFF + S] = csub(a, b);
    }
};

template <int S, int OFF>
struct DftInplace<4, S, OFF> {
    static MLXA_HD void run(float2* v) {
        const float2 a0 = v[OFF], a1 = v[OFF + S], a2 = v[OFF + 2 * S], a3 = v[OFF + 3 * S];
        const float2 t0 = cadd(a0, a2), t1 = csub(a0, a2);
        const float2 t2 = cadd(a1, a3), d = csub(a1, a3);
        v[OFF] = cadd(t0, t2);
        v[OFF + S] = cadd_rot(t1, d);      // t1 + (-i) d
        v[OFF + 2 * S] = csub(t0, t2);
        v[OFF + 3 * S] = csub_rot(t1, d);  // t1 - (-i) d
    }
};

template <int S, int OFF>
struct DftInplace<3, S, OFF> {
    static MLXA_HD void run(float2* v) {
        constexpr float s3 = 0.86602540378443864676f;
        const float2 a0 = v[OFF], a1 = v[OFF + S], a2 = v[OFF + 2 * S];
        const float2 t = cadd(a1, a2), d = cscale(csub(a1, a2), s3);
        const float2 m = caxpy(-0.5f, t, a0);
        v[OFF] = cadd(a0, t);
        v[OFF + S] = cadd_rot(m, d);      // m + (-i*s3)*(a1 - a2)
        v[OFF + 2 * S] = csub_rot(m, d);
    }
};

template <int S, int OFF>
struct DftInplace<5, S, OFF> {
    static MLXA_HD void run(float2* v) {
        constexpr float c1 = float(cos_turn(1, 5)), c2 = float(cos_turn(2, 5));
        constexpr float s1 = float(sin_turn(1, 5)), s2 = float(sin_turn(2, 5));
        const float2 a0 = v[OFF], a1 = v[OFF + S], a2 = v[OFF + 2 * S], a3 = v[OFF + 3 * S],
                     a4 = v[OFF + 4 * S];
        const float2 t1 = cadd(a1, a4), t2 = cadd(a2, a3), t3 = csub(a1, a4), t4 = csub(a2, a3);
        const float2 m1 = caxpy(c2, t2, caxpy(c1, t1, a0));
        const float2 m2 = caxpy(c1, t2, caxpy(c2, t1, a0));
        const float2 n1 = caxpy(s2, t4, cscale(t3, s1));
        const float2 n2 = caxpy(-s1, t4, cscale(t3, s2));
        v[OFF] = cadd(cadd(a0, t1), t2);
        v[OFF + S] = cadd_rot(m1, n1);      // m1 + (-i) n1
        v[OFF + 4 * S] = csub_rot(m1, n1);
        v[OFF + 2 * S] = cadd_rot(m2, n2);
        v[OFF + 3 * S] = csub_rot(m2, n2);
    }
};

template <int R, int S, int OFF>
struct DftInplace {
    static constexpr int A = first_factor(R), B = R / A;
    static_assert(A < R, "unsupported prime radix");
    static MLXA_HD void run(float2* v) {
        // 1. B transforms of length A over the stride-(S*B) subsequences x[n1*B + n2]
        static_for<B>([&](auto n2) { DftInplace<A, S * B, OFF + decltype(n2)::value * S>::run(v); });
        // 2. twiddle by W_R^(k1*n2); position n1p holds k1 = dft_perm(A, n1p)
        static_for<A>([&](auto n1p) {
            static_for<B>([&](auto n2) {
                constexpr int i = decltype(n1p)::value * B + decltype(n2)::value;
                constexpr int k1 = dft_perm(A, decltype(n1p)::value);
                v[OFF + S * i] = mul_tw<k1 * decltype(n2)::value, R>(v[OFF + S * i]);
            });
        });
        // 3. A transforms of length B over the contiguous runs
        static_for<A>([&](auto n1p) { DftInplace<B, S, OFF + decltype(n1p)::value * B * S>::run(v); });
    }
};

template <int R>
MLXA_HD void dft_inplace(float2* v) { DftInplace<R, 1, 0>::run(v); }

}  // namespace mlxa
