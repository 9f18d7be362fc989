// Shared helpers for the sm_100a spectral kernels: compile-time loops, constexpr
// trigonometry (so every in-register twiddle becomes an FFMA immediate), error plumbing.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <type_traits>
#include <utility>

#define MLXA_HD __host__ __device__ __forceinline__
#define MLXA_D __device__ __forceinline__

namespace mlxa {

// ---- compile-time loop: f(integral_constant<int, I>) for I in [0, N) ----------------------
template <class F, int... I>
MLXA_HD void static_for_impl(F&& f, std::integer_sequence<int, I...>) {
    (f(std::integral_constant<int, I>{}), ...);
}
template <int N, class F>
MLXA_HD void static_for(F&& f) {
    static_for_impl(static_cast<F&&>(f), std::make_integer_sequence<int, N>{});
}

// ---- constexpr cos / sin of (2*pi*num/den), exact octant reduction on integers ------------
constexpr double kPi = 3.141592653589793238462643383279502884;

constexpr double taylor_cos(double x) {  // |x| <= pi/4
    double x2 = x * x, term = 1.0, sum = 1.0;
    for (int i = 1; i <= 12; ++i) {
        term *= -x2 / double((2 * i - 1) * (2 * i));
        sum += term;
    }
    return sum;
}
constexpr double taylor_sin(double x) {  // |x| <= pi/4
    double x2 = x * x, term = x, sum = x;
    for (int i = 1; i <= 12; ++i) {
        term *= -x2 / double((2 * i) * (2 * i + 1));
        sum += term;
    }
    return sum;
}
// angle = 2*pi*m/(8*d); m any integer.  8d units make half/quarter/eighth turns integers.
constexpr double cos_units8(long long m, long long d) {
    const long long full = 8 * d;
    m %= full;
    if (m < 0) m += full;
    if (m > 4 * d) m = full - m;           // cos(-x) = cos(x)          -> [0, half turn]
    double sign = 1.0;
    if (m > 2 * d) { m = 4 * d - m; sign = -1.0; }  // cos(pi - x) = -cos x -> [0, quarter]
    if (m == 0) return sign;
    if (m == 2 * d) return 0.0;
    if (m > d) return sign * taylor_sin(2.0 * kPi * double(2 * d - m) / double(full));
    return sign * taylor_cos(2.0 * kPi * double(m) / double(full));
}
constexpr double cos_turn(long long num, long long den) { return cos_units8(8 * num, den); }
constexpr double sin_turn(long long num, long long den) { return cos_units8(8 * num - 2 * den, den); }

// ---- complex helpers on float2 ---------------------------------------------------------------
// On the device every helper is ONE or TWO packed FP32 instructions (sm_100 FADD2 / FMUL2 / FFMA2 on an
// aligned register pair {re, im}; PTX add/mul/fma.rn.f32x2).  ptxas folds the component swaps and
// per-half sign changes written below as pack(...) arguments into operand modifiers (.LO_HI, .NP), so a
// complex add or a multiplication by -i costs one issue slot instead of two.  A packed instruction
// occupies the FMA pipe for two passes -- the flop rate is unchanged (tools/probes/f32x2_probe.cu) --
// but the transform kernels are issue-bound, not pipe-bound.  The host versions (CPU emulation of the
// kernels, tests/emul) perform the same roundings in the same order.
#if defined(__CUDACC__)
typedef unsigned long long mlxa_u64;
MLXA_D mlxa_u64 pk2(float a, float b) { mlxa_u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
MLXA_D mlxa_u64 pk2(float2 a) { return pk2(a.x, a.y); }
MLXA_D float2 up2(mlxa_u64 v) { float2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
MLXA_D mlxa_u64 add2(mlxa_u64 a, mlxa_u64 b) { mlxa_u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
MLXA_D mlxa_u64 mul2(mlxa_u64 a, mlxa_u64 b) { mlxa_u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
MLXA_D mlxa_u64 fma2(mlxa_u64 a, mlxa_u64 b, mlxa_u64 c) { mlxa_u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
#endif

// (a.x + b.x, a.y + b.y) with the second operand's halves optionally swapped / negated
MLXA_HD float2 padd(float2 a, float bx, float by) {
#if defined(__CUDA_ARCH__)
    return up2(add2(pk2(a), pk2(bx, by)));
#else
    return make_float2(a.x + bx, a.y + by);
#endif
}
MLXA_HD float2 cadd(float2 a, float2 b) { return padd(a, b.x, b.y); }
MLXA_HD float2 csub(float2 a, float2 b) { return padd(a, -b.x, -b.y); }
MLXA_HD float2 cadd_rot(float2 a, float2 b) { return padd(a, b.y, -b.x); }    // a + (-i) b
MLXA_HD float2 csub_rot(float2 a, float2 b) { return padd(a, -b.y, b.x); }    // a - (-i) b
MLXA_HD float2 cadd_conj(float2 a, float2 b) { return padd(a, b.x, -b.y); }   // a + conj b
MLXA_HD float2 csub_conj(float2 a, float2 b) { return padd(a, -b.x, b.y); }   // a - conj b
// (a.x * s.x, a.y * s.y) and (a.x * s.x + c.x, a.y * s.y + c.y)
MLXA_HD float2 pmul(float2 a, float sx, float sy) {
#if defined(__CUDA_ARCH__)
    return up2(mul2(pk2(a), pk2(sx, sy)));
#else
    return make_float2(a.x * sx, a.y * sy);
#endif
}
MLXA_HD float2 pfma(float ax, float ay, float sx, float sy, float2 c) {
#if defined(__CUDA_ARCH__)
    return up2(fma2(pk2(ax, ay), pk2(sx, sy), pk2(c)));
#else
    return make_float2(fmaf(ax, sx, c.x), fmaf(ay, sy, c.y));
#endif
}
MLXA_HD float2 cscale(float2 a, float s) { return pmul(a, s, s); }                       // s * a
MLXA_HD float2 caxpy(float s, float2 a, float2 c) { return pfma(a.x, a.y, s, s, c); }    // c + s * a
MLXA_HD float2 cmul(float2 a, float2 b) {  // a.x * (b.x, b.y) + a.y * (-b.y, b.x)
    return pfma(a.y, a.y, -b.y, b.x, pmul(make_float2(a.x, a.x), b.x, b.y));
}
MLXA_HD float2 cmul_conj(float2 a, float2 b) {  // a * conj(b) = a.x * (b.x, -b.y) + a.y * (b.y, b.x)
    return pfma(a.y, a.y, b.y, b.x, pmul(make_float2(a.x, a.x), b.x, -b.y));
}
MLXA_HD float2 mul_neg_i(float2 a) { return make_float2(a.y, -a.x); }  // a * (-i)
MLXA_HD float2 mul_pos_i(float2 a) { return make_float2(-a.y, a.x); }  // a * (+i)
MLXA_HD float2 cswap(float2 a) { return make_float2(a.y, a.x); }

// v * exp(-2*pi*i*NUM/DEN) with every special angle folded at compile time
template <int NUM, int DEN>
MLXA_HD float2 mul_tw(float2 v) {
    constexpr int n = ((NUM % DEN) + DEN) % DEN;
    constexpr float h = 0.70710678118654752440f;
    if constexpr (n == 0) {
        return v;
    } else if constexpr (4 * n == DEN) {
        return mul_neg_i(v);
    } else if constexpr (2 * n == DEN) {
        return make_float2(-v.x, -v.y);
    } else if constexpr (4 * n == 3 * DEN) {
        return mul_pos_i(v);
    } else if constexpr (8 * n == DEN) {  // (1 - i)/sqrt2: h * (x + y, y - x)
        return cscale(padd(v, v.y, -v.x), h);
    } else if constexpr (8 * n == 3 * DEN) {  // (-1 - i)/sqrt2: h * (y - x, -(x + y))
        return cscale(padd(make_float2(-v.x, -v.y), v.y, -v.x), h);
    } else if constexpr (8 * n == 5 * DEN) {  // (-1 + i)/sqrt2: h * (-(x + y), x - y)
        return cscale(padd(make_float2(-v.x, -v.y), -v.y, v.x), h);
    } else if constexpr (8 * n == 7 * DEN) {  // (1 + i)/sqrt2: h * (x - y, x + y)
        return cscale(padd(v, -v.y, v.x), h);
    } else {
        constexpr float c = float(cos_turn(n, DEN));
        constexpr float s = float(sin_turn(n, DEN));
        // (x + iy)(c - is) = (xc + ys) + i(yc - xs)
        return pfma(v.x, v.y, c, c, pmul(make_float2(v.y, v.x), s, -s));
    }
}

constexpr int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace mlxa

// ---- sm_90+/sm_100 async-proxy helpers: mbarrier + 1-D bulk copy (TMA, SASS UBLKCP) ---------
#if defined(__CUDACC__)
namespace mlxa {
MLXA_D uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
MLXA_D void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
MLXA_D void mbar_arrive_one(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
MLXA_D void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared bulk copy; dst/src 16-byte aligned, bytes % 16 == 0; completion lands on bar
MLXA_D void bulk_copy_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// Ampere-style per-thread async copies (LDGSTS), 8 bytes each; used to prefetch a spectrum frame
MLXA_D void cp_async8(void* dst_smem, const void* src_gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
MLXA_D void cp_async4(void* dst_smem, const void* src_gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
MLXA_D void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
MLXA_D void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
MLXA_D void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
MLXA_D void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Synchronise the G lanes that share a transform: a (sub-)warp for G <= 32, the two warps of a 64-lane group
// through named barrier 1 + group index otherwise (barrier 0 is __syncthreads; at most 15 groups per CTA).
template <int G>
MLXA_D void group_sync(int group_in_cta) {
    if constexpr (G <= 32) {
        __syncwarp();
    } else {
        asm volatile("bar.sync %0, %1;" ::"r"(group_in_cta + 1), "n"(G) : "memory");
    }
}

MLXA_D bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// bounded spin: a copy that never lands traps instead of hanging the GPU
MLXA_D void mbar_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
        if (spins > (1u << 24)) __trap();
}
}  // namespace mlxa
#endif
