// CPU emulation of the warp-group FFT engine (test infrastructure).  The per-lane pass
// functions in csrc/cuda/fft_plan.cuh are __host__ __device__; here the lanes of a group are
// executed one after another with the phases separated exactly where the kernels synchronise.
#include <cmath>
#include <vector>
#include "fft_plans_list.cuh"
#include "fft_mirror.cuh"
#include "pcg64.cuh"

using namespace mlxa;

template <class P>
static void run_plan(const float2* in, float2* out) {
    std::vector<float2> buf(P::BUF, make_float2(0.f, 0.f)), tw(P::TW > 0 ? P::TW : 1);
    fill_plan_twiddles<P>(tw.data());
    std::vector<std::vector<float2>> regs(P::G, std::vector<float2>(P::E));
    for (int g = 0; g < P::G; ++g) {
        float2* v = regs[g].data();
        pass_load_fn<P, 0>(g, v, [&](int idx) { return in[idx]; });
        pass_compute<P, 0>(g, v, tw.data());
        pass_store_buf<P, 0>(g, v, buf.data());
    }
    for (int g = 0; g < P::G; ++g) pass_load_buf<P, 1>(g, regs[g].data(), buf.data());
    for (int g = 0; g < P::G; ++g) {
        pass_compute<P, 1>(g, regs[g].data(), tw.data());
        pass_store_buf<P, 1>(g, regs[g].data(), buf.data());
    }
    if constexpr (P::NPASS == 3) {
        for (int g = 0; g < P::G; ++g) pass_load_buf<P, 2>(g, regs[g].data(), buf.data());
        for (int g = 0; g < P::G; ++g) {
            pass_compute<P, 2>(g, regs[g].data(), tw.data());
            pass_store_buf<P, 2>(g, regs[g].data(), buf.data());
        }
    }
    for (int k = 0; k < P::N; ++k) out[k] = buf[P::phys(k)];
}

template <int R>
static void run_radix(const float2* in, float2* out) {
    float2 v[R];
    for (int i = 0; i < R; ++i) v[i] = in[i];
    dft_inplace<R>(v);
    for (int i = 0; i < R; ++i) out[dft_perm(R, i)] = v[i];
}

// mirror-paired power spectra of a frame pair (fft_mirror.cuh), lanes run one after another; rot != 0
// feeds the frames cyclically rotated by G samples, as the odd lane group of a warp does
template <class P>
static void run_mirror(const float* fa, const float* fb, const float* win, int rot, float* pa, float* pb) {
    using M = Mirror<P>;
    std::vector<float2> buf(P::BUF, make_float2(0.f, 0.f)), tw(P::TW);
    fill_plan_twiddles<P>(tw.data());
    for (int g = 0; g < P::G; ++g)
        mirror_pass0<P>(g, [&](auto r_) {
            const int n = (g + P::G * (decltype(r_)::value + (rot ? 1 : 0))) % P::N;
            return make_float2(fa[n] * win[n], fb[n] * win[n]);
        }, buf.data());
    for (int g = 0; g < M::OWNERS; ++g) {
        float2 pp[P::R1];
        MirrorTwiddles<P> mtw;
        mtw.load(g, tw.data());
        mirror_last_pass_powers<P, POW_SQUARE>(g, buf.data(), mtw, 2.f, pp);
        for (int k = 0; k < P::R1; ++k) { pa[M::row(g, k)] = pp[k].x; pb[M::row(g, k)] = pp[k].y; }
    }
}

extern "C" {
int emul_mirror_powers_400(const float* fa, const float* fb, const float* win, int rot, float* pa, float* pb) {
    run_mirror<PlanFor<400>::Plan>(fa, fb, win, rot, pa, pb);
    return 0;
}
// the same __host__ __device__ PCG64 code the device kernel runs, including the per-chunk jump-ahead
void emul_pcg64_uniform(unsigned long long s_hi, unsigned long long s_lo, unsigned long long i_hi, unsigned long long i_lo,
                        double low, double high, long long n, int chunk, float* out) {
    const u128 inc = ((u128)i_hi << 64) | i_lo, s0 = ((u128)s_hi << 64) | s_lo;
    for (long long i0 = 0; i0 < n; i0 += chunk) {
        u128 st = pcg_advance(s0, inc, (unsigned long long)i0);
        for (long long j = i0; j < n && j < i0 + chunk; ++j) out[j] = pcg_uniform_f32(st, inc, low, high - low);
    }
}
int emul_plan_length(int n_fft) {
    switch (n_fft) {
#define X(NF) case NF: return PlanFor<NF>::Plan::N;
        MLXA_FOR_EACH_NFFT(X)
#undef X
    }
    return -1;
}
int emul_plan_fft(int n_fft, const float* in, float* out) {
    switch (n_fft) {
#define X(NF) case NF: run_plan<PlanFor<NF>::Plan>((const float2*)in, (float2*)out); return 0;
        MLXA_FOR_EACH_NFFT(X)
#undef X
    }
    return -1;
}
int emul_radix(int R, const float* in, float* out) {
    switch (R) {
#define X(RR) case RR: run_radix<RR>((const float2*)in, (float2*)out); return 0;
        X(2) X(3) X(4) X(5) X(8) X(9) X(10) X(16) X(20) X(25) X(32) X(64)
#undef X
    }
    return -1;
}
}
