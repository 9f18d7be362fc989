// Autocorrelation pitch detector for one compiled transform size (built once per packed plan with
// -DMLXA_NFFT=<n_fft>), reference pitch.py:118-260 (a per-frame NumPy loop on the host there).
//
// One lane group per frame, everything between the samples and (f0, voiced) stays on the SM:
//   frame - mean, zero-padded to n_fft  ->  real FFT (packed plan)  ->  |X|^2  ->  real FFT again (the inverse
//   transform of a real, even spectrum is its forward transform / n_fft)  ->  r / r[0]  ->  first local maximum above
//   the threshold in [sr/fmax, sr/fmin], else the range's global maximum if above the threshold.
// The second consumer of the FFT engine.  n_fft is the power of two >= 2*frame_length - 1.
#include "fft_plans_list.cuh"
#include "params.cuh"

#ifndef MLXA_NFFT
#error "compile with -DMLXA_NFFT=<n_fft>"
#endif

namespace mlxa {
namespace {

using PF = PlanFor<MLXA_NFFT>;
using P = PF::Plan;
constexpr int NFFT = MLXA_NFFT;
static_assert(PF::MODE == MODE_PACK && P::G <= 32, "packed plans whose lane groups fit a warp");
constexpr int N = P::N;  // complex points
constexpr int THREADS = (P::E > 32) ? 256 : ((P::G == 32) ? 512 : 256);
constexpr int NG = THREADS / P::G;
constexpr int NQ = ceil_div(N + 1, P::G);

template <class F>
MLXA_D void forward_natural(int g, int gi, float2* v, float2* buf, const float2* tw, F&& load) {
    pass_load_fn<P, 0>(g, v, load);
    group_sync<P::G>(gi);  // every lane has read its inputs (they may live in buf)
    pass_compute<P, 0>(g, v, tw);
    pass_store_buf<P, 0>(g, v, buf);
    group_sync<P::G>(gi);
    pass_load_buf<P, 1>(g, v, buf);
    group_sync<P::G>(gi);
    pass_compute<P, 1>(g, v, tw);
    if constexpr (P::NPASS == 3) {
        pass_store_buf<P, 1>(g, v, buf);
        group_sync<P::G>(gi);
        pass_load_buf<P, 2>(g, v, buf);
        group_sync<P::G>(gi);
        pass_compute<P, 2>(g, v, tw);
    }
    pass_store_natural<P, P::NPASS - 1>(g, v, buf);  // Z[k] at buf[k]
    group_sync<P::G>(gi);
}
// bin k (0..N) of the real transform whose packed complex transform sits in buf
MLXA_D float2 real_bin(const float2* buf, const float2* __restrict__ tw_unpack, int k) {
    const float2 zk = buf[k == N ? 0 : k], zm = buf[k == 0 ? 0 : N - k];
    const float2 E = cadd_conj(zk, zm), D = csub_conj(zk, zm);
    return caxpy(0.5f, E, cmul(mul_neg_i(D), __ldg(tw_unpack + k)));
}

__global__ void __launch_bounds__(THREADS) acf_pitch_kernel(const AcfParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int gi = threadIdx.x / P::G, g = threadIdx.x % P::G;
    float2* buf = reinterpret_cast<float2*>(smem_raw) + gi * P::BUF;
    float* fb = reinterpret_cast<float*>(buf);
    const unsigned gmask = (P::G == 32) ? 0xffffffffu : (((1u << (P::G & 31)) - 1u) << (threadIdx.x & 31 & ~(P::G - 1)));
    auto gsum = [&](float x) {
#pragma unroll
        for (int o = P::G / 2; o > 0; o >>= 1) x += __shfl_xor_sync(gmask, x, o, P::G);
        return x;
    };
    const long long frame = (long long)blockIdx.x * NG + gi;
    const bool live = frame < (long long)p.B * p.T;
    const long long b = live ? frame / p.T : 0, t = live ? frame - b * p.T : 0;
    const float* yb = p.y + b * p.ldy;
    const long long s0 = t * p.hop - p.pad;  // clip index of the frame's first sample (zeros outside the clip)
    auto sample = [&](int j) {
        const long long q = s0 + j;
        return (j < p.frame_length && q >= 0 && q < p.L) ? __ldg(yb + q) : 0.f;
    };
    float sum = 0.f;
    for (int j = g; j < p.frame_length; j += P::G) sum += sample(j);
    const float mean = gsum(sum) / float(p.frame_length);

    float2 v[P::E];
    // 1. real FFT of the centred, zero-padded frame (packed: z[n] = x[2n] + i x[2n+1])
    forward_natural(g, gi, v, buf, p.tw_plan, [&](int n) {
        const int j = 2 * n;
        return make_float2(j < p.frame_length ? sample(j) - mean : 0.f, j + 1 < p.frame_length ? sample(j + 1) - mean : 0.f);
    });
    // 2. power spectrum, laid out as the real, even sequence P[0..n_fft) in the group's buffer
    float pk[NQ];
    static_for<NQ>([&](auto q) {
        constexpr int Q = decltype(q)::value;
        const int k = g + Q * P::G;
        if (k <= N) {
            const float2 X = real_bin(buf, p.tw_unpack, k);
            pk[Q] = fmaf(X.x, X.x, X.y * X.y);
        }
    });
    group_sync<P::G>(gi);
    static_for<NQ>([&](auto q) {
        constexpr int Q = decltype(q)::value;
        const int k = g + Q * P::G;
        if (k <= N) {
            fb[k] = pk[Q];
            if (k > 0 && k < N) fb[NFFT - k] = pk[Q];
        }
    });
    group_sync<P::G>(gi);
    // 3. r = irfft(P) = Re(rfft(P)) / n_fft (P real and even)
    forward_natural(g, gi, v, buf, p.tw_plan, [&](int n) { return buf[n]; });
    static_for<NQ>([&](auto q) {
        constexpr int Q = decltype(q)::value;
        const int k = g + Q * P::G;
        if (k <= N) pk[Q] = real_bin(buf, p.tw_unpack, k).x * (1.0f / float(NFFT));
    });
    group_sync<P::G>(gi);
    static_for<NQ>([&](auto q) {
        constexpr int Q = decltype(q)::value;
        const int k = g + Q * P::G;
        if (k <= N) fb[k] = pk[Q];
    });
    group_sync<P::G>(gi);
    // 4. peak picking on r / r[0] (pitch.py:223-258)
    const float r0 = fb[0];
    float f0 = 0.f;
    int voiced = 0;
    const int len = p.max_lag - p.min_lag + 1;  // search range r[min_lag .. max_lag]
    if (p.voiced == nullptr) {  // periodicity (pitch.py:267-383): the maximum of r / r[0] over the range, 0 for silent frames
        float best = 0.f;
        if (r0 > 1e-10f && len > 0) {
            best = -INFINITY;
            for (int i = g; i < len; i += P::G) best = fmaxf(best, fb[p.min_lag + i] / r0);
#pragma unroll
            for (int o = P::G / 2; o > 0; o >>= 1) best = fmaxf(best, __shfl_xor_sync(gmask, best, o, P::G));
        }
        if (live && g == 0) p.f0[frame] = best;
        return;
    }
    if (r0 > 1e-10f && len > 0) {  // (group-uniform)
        auto rn = [&](int lag) { return fb[lag] / r0; };
        int first = 0x7fffffff;
        for (int i = 1 + g; i < len - 1; i += P::G) {
            const int lag = p.min_lag + i;
            const float c = rn(lag);
            if (c > rn(lag - 1) && c > rn(lag + 1) && c > p.threshold) { first = lag; break; }
        }
#pragma unroll
        for (int o = P::G / 2; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(gmask, first, o, P::G));
        if (first == 0x7fffffff) {  // no local maximum: the global maximum of the range, first index on ties
            float best = -INFINITY;
            int arg = 0x7fffffff;
            for (int i = g; i < len; i += P::G) {
                const float c = rn(p.min_lag + i);
                if (c > best) { best = c; arg = p.min_lag + i; }
            }
#pragma unroll
            for (int o = P::G / 2; o > 0; o >>= 1) {
                const float ob = __shfl_xor_sync(gmask, best, o, P::G);
                const int oa = __shfl_xor_sync(gmask, arg, o, P::G);
                if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
            }
            if (best > p.threshold) first = arg;
        }
        if (first != 0x7fffffff && first > 0) {
            f0 = float(double(p.sr) / double(first));
            voiced = 1;
        }
    }
    if (live && g == 0) {
        p.f0[frame] = f0;
        p.voiced[frame] = (unsigned char)voiced;
    }
}

}  // namespace

#define MLXA_CAT2(a, b) a##b
#define MLXA_CAT(a, b) MLXA_CAT2(a, b)

cudaError_t MLXA_CAT(launch_acf_, MLXA_NFFT)(const AcfParams& p, cudaStream_t s) {
    const size_t smem = size_t(NG) * P::BUF * 8;
    cudaError_t e = cudaFuncSetAttribute(acf_pitch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const long long frames = (long long)p.B * p.T, grid = (frames + NG - 1) / NG;
    if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    acf_pitch_kernel<<<(unsigned)grid, THREADS, smem, s>>>(p);
    return cudaGetLastError();
}

}  // namespace mlxa
