// Fused irFFT -> window -> overlap-add -> normalise -> trim kernel for one compiled n_fft
// (built once per size with -DMLXA_NFFT=<n_fft>).
//
// One CTA = one clip x one tile of tile_hops*hop OUTPUT samples.  It inverse-transforms every
// frame that touches the tile (the r-1 = ceil(n_fft/hop)-1 frames before the tile are
// recomputed as a halo), adds the windowed frames into a shared-memory accumulator and writes
// each output sample once, already divided by max(sum w^2, 1e-8) and shifted by the centre
// trim.  No atomics: within a round, frames whose index differs by a multiple of r cannot
// overlap, so the adds run in r barrier-separated phases and the summation order is fixed
// (deterministic results).  The (B, T, n_fft) frame tensor of the reference
// (stft.py:295 -> overlap_add.metal:16) never exists.
#include "fft_plans_list.cuh"
#include "params.cuh"

#ifndef MLXA_NFFT
#error "compile with -DMLXA_NFFT=<n_fft>"
#endif

namespace mlxa {
namespace {  // per-translation-unit kernels: every n_fft gets its own copy

using PF = PlanFor<MLXA_NFFT>;
using P = PF::Plan;
constexpr int NFFT = MLXA_NFFT;
constexpr int FPT = (PF::MODE == MODE_PAIR) ? 2 : 1;
constexpr int THREADS = (P::E > 32) ? 128 : 256;
constexpr int NG = THREADS / P::G;

__global__ void __launch_bounds__(THREADS) inv_kernel(const InvParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int TS = p.tile_hops * p.hop;
    float* s_acc = reinterpret_cast<float*>(smem_raw);
    float* s_win = s_acc + ((TS + 3) & ~3);
    float2* s_buf = reinterpret_cast<float2*>(s_win + NFFT);

    const int b = blockIdx.y;
    const long long o0 = (long long)blockIdx.x * TS;
    for (int i = threadIdx.x; i < TS; i += THREADS) s_acc[i] = 0.f;
    for (int i = threadIdx.x; i < NFFT; i += THREADS) s_win[i] = __ldg(p.window + i);

    // frames touching [o0, min(o0 + TS, ola_len))
    const long long o_end = min(o0 + (long long)TS, p.ola_len);
    const int r = (NFFT + p.hop - 1) / p.hop;
    int f_lo = (o0 < NFFT) ? 0 : int((o0 - NFFT) / p.hop) + 1;
    int f_hi = (o_end > o0) ? int(min((long long)(p.T - 1), (o_end - 1) / p.hop)) : -1;
    __syncthreads();

    const int gi = threadIdx.x / P::G, g = threadIdx.x % P::G;
    float2* buf = s_buf + gi * P::BUF;
    const float2* specb = p.spec + (long long)b * p.T * p.F_in;
    const float inv_n = 1.0f / float(P::N);

    for (int base = f_lo; base <= f_hi; base += NG * FPT) {
        const int fa = base + gi * FPT;
        const bool va = fa <= f_hi;
        float2 v[P::E];

        // ---- spectrum -> packed complex input (swapped: inverse = swap . forward . swap) ---
        if constexpr (PF::MODE == MODE_PACK) {
            constexpr int N = P::N;
            const float2* X = specb + (long long)(va ? fa : 0) * p.F_in;
            for (int k = g; k < N; k += P::G) {
                float2 xk = make_float2(0.f, 0.f), xm = make_float2(0.f, 0.f);
                if (va && k < p.F_in) xk = __ldg(X + k);
                if (va && N - k < p.F_in) xm = __ldg(X + N - k);
                if (k == 0) { xk.y = 0.f; xm.y = 0.f; }  // c2r ignores imag of DC / Nyquist
                const float2 w = __ldg(p.tw_unpack + k);
                const float ex = 0.5f * (xk.x + xm.x), ey = 0.5f * (xk.y - xm.y);
                const float2 d = make_float2(0.5f * (xk.x - xm.x), 0.5f * (xk.y + xm.y));
                const float2 o = cmul_conj(d, w);
                buf[k] = make_float2(ey + o.x, ex - o.y);  // swap(E + i*O)
            }
        } else {
            constexpr int N = P::N;
            const bool vb = fa + 1 <= f_hi;
            const float2* Xa = specb + (long long)(va ? fa : 0) * p.F_in;
            const float2* Xb = specb + (long long)(vb ? fa + 1 : 0) * p.F_in;
            for (int k = g; k <= N / 2; k += P::G) {
                float2 a = make_float2(0.f, 0.f), c = make_float2(0.f, 0.f);
                if (va && k < p.F_in) a = __ldg(Xa + k);
                if (vb && k < p.F_in) c = __ldg(Xb + k);
                if (k == 0 || 2 * k == N) { a.y = 0.f; c.y = 0.f; }
                buf[k] = make_float2(a.y + c.x, a.x - c.y);  // swap(Xa + i*Xb)
                if (k > 0 && 2 * k < N) buf[N - k] = make_float2(c.x - a.y, a.x + c.y);  // swap(conj Xa + i*conj Xb)
            }
        }
        __syncwarp();
        pass_load_fn<P, 0>(g, v, [&](int n) { return buf[n]; });
        __syncwarp();
        pass_compute<P, 0>(g, v, p.tw_plan);
        pass_store_buf<P, 0>(g, v, buf);
        __syncwarp();
        pass_load_buf<P, 1>(g, v, buf);
        __syncwarp();
        pass_compute<P, 1>(g, v, p.tw_plan);
        if constexpr (P::NPASS == 3) {
            pass_store_buf<P, 1>(g, v, buf);
            __syncwarp();
            pass_load_buf<P, 2>(g, v, buf);
            __syncwarp();
            pass_compute<P, 2>(g, v, p.tw_plan);
        }

        // ---- windowed overlap-add into the tile accumulator, r conflict-free phases -------
        for (int ph = 0; ph < r; ++ph) {
            if constexpr (PF::MODE == MODE_PACK) {
                if (va && ((fa - f_lo) % r) == ph) {
                    const long long off = (long long)fa * p.hop - o0;
                    pass_store_fn<P, P::NPASS - 1>(g, v, [&](int n, float2 val) {
                        const long long q = off + 2 * n;
                        if (q >= 0 && q < TS) s_acc[q] = fmaf(s_win[2 * n], val.y * inv_n, s_acc[q]);
                        if (q + 1 >= 0 && q + 1 < TS) s_acc[q + 1] = fmaf(s_win[2 * n + 1], val.x * inv_n, s_acc[q + 1]);
                    });
                }
            } else {
                if (va && ((fa - f_lo) % r) == ph) {
                    const long long off = (long long)fa * p.hop - o0;
                    pass_store_fn<P, P::NPASS - 1>(g, v, [&](int n, float2 val) {
                        const long long q = off + n;
                        if (q >= 0 && q < TS) s_acc[q] = fmaf(s_win[n], val.y * inv_n, s_acc[q]);
                    });
                }
                if ((fa + 1 <= f_hi) && ((fa + 1 - f_lo) % r) == ph) {
                    const long long off = (long long)(fa + 1) * p.hop - o0;
                    pass_store_fn<P, P::NPASS - 1>(g, v, [&](int n, float2 val) {
                        const long long q = off + n;
                        if (q >= 0 && q < TS) s_acc[q] = fmaf(s_win[n], val.x * inv_n, s_acc[q]);
                    });
                }
            }
            __syncthreads();
        }
    }
    __syncthreads();

    // ---- normalise, trim, store ---------------------------------------------------------
    float* yb = p.y + (long long)b * p.ldy;
    for (int i = threadIdx.x; i < TS; i += THREADS) {
        const long long o = o0 + i;
        const long long j = o - p.trim;
        if (j < 0 || j >= p.out_len) continue;
        float val = 0.f;
        if (o < p.ola_len) val = s_acc[i] / fmaxf(__ldg(p.wss + o), 1e-8f);
        yb[j] = val;
    }
}

static size_t inv_smem_bytes(int hop, int TH) {
    return size_t((TH * hop + 3) & ~3) * 4 + size_t(NFFT) * 4 + size_t(NG) * P::BUF * 8;
}

}  // namespace

#define MLXA_CAT2(a, b) a##b
#define MLXA_CAT(a, b) MLXA_CAT2(a, b)

cudaError_t MLXA_CAT(launch_inv_, MLXA_NFFT)(InvParams& p, cudaStream_t s) {
    constexpr size_t kMaxSmem = 227 * 1024;
    const int r = (NFFT + p.hop - 1) / p.hop;
    const long long span = (p.ola_len > p.out_len + p.trim) ? p.ola_len : p.out_len + p.trim;
    // tile of TH hops: as large as fits two CTAs per SM, but at least 4 halos long
    int TH = 1;
    while (inv_smem_bytes(p.hop, TH * 2) <= kMaxSmem / 2) TH *= 2;
    while (TH < 4 * r && inv_smem_bytes(p.hop, TH * 2) <= kMaxSmem) TH *= 2;
    while ((long long)(TH / 2) * p.hop >= span && TH > 1) TH /= 2;  // short clips
    if (inv_smem_bytes(p.hop, TH) > kMaxSmem) return cudaErrorInvalidConfiguration;
    p.tile_hops = TH;
    const size_t smem = inv_smem_bytes(p.hop, TH);
    const long long TS = (long long)TH * p.hop;
    dim3 grid((unsigned)((span + TS - 1) / TS), p.B);
    cudaError_t e = cudaFuncSetAttribute(inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    inv_kernel<<<grid, THREADS, smem, s>>>(p);
    return cudaGetLastError();
}

}  // namespace mlxa
