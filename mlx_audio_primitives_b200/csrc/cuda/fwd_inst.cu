// Fused pad -> frame -> window -> real FFT (+ epilogue) kernel for one compiled n_fft
// (built once per size with -DMLXA_NFFT=<n_fft>).
//
// One CTA = one clip x one tile of consecutive frames.  The hop-overlapped span of samples the
// tile needs, (tile-1)*hop + n_fft, is staged ONCE in shared memory (padding is index
// arithmetic on the way in), so HBM sees each input sample ~once instead of n_fft/hop times
// plus three inflated round trips (reference stft.py:118-130).  Groups of G lanes then run the
// Stockham plan per frame (or per frame pair), exchange through a padded smem buffer with
// __syncwarp only, unpack the real spectrum and hand every bin to the epilogue in registers.
#include "fft_plans_list.cuh"
#include "fwd_epilogue.cuh"

#ifndef MLXA_NFFT
#error "compile with -DMLXA_NFFT=<n_fft>"
#endif

namespace mlxa {
namespace {  // per-translation-unit kernels: every n_fft gets its own copy

using PF = PlanFor<MLXA_NFFT>;
using P = PF::Plan;
constexpr int NFFT = MLXA_NFFT;
constexpr int FPT = (PF::MODE == MODE_PAIR) ? 2 : 1;          // frames per transform
constexpr int THREADS = (P::E > 32) ? 128 : 256;              // register-heavy plans run fewer warps
constexpr int NG = THREADS / P::G;                            // transforms in flight per CTA

constexpr int round_up4(int v) { return (v + 3) & ~3; }

template <int EP>
__global__ void __launch_bounds__(THREADS) fwd_kernel(const FwdParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int TT = p.tile_frames;
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * TT;
    const int nt = min(TT, p.T - t0);
    const int tile_len = (nt - 1) * p.hop + NFFT;

    float* s_in = reinterpret_cast<float*>(smem_raw);
    float* s_win = s_in + round_up4((TT - 1) * p.hop + NFFT);
    float2* s_buf = reinterpret_cast<float2*>(s_win + NFFT);
    float* s_ep = reinterpret_cast<float*>(s_buf + NG * P::BUF);
    __shared__ float s_red[THREADS / 32];

    // ---- stage the tile's samples and the window --------------------------------------
    {
        const float* yb = p.y + (long long)b * p.ldy;
        const int src0 = t0 * p.hop - p.pad;
        for (int i = threadIdx.x; i < tile_len; i += THREADS)
            s_in[i] = load_padded(yb, p.L, src0 + i, p.pad_mode);
        for (int i = threadIdx.x; i < NFFT; i += THREADS) s_win[i] = __ldg(p.window + i);
    }
    __syncthreads();

    const int gi = threadIdx.x / P::G, g = threadIdx.x % P::G;
    float2* buf = s_buf + gi * P::BUF;
    const int ep_stride = TT + 1;
    const bool hop_even = (p.hop & 1) == 0;

    for (int base = 0; base < nt; base += NG * FPT) {
        const int f0 = base + gi * FPT;
        const bool va = (f0 < nt) && (t0 + f0 < p.T_valid);
        float2 v[P::E];

        // ---- pass 0: windowed samples straight from the staged tile --------------------
        if constexpr (PF::MODE == MODE_PACK) {
            const float* src = s_in + (va ? f0 * p.hop : 0);
            if (hop_even) {
                pass_load_fn<P, 0>(g, v, [&](int n) {
                    const float2 x = *reinterpret_cast<const float2*>(src + 2 * n);
                    const float2 w = *reinterpret_cast<const float2*>(s_win + 2 * n);
                    return va ? make_float2(x.x * w.x, x.y * w.y) : make_float2(0.f, 0.f);
                });
            } else {
                pass_load_fn<P, 0>(g, v, [&](int n) {
                    const float2 w = *reinterpret_cast<const float2*>(s_win + 2 * n);
                    return va ? make_float2(src[2 * n] * w.x, src[2 * n + 1] * w.y) : make_float2(0.f, 0.f);
                });
            }
        } else {
            const bool vb = (f0 + 1 < nt) && (t0 + f0 + 1 < p.T_valid);
            const float* sa = s_in + (va ? f0 * p.hop : 0);
            const float* sb = s_in + (vb ? (f0 + 1) * p.hop : 0);
            pass_load_fn<P, 0>(g, v, [&](int n) {
                const float w = s_win[n];
                return make_float2(va ? sa[n] * w : 0.f, vb ? sb[n] * w : 0.f);
            });
        }
        pass_compute<P, 0>(g, v, p.tw_plan);
        pass_store_buf<P, 0>(g, v, buf);
        __syncwarp();
        pass_load_buf<P, 1>(g, v, buf);
        __syncwarp();
        pass_compute<P, 1>(g, v, p.tw_plan);
        pass_store_buf<P, 1>(g, v, buf);
        __syncwarp();
        if constexpr (P::NPASS == 3) {
            pass_load_buf<P, 2>(g, v, buf);
            __syncwarp();
            pass_compute<P, 2>(g, v, p.tw_plan);
            pass_store_buf<P, 2>(g, v, buf);
            __syncwarp();
        }

        // ---- unpack the real spectrum, feed the epilogue --------------------------------
        if constexpr (PF::MODE == MODE_PACK) {
            constexpr int N = P::N;  // n_fft / 2
            if (f0 < nt) {
                for (int k = g; k <= N; k += P::G) {
                    const float2 zk = buf[P::phys(k == N ? 0 : k)];
                    const float2 zm = buf[P::phys(k == 0 ? 0 : N - k)];
                    const float2 w = __ldg(p.tw_unpack + k);
                    const float ex = 0.5f * (zk.x + zm.x), ey = 0.5f * (zk.y - zm.y);
                    const float ox = 0.5f * (zk.y + zm.y), oy = -0.5f * (zk.x - zm.x);
                    const float2 X = make_float2(ex + fmaf(ox, w.x, -(oy * w.y)), ey + fmaf(ox, w.y, oy * w.x));
                    epilogue_bin<EP>(p, b, t0 + f0, f0, k, X, s_ep, ep_stride);
                }
            }
        } else {
            constexpr int N = P::N;  // n_fft
            if (f0 < nt) {
                const bool fb = f0 + 1 < nt;
                for (int k = g; k <= N / 2; k += P::G) {
                    const float2 zk = buf[P::phys(k)];
                    const float2 zm = buf[P::phys(k == 0 ? 0 : N - k)];
                    const float2 Xa = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
                    const float2 Xb = make_float2(0.5f * (zk.y + zm.y), -0.5f * (zk.x - zm.x));
                    epilogue_bin<EP>(p, b, t0 + f0, f0, k, Xa, s_ep, ep_stride);
                    if (fb) epilogue_bin<EP>(p, b, t0 + f0 + 1, f0 + 1, k, Xb, s_ep, ep_stride);
                }
            }
        }
        __syncwarp();
    }

    if constexpr (EP == EP_MEL) {
        __syncthreads();
        mel_phase<THREADS>(p, b, t0, nt, s_ep, TT, s_red);
    }
}

static size_t fwd_smem_bytes(int ep, int hop, int TT, int F) {
    size_t s = size_t(round_up4((TT - 1) * hop + NFFT)) * 4 + size_t(NFFT) * 4 + size_t(NG) * P::BUF * 8;
    if (ep == EP_MEL) s += size_t(F) * (TT + 1) * 4;
    return s;
}

}  // namespace

#define MLXA_CAT2(a, b) a##b
#define MLXA_CAT(a, b) MLXA_CAT2(a, b)

cudaError_t MLXA_CAT(launch_fwd_, MLXA_NFFT)(int ep, FwdParams& p, cudaStream_t s) {
    constexpr size_t kMaxSmem = 227 * 1024;
    int TT;
    if (ep == EP_MEL) {
        TT = 32;  // lanes run along the tile's frames in the projection phase
        while (TT > 1 && fwd_smem_bytes(ep, p.hop, TT, p.F) > kMaxSmem) TT >>= 1;
        // prefer two CTAs per SM when that is possible with a tile of >= 16 frames
        if (TT == 32 && fwd_smem_bytes(ep, p.hop, 32, p.F) > kMaxSmem / 2 &&
            fwd_smem_bytes(ep, p.hop, 16, p.F) <= kMaxSmem / 2)
            TT = 16;
    } else {
        TT = 2 * NG * FPT;  // two rounds of transforms per staged tile
        while (TT > 1 && fwd_smem_bytes(ep, p.hop, TT, p.F) > kMaxSmem / 2) TT >>= 1;
    }
    if (fwd_smem_bytes(ep, p.hop, TT, p.F) > kMaxSmem) return cudaErrorInvalidConfiguration;
    p.tile_frames = TT;
    const size_t smem = fwd_smem_bytes(ep, p.hop, TT, p.F);
    dim3 grid((p.T + TT - 1) / TT, p.B);
    cudaError_t e;
#define MLXA_LAUNCH(EPV)                                                                          \
    e = cudaFuncSetAttribute(fwd_kernel<EPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return e;                                                               \
    fwd_kernel<EPV><<<grid, THREADS, smem, s>>>(p);
    if (ep == EP_STFT) { MLXA_LAUNCH(EP_STFT) }
    else if (ep == EP_MEL) { MLXA_LAUNCH(EP_MEL) }
    else { MLXA_LAUNCH(EP_GL) }
#undef MLXA_LAUNCH
    return cudaGetLastError();
}

// host tables: plan twiddles and the real-unpack twiddle exp(-i*pi*k/N)
void MLXA_CAT(plan_tables_, MLXA_NFFT)(float2* tw_plan_host, int* n_plan, float2* tw_unpack_host, int* n_unpack) {
    *n_plan = P::TW;
    *n_unpack = (PF::MODE == MODE_PACK) ? P::N + 1 : 0;
    if (tw_plan_host) fill_plan_twiddles<P>(tw_plan_host);
    if (tw_unpack_host && PF::MODE == MODE_PACK)
        for (int k = 0; k <= P::N; ++k) {
            const double a = -kPi * double(k) / double(P::N);
            tw_unpack_host[k] = make_float2(float(__builtin_cos(a)), float(__builtin_sin(a)));
        }
}

}  // namespace mlxa
