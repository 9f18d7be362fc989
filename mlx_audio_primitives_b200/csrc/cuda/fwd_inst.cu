// Fused pad -> frame -> window -> real FFT (+ epilogue) kernel for one compiled n_fft
// (built once per size with -DMLXA_NFFT=<n_fft>).
//
// One CTA = one clip x one tile of consecutive frames.  The hop-overlapped span of samples the
// tile needs, (tile-1)*hop + n_fft, is staged ONCE in shared memory -- by a single 1-D bulk
// async copy (TMA, cp.async.bulk + mbarrier) for interior tiles, by index arithmetic (the
// padding rules) for the first/last tiles of a clip -- so HBM sees each input sample ~once
// instead of n_fft/hop times plus three inflated round trips (reference stft.py:118-130).
// Groups of G lanes then run the Stockham plan per frame (or per frame pair), exchange through a
// padded smem buffer with __syncwarp only, unpack the real spectrum and hand every bin to the
// epilogue in registers.  Window, twiddles and the band-sparse filterbank are smem-resident.
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
constexpr bool PACK = (PF::MODE == MODE_PACK);
constexpr int FPT = PACK ? 1 : 2;                              // frames per transform
// warps per CTA: sized so that the exchange buffers of all resident transforms fill the SM's shared memory
// (n_fft 2048: 16 warps x 8.4 KB; n_fft 4096: 8 warps x 16.6 KB with 64 complex values per lane)
constexpr int THREADS = (P::E > 32) ? 256 : ((P::G == 32) ? 512 : 256);
constexpr int NG = THREADS / P::G;                            // transforms in flight per CTA
static_assert((NG * P::BUF) % 2 == 0 && NFFT % 4 == 0, "smem carve-up assumes 16-byte multiples");
constexpr int NBINS = NFFT / 2 + 1;
constexpr int NUNPACK = PACK ? P::N + 1 : 0;                   // exp(-i*pi*k/N) entries
constexpr bool TW_SMEM = (P::TW + NUNPACK) * 8 <= 20 * 1024;   // twiddles staged in smem when small

constexpr int TWP = (P::TW + 1) & ~1, TWU = (NUNPACK + 1) & ~1;  // table sizes as uploaded (even counts)

constexpr int round_up4(int v) { return (v + 3) & ~3; }

struct SmemLayout {
    int in_floats, tw_f2, ep_floats, mel_floats;
    size_t bytes;
};
__host__ __device__ inline SmemLayout smem_layout(int ep, int hop, int TT, int n_bands, long long n_w4) {
    SmemLayout s;
    s.in_floats = round_up4((TT - 1) * hop + NFFT + 8);  // +8: room for a 16-byte alignment lead + tail
    s.tw_f2 = TW_SMEM ? (TWP + TWU) : 0;  // both tables padded to even counts (16-byte multiples)
    s.ep_floats = (ep == EP_MEL) ? round_up4(n_bands * (TT + 1)) : 0;  // mel staging tile [n_bands][TT+1]
    s.mel_floats = (ep == EP_MEL) ? (int)packed_bank_words(n_bands, n_w4) : 0;
    s.bytes = size_t(s.in_floats + NFFT + s.ep_floats + s.mel_floats) * 4 + size_t(s.tw_f2 + NG * P::BUF) * 8 + 16;
    return s;
}

template <int EP, int PW>
__global__ void __launch_bounds__(THREADS) fwd_kernel(const FwdParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int TT = p.tile_frames;
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * TT;
    const int nt = min(TT, p.T - t0);
    const int tile_len = (nt - 1) * p.hop + NFFT;
    const SmemLayout lay = smem_layout(EP, p.hop, TT, p.n_bands, p.n_w4);

    float* s_in = reinterpret_cast<float*>(smem_raw);
    float* s_win = s_in + lay.in_floats;
    float2* s_tw = reinterpret_cast<float2*>(s_win + NFFT);
    float2* s_buf = s_tw + lay.tw_f2;
    float* s_ep = reinterpret_cast<float*>(s_buf + NG * P::BUF);
    float* s_mel = s_ep + lay.ep_floats;
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_mel + lay.mel_floats);
    __shared__ float s_red[THREADS / 32];

    // ---- stage the tile's samples and the constants ------------------------------------------
    const float* yb = p.y + (long long)b * p.ldy;
    const int src0 = t0 * p.hop - p.pad;
    // interior tile: every sample exists -> one bulk async copy from the 16-byte aligned address
    // at or below the first sample ("lead" extra floats in front)
    const int lead = int((reinterpret_cast<uintptr_t>(yb + src0) & 15) >> 2);
    const int n_bulk = round_up4(lead + tile_len);
    const bool bulk = (src0 - lead >= 0) && (src0 - lead + n_bulk <= p.L) && ((reinterpret_cast<uintptr_t>(yb) & 3) == 0);
    const int in_off = bulk ? lead : 0;
    const bool cbulk = p.const_bulk != 0;
    const uint32_t mel_bytes = (EP == EP_MEL) ? uint32_t(lay.mel_floats) * 4u : 0u;
    if (threadIdx.x == 0) mbar_init(s_bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tx = (bulk ? n_bulk * 4 : 0) + (TW_SMEM ? (TWP + TWU) * 8 : 0) + (cbulk ? NFFT * 4 + mel_bytes : 0);
        mbar_arrive_expect_tx(s_bar, tx);
        if (bulk) bulk_copy_g2s(s_in, yb + src0 - lead, n_bulk * 4, s_bar);
        if constexpr (TW_SMEM) {
            if (TWP) bulk_copy_g2s(s_tw, p.tw_plan, TWP * 8, s_bar);
            if (TWU) bulk_copy_g2s(s_tw + TWP, p.tw_unpack, TWU * 8, s_bar);
        }
        if (cbulk) {
            bulk_copy_g2s(s_win, p.window, NFFT * 4, s_bar);
            if (EP == EP_MEL) bulk_copy_g2s(s_mel, p.bank, mel_bytes, s_bar);
        }
    }
    if (!bulk) {
        for (int i = threadIdx.x; i < tile_len; i += THREADS)
            s_in[i] = load_padded(yb, p.L, src0 + i, p.pad_mode);
    }
    if (!cbulk) {
        for (int i = threadIdx.x; i < NFFT; i += THREADS) s_win[i] = __ldg(p.window + i);
        if (EP == EP_MEL)
            for (int i = threadIdx.x; i < lay.mel_floats; i += THREADS) s_mel[i] = __ldg(p.bank + i);
    }
    const float2* tw_plan = TW_SMEM ? s_tw : p.tw_plan;
    const float2* tw_unpack = TW_SMEM ? s_tw + TWP : p.tw_unpack;
    MelSmem ms{};
    const int ep_stride = TT + 1;
    if constexpr (EP == EP_MEL) {
        ms = mel_smem_carve(s_mel, p.n_bands, p.n_w4);
    }
    __syncthreads();
    mbar_wait(s_bar, 0);

    const int gi = threadIdx.x / P::G, g = threadIdx.x % P::G;
    float2* buf = s_buf + gi * P::BUF;
    const bool even_off = (((p.hop | in_off) & 1) == 0);
    const float* tile = s_in + in_off;

    for (int base = 0; base < nt; base += NG * FPT) {
        const int f0 = base + gi * FPT;
        const bool va = (f0 < nt) && (t0 + f0 < p.T_valid);
        float2 v[P::E];

        // ---- pass 0: windowed samples straight from the staged tile --------------------
        if constexpr (PACK) {
            const float* src = tile + (va ? f0 * p.hop : 0);
            if (even_off) {
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
            const float* sa = tile + (va ? f0 * p.hop : 0);
            const float* sb = tile + (vb ? (f0 + 1) * p.hop : 0);
            pass_load_fn<P, 0>(g, v, [&](int n) {
                const float w = s_win[n];
                return make_float2(va ? sa[n] * w : 0.f, vb ? sb[n] * w : 0.f);
            });
        }
        pass_compute<P, 0>(g, v, tw_plan);
        pass_store_buf<P, 0>(g, v, buf);
        __syncwarp();
        pass_load_buf<P, 1>(g, v, buf);
        __syncwarp();
        pass_compute<P, 1>(g, v, tw_plan);
        if constexpr (P::NPASS == 3) {
            pass_store_buf<P, 1>(g, v, buf);
            __syncwarp();
            pass_load_buf<P, 2>(g, v, buf);
            __syncwarp();
            pass_compute<P, 2>(g, v, tw_plan);
        }
        pass_store_natural<P, P::NPASS - 1>(g, v, buf);  // Z[k] at buf[k]
        __syncwarp();

        // ---- unpack the real spectrum, feed the epilogue --------------------------------
        constexpr int NQ = ceil_div(NBINS, P::G);
        if constexpr (EP == EP_MEL) {
            // |X|^p of every bin into registers, then parked in the group's own exchange buffer
            // (Z is dead by then) where the band-sparse projection reads it.
            float pw[NQ * FPT];
            static_for<NQ>([&](auto q) {
                constexpr int Q = decltype(q)::value;
                const int k = g + Q * P::G;
                if (Q + 1 < NQ || k < NBINS) {
                    if constexpr (PACK) {
                        constexpr int N = P::N;
                        const float2 zk = buf[(Q + 1 == NQ && k == N) ? 0 : k];
                        const float2 zm = buf[(Q == 0 && k == 0) ? 0 : N - k];
                        const float2 w = tw_unpack[k];
                        const float ex = 0.5f * (zk.x + zm.x), ey = 0.5f * (zk.y - zm.y);
                        const float ox = 0.5f * (zk.y + zm.y), oy = -0.5f * (zk.x - zm.x);
                        const float2 X = make_float2(ex + fmaf(ox, w.x, -(oy * w.y)), ey + fmaf(ox, w.y, oy * w.x));
                        pw[Q] = spectral_power<PW>(X, p.power);
                    } else {
                        constexpr int N = P::N;
                        const float2 zk = buf[k];
                        const float2 zm = buf[(Q == 0 && k == 0) ? 0 : N - k];
                        pw[2 * Q] = spectral_power<PW>(make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y)), p.power);
                        pw[2 * Q + 1] = spectral_power<PW>(make_float2(0.5f * (zk.y + zm.y), -0.5f * (zk.x - zm.x)), p.power);
                    }
                }
            });
            __syncwarp();
            float* pbuf = reinterpret_cast<float*>(buf);
            static_for<NQ>([&](auto q) {
                constexpr int Q = decltype(q)::value;
                const int k = g + Q * P::G;
                if (Q + 1 < NQ || k < NBINS) {
                    if constexpr (PACK) pbuf[k] = pw[Q];
                    else reinterpret_cast<float2*>(pbuf)[k] = make_float2(pw[2 * Q], pw[2 * Q + 1]);
                }
            });
            __syncwarp();
            if (f0 < nt) mel_project_group<P::G, FPT>(ms, p.n_bands, g, pbuf, s_ep, ep_stride, f0);
        } else if (f0 < nt) {
            const long long obase = ((long long)b * p.T + t0 + f0) * p.F;
            if constexpr (PACK) {
                constexpr int N = P::N;  // n_fft / 2
                static_for<NQ>([&](auto q) {
                    constexpr int Q = decltype(q)::value;
                    const int k = g + Q * P::G;
                    if (Q + 1 < NQ || k <= N) {
                        const float2 zk = buf[(Q + 1 == NQ && k == N) ? 0 : k];
                        const float2 zm = buf[(Q == 0 && k == 0) ? 0 : N - k];
                        const float2 w = tw_unpack[k];
                        const float ex = 0.5f * (zk.x + zm.x), ey = 0.5f * (zk.y - zm.y);
                        const float ox = 0.5f * (zk.y + zm.y), oy = -0.5f * (zk.x - zm.x);
                        const float2 X = make_float2(ex + fmaf(ox, w.x, -(oy * w.y)), ey + fmaf(ox, w.y, oy * w.x));
                        epilogue_bin_global<EP>(p, obase + k, X);
                    }
                });
            } else {
                constexpr int N = P::N;  // n_fft
                const bool fb = f0 + 1 < nt;
                static_for<NQ>([&](auto q) {
                    constexpr int Q = decltype(q)::value;
                    const int k = g + Q * P::G;
                    if (Q + 1 < NQ || k <= N / 2) {
                        const float2 zk = buf[k];
                        const float2 zm = buf[(Q == 0 && k == 0) ? 0 : N - k];
                        epilogue_bin_global<EP>(p, obase + k, make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y)));
                        if (fb) epilogue_bin_global<EP>(p, obase + p.F + k, make_float2(0.5f * (zk.y + zm.y), -0.5f * (zk.x - zm.x)));
                    }
                });
            }
        }
        __syncwarp();
    }

    if constexpr (EP == EP_MEL) {
        __syncthreads();
        mel_store_tile<THREADS>(p, b, t0, nt, s_ep, TT, s_red);
    }
}

}  // namespace

#define MLXA_CAT2(a, b) a##b
#define MLXA_CAT(a, b) MLXA_CAT2(a, b)

cudaError_t MLXA_CAT(launch_fwd_, MLXA_NFFT)(int ep, FwdParams& p, cudaStream_t s) {
    constexpr size_t kMaxSmem = 227 * 1024;
    auto bytes = [&](int TT) { return smem_layout(ep, p.hop, TT, p.n_bands, p.n_w4).bytes; };
    int TT;
    if (ep == EP_MEL) {
        TT = 32;  // lanes run along the tile's frames in the projection phase
        while (TT > 2 && bytes(TT) > kMaxSmem) TT >>= 1;
        // prefer two CTAs per SM when that is possible with a tile of >= 16 frames
        if (TT == 32 && bytes(32) > kMaxSmem / 2 && bytes(16) <= kMaxSmem / 2) TT = 16;
    } else {
        TT = 2 * NG * FPT;  // two rounds of transforms per staged tile
        while (TT > 2 && bytes(TT) > kMaxSmem / 2) TT >>= 1;
    }
    if (bytes(TT) > kMaxSmem) return cudaErrorInvalidConfiguration;
    p.tile_frames = TT;
    const size_t smem = bytes(TT);
    dim3 grid((p.T + TT - 1) / TT, p.B);
    cudaError_t e;
#define MLXA_LAUNCH(EPV, PWV)                                                                          \
    e = cudaFuncSetAttribute(fwd_kernel<EPV, PWV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return e;                                                                    \
    fwd_kernel<EPV, PWV><<<grid, THREADS, smem, s>>>(p);
    if (ep == EP_STFT) { MLXA_LAUNCH(EP_STFT, POW_SQUARE) }
    else if (ep == EP_GL) { MLXA_LAUNCH(EP_GL, POW_SQUARE) }
    else if (p.power_mode == POW_SQUARE) { MLXA_LAUNCH(EP_MEL, POW_SQUARE) }
    else if (p.power_mode == POW_ABS) { MLXA_LAUNCH(EP_MEL, POW_ABS) }
    else { MLXA_LAUNCH(EP_MEL, POW_GENERAL) }
#undef MLXA_LAUNCH
    return cudaGetLastError();
}

// host tables: plan twiddles and the real-unpack twiddle exp(-i*pi*k/N)
void MLXA_CAT(plan_tables_, MLXA_NFFT)(float2* tw_plan_host, int* n_plan, float2* tw_unpack_host, int* n_unpack) {
    *n_plan = TWP;       // even counts: the tables are bulk-copied in 16-byte units
    *n_unpack = TWU;
    if (tw_plan_host) {
        for (int i = 0; i < TWP; ++i) tw_plan_host[i] = make_float2(0.f, 0.f);
        fill_plan_twiddles<P>(tw_plan_host);
    }
    if (tw_unpack_host) for (int i = 0; i < TWU; ++i) tw_unpack_host[i] = make_float2(0.f, 0.f);
    if (tw_unpack_host && PACK)
        for (int k = 0; k <= P::N; ++k) {
            const double a = -kPi * double(k) / double(P::N);
            tw_unpack_host[k] = make_float2(float(__builtin_cos(a)), float(__builtin_sin(a)));
        }
}

}  // namespace mlxa
