// Fused irFFT -> window -> overlap-add -> normalise -> trim kernel for one compiled n_fft
// (built once per size with -DMLXA_NFFT=<n_fft>).
//
// One CTA = one clip x one tile of tile_hops*hop OUTPUT samples.  It inverse-transforms every
// frame that touches the tile (the r-1 = ceil(n_fft/hop)-1 frames before the tile are
// recomputed as a halo), adds the windowed frames into a shared-memory accumulator and writes
// each output sample once, already divided by max(sum w^2, 1e-8) and shifted by the centre
// trim.  No atomics: frames whose index differs by a multiple of r cannot overlap.  When a tile
// of NG*r hops fits (p.spaced), a round transforms NG frames that are r apart -- all groups add
// at once and a round costs ONE CTA barrier; otherwise a round takes NG consecutive frames and
// the adds run in r barrier-separated phases.  Either way the summation order is fixed
// (deterministic results).  The (B, T, n_fft) frame tensor of the reference
// (stft.py:295 -> overlap_add.metal:16) never exists.
//
// Griffin-Lim's momentum extrapolation (griffinlim.py:176-178: rebuilt = new + m*(new - tprev), then
// istft(rebuilt)) is applied in the SIGNAL domain: the inverse STFT is linear, so
// istft(new + m*(new - tprev)) = u + m*(u - u_prev) with u = istft(new), u_prev = istft(tprev).  The kernel
// writes u (for the next iteration) and y = u + m*(u - u_prev) from its normalise-and-store loop: the
// previous projection (8*F*T bytes per clip) is never read again and the per-bin extrapolation arithmetic
// disappears from the transform loop; the price is 8*L bytes per clip of signal traffic.
#include "fft_plans_list.cuh"
#include "params.cuh"

#include <cstdlib>

#ifndef MLXA_NFFT
#error "compile with -DMLXA_NFFT=<n_fft>"
#endif

namespace mlxa {
namespace {  // per-translation-unit kernels: every n_fft gets its own copy

using PF = PlanFor<MLXA_NFFT>;
using P = PF::Plan;
constexpr int NFFT = MLXA_NFFT;
constexpr bool PACK = (PF::MODE == MODE_PACK);
constexpr int FPT = PACK ? 1 : 2;
// warps per CTA: sized so that the exchange buffers of all resident transforms fill the SM's shared memory
// (n_fft 2048: 16 warps x 8.4 KB; n_fft 4096: 8 warps x 16.6 KB with 64 complex values per lane)
#ifdef MLXA_INV_THREADS
constexpr int THREADS = MLXA_INV_THREADS;
#else
// n_fft 1024 (16 lanes x 32 values): 8 groups, so that a tile of 8 * r hops (spaced rounds) fits THREE times per SM --
// the same 12 warps as two CTAs of 12 groups, but three CTAs interleave their prologue / transform / store phases
// more finely (inverse kernel 270 -> 258 us; four CTAs of 6 groups: 282 us, the halo recompute grows to 3 in 24)
constexpr int THREADS = (P::E > 32) ? 256 : ((P::G >= 32) ? 512 : ((P::G == 16 && P::E == 32) ? 128 : 256));
#endif
#ifndef MLXA_INV_CTAS
#define MLXA_INV_CTAS ((P::G == 16 && P::E == 32 && THREADS == 128) ? 3 : 2)
#endif
constexpr int NG = THREADS / P::G;
// (the unpack table is read once per thread -- entry g, the base of the compile-time multiples -- straight from
// global memory: it takes no shared memory)
constexpr int TWP = (P::TW + 1) & ~1, TWU = 0;
constexpr bool TW_SMEM = (TWP + TWU) * 8 <= 20 * 1024;

constexpr int round_up4(int v) { return (v + 3) & ~3; }
// Spaced rounds: a frame only overlaps frames of its own lane group (earlier rounds, same lanes) and of the two
// neighbouring groups, which live in the same or an adjacent warp.  So a round's overlap-add does not need the whole
// CTA behind a barrier: a warp waits until its two neighbour warps have finished the PREVIOUS round's adds (one
// mbarrier per warp and round parity: a neighbour can be at most one round ahead) and signals its own.  The summation
// order is unchanged.  Two-warp groups (G = 64) keep the CTA barrier.  -DMLXA_INV_CTA_ROUNDS: the barrier everywhere.
constexpr int NW = THREADS / 32;
#ifdef MLXA_INV_CTA_ROUNDS
constexpr bool NEIGHBOUR_ROUNDS = false;
#else
constexpr bool NEIGHBOUR_ROUNDS = (P::G <= 32) && NW >= 2;
#endif
// PACK plans: the spectrum -> packed-input step happens inside pass 0's loads (-DMLXA_INV_UNPACK_IN_PLACE: the
// earlier form, which rewrites the buffer in place and reads it back)
#ifdef MLXA_INV_UNPACK_IN_PLACE
constexpr bool FUSED_UNPACK = false;
#else
constexpr bool FUSED_UNPACK = PACK;
#endif

MLXA_D float2 load_bin(const float2* __restrict__ X, int k, bool ok) {
    return ok ? __ldg(X + k) : make_float2(0.f, 0.f);
}

// FULLF: the spectra carry all N + 1 bins the plan reads (the usual case: no per-bin range checks)
template <bool FULLF>
__global__ void __launch_bounds__(THREADS) inv_kernel(const InvParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int TS = p.tile_hops * p.hop;
    float* s_acc = reinterpret_cast<float*>(smem_raw);
    float* s_win = s_acc + round_up4(TS);
    float2* s_tw = reinterpret_cast<float2*>(s_win + NFFT);
    float2* s_buf = s_tw + (TW_SMEM ? TWP + TWU : 0);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_buf + NG * P::BUF);
    uint64_t* s_done = s_bar + 2;  // [warp][round parity]: "this warp's overlap-add of the round has landed"

    const int b = blockIdx.y;
    const long long o0 = (long long)blockIdx.x * TS;
    const bool cbulk = p.const_bulk != 0;
    if (threadIdx.x == 0) mbar_init(s_bar, 1);
    if constexpr (NEIGHBOUR_ROUNDS) {
        if (threadIdx.x < 2 * NW) mbar_init(s_done + threadIdx.x, 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(s_bar, (TW_SMEM ? (TWP + TWU) * 8 : 0) + (cbulk ? NFFT * 4 : 0));
        if constexpr (TW_SMEM) {
            if (TWP) bulk_copy_g2s(s_tw, p.tw_plan, TWP * 8, s_bar);
            if (TWU) bulk_copy_g2s(s_tw + TWP, p.tw_unpack, TWU * 8, s_bar);
        }
        if (cbulk) bulk_copy_g2s(s_win, p.window, NFFT * 4, s_bar);
    }
    for (int i = threadIdx.x; i < TS; i += THREADS) s_acc[i] = 0.f;
    if (p.u_prev != nullptr) {  // the previous inverse of this tile is read at the very end: pull its lines into L2 now
        const long long j0 = max(0LL, o0 - p.trim), j1 = min(p.out_len, o0 + TS - p.trim);
        const char* up = reinterpret_cast<const char*>(p.u_prev + (long long)b * p.ldy);
        for (long long off = j0 * 4 + threadIdx.x * 128LL; off < j1 * 4; off += THREADS * 128LL) prefetch_l2(up + off);
    }
    if (!cbulk)
        for (int i = threadIdx.x; i < NFFT; i += THREADS) s_win[i] = __ldg(p.window + i);
    const float2* tw_plan = TW_SMEM ? s_tw : p.tw_plan;
    const float2* tw_unpack = p.tw_unpack;

    // frames touching [o0, min(o0 + TS, ola_len))
    const long long o_end = min(o0 + (long long)TS, p.ola_len);
    const int r = (NFFT + p.hop - 1) / p.hop;
    const int f_lo = (o0 < NFFT) ? 0 : int((o0 - NFFT) / p.hop) + 1;
    const int f_hi = (o_end > o0) ? int(min((long long)(p.T - 1), (o_end - 1) / p.hop)) : -1;
    __syncthreads();
    mbar_wait(s_bar, 0);
    // the 1/N of the inverse transform rides on the staged window: the overlap-add is one packed FFMA per sample pair
    for (int i = threadIdx.x; i < NFFT; i += THREADS) s_win[i] *= 1.0f / float(P::N);
    __syncthreads();

    const int gi = threadIdx.x / P::G, g = threadIdx.x % P::G;
    float2* buf = s_buf + gi * P::BUF;
    [[maybe_unused]] const float2 tw_base = PACK ? __ldg(tw_unpack + g) : make_float2(0.f, 0.f);
    const long long clip = (long long)b * p.T * p.F_in;
    const bool hop_even = (p.hop & 1) == 0;

    // PACK plans: the spectrum of a group's NEXT frame is fetched by per-lane async copies
    // (cp.async, 8 bytes) straight into the group's exchange buffer as soon as the last pass has
    // read it, so the HBM latency of round r+1 hides behind the butterflies and the overlap-add of
    // round r; the previous-projection lines of the Griffin-Lim extrapolation are pulled into L2.
    constexpr int NSPEC = PACK ? P::N + 1 : 0;
    [[maybe_unused]] const int kmax_c = FULLF ? NSPEC : min(p.F_in, NSPEC);
    [[maybe_unused]] auto prefetch_frame = [&](int f) {
        if constexpr (PACK) {
            constexpr int NQ1 = ceil_div(NSPEC, P::G);
            const long long fo = clip + (long long)f * p.F_in;
            const float2* X = p.spec + fo;
            static_for<NQ1>([&](auto q) {
                constexpr int Q = decltype(q)::value;
                const int k = g + Q * P::G;
                if (Q + 1 < NQ1 || k < NSPEC) {
                    if (FULLF || k < kmax_c) cp_async8(buf + k, X + k);
                    else buf[k] = make_float2(0.f, 0.f);
                }
            });
            cp_async_commit();
        }
    };
    // frame of (round q, this group, slot `which` of a frame pair)
    const int n_f = f_hi - f_lo + 1, per_super = NG * FPT * r;
    const bool spaced = p.spaced != 0;
    auto frame_of = [&](int q, int which) {
        if (spaced) return f_lo + (q / r) * per_super + (which * NG + gi) * r + (q % r);
        return f_lo + q * NG * FPT + gi * FPT + which;
    };
    const int n_rounds = n_f <= 0 ? 0 : (spaced ? ((n_f + per_super - 1) / per_super) * r : (n_f + NG * FPT - 1) / (NG * FPT));
    const int n_phases = spaced ? 1 : r;
    if constexpr (PACK) {
        if (n_rounds > 0 && frame_of(0, 0) <= f_hi) prefetch_frame(frame_of(0, 0));
    }

    for (int q = 0; q < n_rounds; ++q) {
        const int fa = frame_of(q, 0);
        const bool va = fa <= f_hi;
        [[maybe_unused]] const int f_next = (q + 1 < n_rounds) ? frame_of(q + 1, 0) : f_hi + 1;
        float2 v[P::E];

        // ---- spectrum -> packed complex input (swapped: inverse = swap . forward . swap) ---
        if constexpr (PACK) {
            // 1) every bin X[0..N] is read from HBM exactly once and parked in buf[k] (prefetched);
            // 2) the lane that owns the pair (k, N-k) turns it into Z[k], Z[N-k] in place:
            //    Z[k] = E + iO, Z[N-k] = conj(E) + i*conj(O), E = (X[k] + conj X[N-k])/2,
            //    O = conj(w^k) (X[k] - conj X[N-k])/2.
            constexpr int N = P::N;
            static_assert(P::BUF >= N + 1, "exchange buffer must hold the Nyquist bin");
            cp_async_wait_all();
            group_sync<P::G>(gi);
            if constexpr (FUSED_UNPACK) {
                // ... and pass 0 builds each of its inputs Z[n] from X[n] and X[N-n] as it loads them: the same formula
                // holds for every n in [0, N) (E[N-k] = conj E[k], and conj(w^(N-k)) = -w^k makes O[N-k] = conj O[k]).
                // Each pair is evaluated twice (once per end) but Z never takes a round trip through shared memory.
                constexpr int R = P::radix(0), NB = P::nb(0), RD = P::rounds(0);
                static_for<RD>([&](auto rd_) {
                    constexpr int rd = decltype(rd_)::value;
                    const int bb = g + rd * P::G;
                    if ((NB % P::G == 0) || bb < NB) {
                        static_for<R>([&](auto r_) {
                            constexpr int r = decltype(r_)::value;
                            constexpr int C = rd * P::G + r * NB;
                            const int n = g + C;
                            float2 xk = buf[n], xm = buf[N - n];
                            if constexpr (C == 0) {
                                if (g == 0) { xk.y = 0.f; xm.y = 0.f; }  // c2r ignores imag of DC / Nyquist
                            }
                            const float2 w = mul_tw<C, 2 * P::N>(tw_base);  // 0.5 * exp(-i*pi*n/N)
                            const float2 e2 = cadd_conj(xk, xm);                   // 2 E
                            const float2 o = cmul_conj(csub_conj(xk, xm), w);      // O (the 1/2 rides on w)
                            v[rd * R + r] = pfma(e2.y, e2.x, 0.5f, 0.5f, make_float2(o.x, -o.y));  // swap(E + iO)
                        });
                    }
                });
            } else {
            constexpr int NQ2 = ceil_div(N / 2 + 1, P::G);
            static_for<NQ2>([&](auto q) {
                constexpr int Q = decltype(q)::value;
                const int k = g + Q * P::G;
                if (Q + 1 < NQ2 || k <= N / 2) {
                    float2 xk = buf[k], xm = buf[N - k];
                    if (Q == 0 && k == 0) { xk.y = 0.f; xm.y = 0.f; }  // c2r ignores imag of DC / Nyquist
                    // 0.5 * exp(-i*pi*k/N) (the 1/2 of O rides on it) = the lane's table entry for k = g times a compile-time constant
                    const float2 w = mul_tw<Q * P::G, 2 * P::N>(tw_base);
                    const float ex = 0.5f * (xk.x + xm.x), ey = 0.5f * (xk.y - xm.y);
                    const float2 d = make_float2(xk.x - xm.x, xk.y + xm.y);
                    const float2 o = cmul_conj(d, w);
                    buf[k] = make_float2(ey + o.x, ex - o.y);                                   // swap(E + iO)
                    if (!(Q == 0 && k == 0)) buf[N - k] = make_float2(o.x - ey, ex + o.y);      // swap(conj E + i conj O)
                }
            });
            }
        } else {
            constexpr int N = P::N;
            constexpr int NQ = ceil_div(N / 2 + 1, P::G);
            const int fb = frame_of(q, 1);
            const bool vb = fb <= f_hi;
            const long long foa = clip + (long long)(va ? fa : 0) * p.F_in, fob = clip + (long long)(vb ? fb : 0) * p.F_in;
            const float2 *Xa = p.spec + foa, *Xb = p.spec + fob;
            static_for<NQ>([&](auto q) {
                constexpr int Q = decltype(q)::value;
                const int k = g + Q * P::G;
                if (Q + 1 < NQ || k <= N / 2) {
                    float2 a = load_bin(Xa, k, va && (FULLF || k < p.F_in));
                    float2 c = load_bin(Xb, k, vb && (FULLF || k < p.F_in));
                    if ((Q == 0 && k == 0) || 2 * k == N) { a.y = 0.f; c.y = 0.f; }
                    buf[k] = make_float2(a.y + c.x, a.x - c.y);  // swap(Xa + i*Xb)
                    if (!(Q == 0 && k == 0) && 2 * k < N) buf[N - k] = make_float2(c.x - a.y, a.x + c.y);  // swap(conj Xa + i*conj Xb)
                }
            });
        }
        if constexpr (!FUSED_UNPACK) {
            group_sync<P::G>(gi);
            pass_load_fn<P, 0>(g, v, [&](int n) { return buf[n]; });
        }
        group_sync<P::G>(gi);
        pass_compute<P, 0>(g, v, tw_plan);
        pass_store_buf<P, 0>(g, v, buf);
        group_sync<P::G>(gi);
        pass_load_buf<P, 1>(g, v, buf);
        group_sync<P::G>(gi);
        if constexpr (PACK && P::NPASS == 2) {
            if (f_next <= f_hi) prefetch_frame(f_next);  // buffer is free: fetch the next round's frame
        }
        pass_compute<P, 1>(g, v, tw_plan);
        if constexpr (P::NPASS == 3) {
            pass_store_buf<P, 1>(g, v, buf);
            group_sync<P::G>(gi);
            pass_load_buf<P, 2>(g, v, buf);
            group_sync<P::G>(gi);
            if constexpr (PACK) {
                if (f_next <= f_hi) prefetch_frame(f_next);
            }
            pass_compute<P, 2>(g, v, tw_plan);
        }

        // ---- windowed overlap-add into the tile accumulator, r conflict-free phases -------
        // lane-private part of each frame: last pass leaves element n = b + k*NS in v[], i.e. samples 2n, 2n+1
        const bool nbr = NEIGHBOUR_ROUNDS && spaced;
        if (nbr && q > 0) {  // the neighbours' adds of round q - 1 (k-th use of that parity's barrier: parity k & 1)
            const int w = threadIdx.x >> 5;
            const uint32_t par = uint32_t((q - 1) >> 1) & 1u;
            mbar_wait(s_done + 2 * ((w + NW - 1) % NW) + ((q - 1) & 1), par);
            mbar_wait(s_done + 2 * ((w + 1) % NW) + ((q - 1) & 1), par);
        }
        for (int ph = 0; ph < n_phases; ++ph) {
            if constexpr (PACK) {
                if (va && (spaced || ((fa - f_lo) % r) == ph)) {
                    const int off = int((long long)fa * p.hop - o0);  // tile-local start of the frame
                    if (off >= 0 && off + NFFT <= TS && hop_even) {  // frame fully inside the tile
                        // samples 2n, 2n + 1 = (im, re) of the swapped inverse.  Taken OLA_CH elements at a time -- all window
                        // and accumulator reads of a chunk, then the packed FMAs, then the writes -- because the compiler
                        // cannot tell that the window and the accumulator never overlap and would otherwise leave every
                        // read behind the previous element's write (one exposed shared-memory latency per element).
                        float2* acc2 = reinterpret_cast<float2*>(s_acc + off);
                        const float2* w2 = reinterpret_cast<const float2*>(s_win);
                        constexpr int LP = P::NPASS - 1;
                        constexpr int R = P::radix(LP), NB = P::nb(LP), RD = P::rounds(LP), NS = P::ns(LP);
                        constexpr int OLA_CH = (R % 8 == 0) ? 8 : (R % 5 == 0) ? 5 : (R % 4 == 0) ? 4 : (R % 3 == 0) ? 3 : (R % 2 == 0) ? 2 : 1;
                        static_for<RD>([&](auto rd_) {
                            constexpr int rd = decltype(rd_)::value;
                            const int bb = g + rd * P::G;
                            if ((NB % P::G == 0) || bb < NB) {
                                const int base = (bb / NS) * NS * R + (bb % NS);
                                static_for<R / OLA_CH>([&](auto c_) {
                                    constexpr int c0 = decltype(c_)::value * OLA_CH;
                                    float2 wv[OLA_CH], av[OLA_CH];
                                    static_for<OLA_CH>([&](auto j_) {
                                        constexpr int j = decltype(j_)::value;
                                        const int n = base + dft_perm(R, c0 + j) * NS;
                                        wv[j] = w2[n];
                                        av[j] = acc2[n];
                                    });
                                    static_for<OLA_CH>([&](auto j_) {
                                        constexpr int j = decltype(j_)::value;
                                        const float2 val = v[rd * R + c0 + j];
                                        av[j] = pfma(val.y, val.x, wv[j].x, wv[j].y, av[j]);
                                    });
                                    static_for<OLA_CH>([&](auto j_) {
                                        constexpr int j = decltype(j_)::value;
                                        acc2[base + dft_perm(R, c0 + j) * NS] = av[j];
                                    });
                                });
                            }
                        });
                    } else {
                        pass_store_fn<P, P::NPASS - 1>(g, v, [&](int n, float2 val) {
                            const int q = off + 2 * n;
                            if (q >= 0 && q < TS) s_acc[q] = fmaf(s_win[2 * n], val.y, s_acc[q]);
                            if (q + 1 >= 0 && q + 1 < TS) s_acc[q + 1] = fmaf(s_win[2 * n + 1], val.x, s_acc[q + 1]);
                        });
                    }
                }
            } else {
#pragma unroll
                for (int which = 0; which < 2; ++which) {
                    const int f = frame_of(q, which);
                    if (f <= f_hi && (spaced || ((f - f_lo) % r) == ph)) {
                        const int off = int((long long)f * p.hop - o0);
                        const bool inside = off >= 0 && off + NFFT <= TS;
                        pass_store_fn<P, P::NPASS - 1>(g, v, [&](int n, float2 val) {
                            const int q = off + n;
                            if (inside || (q >= 0 && q < TS))
                                s_acc[q] = fmaf(s_win[n], which ? val.x : val.y, s_acc[q]);
                        });
                    }
                }
            }
            if (nbr) {
                __syncwarp();
                if ((threadIdx.x & 31) == 0) mbar_arrive_one(s_done + 2 * (threadIdx.x >> 5) + (q & 1));
            } else {
                __syncthreads();
            }
        }
    }
    __syncthreads();

    // ---- normalise, trim, store: the window-sum envelope of the tile is bulk-copied into the
    // (now idle) exchange buffers when it is 16-byte aligned ----------------------------------
    float* s_wss = reinterpret_cast<float*>(s_buf);
    const long long n_w = max(0LL, min((long long)TS, p.ola_len - o0));
    const bool wbulk = ((reinterpret_cast<uintptr_t>(p.wss + o0) & 15) == 0) && (n_w % 4 == 0) && n_w > 0 &&
                       (size_t)n_w * 4 <= size_t(NG) * P::BUF * 8;
    if (wbulk) {
        if (threadIdx.x == 0) {
            mbar_arrive_expect_tx(s_bar, (uint32_t)n_w * 4);
            bulk_copy_g2s(s_wss, p.wss + o0, (uint32_t)n_w * 4, s_bar);
        }
        mbar_wait(s_bar, 1);
    }
    float* yb = p.y + (long long)b * p.ldy;
    const float* upb = p.u_prev ? p.u_prev + (long long)b * p.ldy : nullptr;
    float* uob = p.u_out ? p.u_out + (long long)b * p.ldy : nullptr;
    auto one = [&](int i, float up) {  // sample i of the tile: normalise, keep u, momentum step, store
        const long long o = o0 + i, j = o - p.trim;
        float val = 0.f;
        if (o < p.ola_len) val = __fdividef(s_acc[i], fmaxf(wbulk ? s_wss[i] : __ldg(p.wss + o), 1e-8f));
        if (uob) uob[j] = val;                                    // u = istft(spec), kept for the next call
        if (upb) val = fmaf(p.momentum, val - up, val);           // y = u + m * (u - u_prev)
        yb[j] = val;
    };
    // 128-bit path: tile, trim and rows are multiples of four samples and the whole quad lies inside the output
    const bool vec = ((TS | int(p.trim & 3) | int(p.ldy & 3)) & 3) == 0 && wbulk &&
                     ((reinterpret_cast<uintptr_t>(p.y) | reinterpret_cast<uintptr_t>(p.u_prev) | reinterpret_cast<uintptr_t>(p.u_out)) & 15) == 0;
    // a tile that lies wholly inside the trimmed output (all but a clip's first and last): no per-quad range checks,
    // 32-bit indexing, every previous-inverse read of the thread in flight before the first quad is finished
    const long long jb = o0 - p.trim;
    const bool interior = vec && jb >= 0 && jb + TS <= p.out_len && o0 + TS <= p.ola_len;
    if (interior) {
        const float4* a4 = reinterpret_cast<const float4*>(s_acc);
        const float4* w4 = reinterpret_cast<const float4*>(s_wss);
        float4* y4 = reinterpret_cast<float4*>(yb + jb);
        const float4* up4 = upb ? reinterpret_cast<const float4*>(upb + jb) : nullptr;
        float4* uo4 = uob ? reinterpret_cast<float4*>(uob + jb) : nullptr;
        const int nq = TS >> 2;
        const float m = p.momentum;
        constexpr int U = 8;
        for (int i0 = threadIdx.x; i0 < nq; i0 += THREADS * U) {
            float4 up[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + u * THREADS;
                up[u] = (up4 && i < nq) ? __ldg(up4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + u * THREADS;
                if (i < nq) {
                    const float4 a = a4[i], w = w4[i];
                    float4 v = make_float4(__fdividef(a.x, fmaxf(w.x, 1e-8f)), __fdividef(a.y, fmaxf(w.y, 1e-8f)),
                                           __fdividef(a.z, fmaxf(w.z, 1e-8f)), __fdividef(a.w, fmaxf(w.w, 1e-8f)));
                    if (uo4) uo4[i] = v;
                    if (up4)
                        v = make_float4(fmaf(m, v.x - up[u].x, v.x), fmaf(m, v.y - up[u].y, v.y), fmaf(m, v.z - up[u].z, v.z),
                                        fmaf(m, v.w - up[u].w, v.w));
                    y4[i] = v;
                }
            }
        }
    } else if (vec) {
        constexpr int U = 4;  // quads per thread in flight (the u_prev reads are the only long-latency operation here)
        for (int i0 = threadIdx.x * 4; i0 < TS; i0 += THREADS * 4 * U) {
            float4 up[U];
            bool full[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + u * THREADS * 4;
                const long long j = o0 + i - p.trim;
                full[u] = i < TS && j >= 0 && j + 3 < p.out_len && o0 + i + 3 < p.ola_len;
                up[u] = (full[u] && upb) ? __ldg(reinterpret_cast<const float4*>(upb + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + u * THREADS * 4;
                if (i >= TS) continue;
                const long long j = o0 + i - p.trim;
                if (full[u]) {
                    const float4 a = *reinterpret_cast<const float4*>(s_acc + i), w = *reinterpret_cast<const float4*>(s_wss + i);
                    float4 v = make_float4(__fdividef(a.x, fmaxf(w.x, 1e-8f)), __fdividef(a.y, fmaxf(w.y, 1e-8f)),
                                           __fdividef(a.z, fmaxf(w.z, 1e-8f)), __fdividef(a.w, fmaxf(w.w, 1e-8f)));
                    if (uob) *reinterpret_cast<float4*>(uob + j) = v;
                    if (upb) {
                        const float m = p.momentum;
                        v = make_float4(fmaf(m, v.x - up[u].x, v.x), fmaf(m, v.y - up[u].y, v.y), fmaf(m, v.z - up[u].z, v.z),
                                        fmaf(m, v.w - up[u].w, v.w));
                    }
                    *reinterpret_cast<float4*>(yb + j) = v;
                } else {  // a quad that straddles the trim, the end of the signal or the end of the overlap-add
                    for (int e = 0; e < 4; ++e) {
                        const long long je = j + e;
                        if (je >= 0 && je < p.out_len) one(i + e, upb ? __ldg(upb + je) : 0.f);
                    }
                }
            }
        }
    } else {
        constexpr int U = 4;
        for (int i0 = threadIdx.x; i0 < TS; i0 += THREADS * U) {
            float up[U];
            bool ok[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + u * THREADS;
                const long long j = o0 + i - p.trim;
                ok[u] = i < TS && j >= 0 && j < p.out_len;
                up[u] = (ok[u] && upb) ? __ldg(upb + j) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (ok[u]) one(i0 + u * THREADS, up[u]);
        }
    }
}

static size_t inv_smem_bytes(int hop, int TH) {
    return size_t(round_up4(TH * hop)) * 4 + size_t(NFFT) * 4 + size_t(TW_SMEM ? TWP + TWU : 0) * 8 +
           size_t(NG) * P::BUF * 8 + 16 + size_t(2 * NW) * 8;
}

}  // namespace

#define MLXA_CAT2(a, b) a##b
#define MLXA_CAT(a, b) MLXA_CAT2(a, b)

cudaError_t MLXA_CAT(launch_inv_, MLXA_NFFT)(InvParams& p, cudaStream_t s) {
    constexpr size_t kMaxSmem = 227 * 1024;
    const int r = (NFFT + p.hop - 1) / p.hop;
    const long long span = (p.ola_len > p.out_len + p.trim) ? p.ola_len : p.out_len + p.trim;
    // A tile of TH hops needs the TH + r - 1 frames that touch it.  Frames are transformed in rounds of
    // NG*FPT, so TH is chosen as m*NG*FPT - (r - 1): every round is full (no idle transform slots) and
    // the halo recompute is (r - 1) frames per m rounds.  m grows while two CTAs still fit per SM.
    const int per_round = NG * FPT;
    p.spaced = 0;
    // spaced rounds: a tile of m * per_round * r - (r - 1) hops keeps every slot of every round busy; taken when two
    // such CTAs fit per SM and the clip is long enough to fill the tile
    {
        constexpr size_t kHalf = (228 * 1024) / MLXA_INV_CTAS - 1024;
        const int per_super = per_round * r;
        auto th_sp = [&](int m) { return m * per_super - (r - 1); };
        static const bool no_spaced = getenv("MLXA_INV_NO_SPACED") != nullptr;
        if (!no_spaced && th_sp(1) >= 1 && inv_smem_bytes(p.hop, th_sp(1)) <= kHalf && (long long)th_sp(1) * p.hop <= span) {
            int m = 1;
            while (inv_smem_bytes(p.hop, th_sp(m + 1)) <= kHalf && (long long)th_sp(m + 1) * p.hop <= span) ++m;
            p.spaced = 1;
            p.tile_hops = th_sp(m);
        }
    }
    auto th_for = [&](int m) { return m * per_round - (r - 1); };
    int m = 1;
    while (th_for(m) < 1) ++m;
    while (inv_smem_bytes(p.hop, th_for(m + 1)) <= kMaxSmem / 2) ++m;
    if (inv_smem_bytes(p.hop, th_for(m)) > kMaxSmem / 2) {  // not even one round fits twice: go for one CTA per SM
        while (inv_smem_bytes(p.hop, th_for(m + 1)) <= kMaxSmem && m < 4) ++m;
    }
    int TH = th_for(m);
    while (m > 1 && (long long)th_for(m - 1) * p.hop >= span) TH = th_for(--m);  // short clips
    if ((long long)TH * p.hop > span) {  // very short clip: one tile that just covers it
        TH = (int)((span + p.hop - 1) / p.hop);
        if (TH < 1) TH = 1;
    }
    while (TH > 1 && inv_smem_bytes(p.hop, TH) > kMaxSmem) TH = (TH + 1) / 2;  // huge hops: partial rounds
    if (inv_smem_bytes(p.hop, TH) > kMaxSmem) return cudaErrorInvalidConfiguration;
    if (p.spaced) TH = p.tile_hops;
    p.tile_hops = TH;
    const size_t smem = inv_smem_bytes(p.hop, TH);
    const long long TS = (long long)TH * p.hop;
    dim3 grid((unsigned)((span + TS - 1) / TS), p.B);
    cudaError_t e;
    constexpr int NSPEC_ALL = PACK ? P::N + 1 : P::N / 2 + 1;
    if (p.F_in >= NSPEC_ALL) {
        e = cudaFuncSetAttribute(inv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        inv_kernel<true><<<grid, THREADS, smem, s>>>(p);
    } else {
        e = cudaFuncSetAttribute(inv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        inv_kernel<false><<<grid, THREADS, smem, s>>>(p);
    }
    return cudaGetLastError();
}

}  // namespace mlxa
