// Fused pad -> frame -> window -> real FFT (+ epilogue) kernel for one compiled n_fft
// (built once per size with -DMLXA_NFFT=<n_fft>).
//
// PERSISTENT: the grid is one CTA per resident slot (SM count x occupancy); each CTA loops over
// (clip, frame-tile) work items.  Window, twiddles and the packed filterbank are bulk-copied into
// shared memory once per CTA.  A tile's hop-overlapped span of samples, (tile-1)*hop + n_fft, is
// staged by ONE 1-D bulk async copy (TMA: cp.async.bulk + mbarrier) -- with two staging buffers
// the copy of the next tile is in flight while the current tile is transformed; the first/last
// tiles of a clip (where the padding rules apply) are staged by index arithmetic instead.  HBM
// therefore sees each input sample ~once instead of n_fft/hop times plus three inflated round
// trips (reference stft.py:118-130).  Groups of G lanes run the Stockham plan per frame (or per
// frame pair), exchange through a padded smem buffer with __syncwarp only, unpack the real
// spectrum and hand every bin to the epilogue in registers.
#include <cstdlib>
#include "fft_plans_list.cuh"
#include "fwd_epilogue.cuh"
#include "fft_mirror.cuh"
#include "mel_project.cuh"

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
// Threads per CTA, sized so that the exchange buffers of the resident transforms fill the SM's shared
// memory: the mel epilogue runs ONE 16-warp CTA per SM for the mid-size plans (its staging tile and the
// double-buffered input want the whole SM); the store-through epilogues (STFT, Griffin-Lim) run two
// 8-warp CTAs per SM so one CTA's global stores overlap the other's butterflies.  n_fft 2048 needs 16
// warps in one CTA either way (8.4 KB of exchange buffer per warp), n_fft 4096 runs 8 warps (64 complex
// values per lane).
constexpr int threads_for(int ep) {
    return (P::E > 32) ? 256 : ((P::G >= 32 && P::E == 32) ? 512 : ((ep == EP_MEL && P::N >= 400) ? 512 : 256));
}
static_assert(P::BUF % 2 == 0 && NFFT % 4 == 0, "smem carve-up assumes 16-byte multiples");
constexpr int NBINS = NFFT / 2 + 1;
constexpr int NUNPACK = PACK ? P::N + 1 : 0;                   // 0.5*exp(-i*pi*k/N) entries
constexpr bool TW_SMEM = (P::TW + NUNPACK) * 8 <= 20 * 1024;   // twiddles staged in smem when small
constexpr int TWP = (P::TW + 1) & ~1, TWU = (NUNPACK + 1) & ~1;  // table sizes as uploaded (even counts)

constexpr int round_up4(int v) { return (v + 3) & ~3; }

struct SmemLayout {
    int in_floats, tw_f2, xch_bytes, mel_floats;
    size_t bytes;
};
// EP_MEL: the tile's power spectra are laid out [bin][frame] over the (then dead) exchange buffers and
// projected with lanes along frames (mel_project.cuh); the bank is in ROW format.
__host__ __device__ inline SmemLayout smem_layout(int ep, int NG, int hop, int TT, int n_in_buf, int n_bands, long long n_w4, int bank_in_smem) {
    SmemLayout s;
    s.in_floats = round_up4((TT - 1) * hop + NFFT + 8);  // +8: room for a 16-byte alignment lead + tail
    s.tw_f2 = TW_SMEM ? (TWP + TWU) : 0;  // both tables padded to even counts (16-byte multiples)
    const int pt_bytes = (ep == EP_MEL) ? power_tile_rows(NBINS) * power_tile_stride(TT) * 4 : 0;
    s.xch_bytes = (NG * P::BUF * 8 > pt_bytes) ? NG * P::BUF * 8 : ((pt_bytes + 15) & ~15);
    s.mel_floats = (ep == EP_MEL && bank_in_smem) ? (int)packed_bank_words(n_bands, n_w4, -1) : 0;
    s.bytes = size_t(n_in_buf * s.in_floats + NFFT + s.mel_floats) * 4 + size_t(s.tw_f2) * 8 + size_t(s.xch_bytes) + 48;
    return s;
}

// Geometry of one work item.  The tile's samples src0 .. src0 + tile_len (clip coordinates; negative or
// >= L where the centre padding applies) sit in a staging buffer at [lead, lead + tile_len), lead chosen
// so that smem and global addresses agree mod 16 bytes.  The part of the tile that exists in the clip is
// fetched by ONE bulk async copy (16-byte units: it may start up to 3 samples before the tile and end up
// to 3 after); only the nl leading / nr trailing samples outside it -- the padding of a clip's first and
// last tiles -- are filled by index arithmetic.
struct Tile {
    int b, t0, nt, tile_len, src0, lead, bulk_lo, n_bulk, nl, nr;
    const float* yb;
};
MLXA_D Tile tile_at(const FwdParams& p, int TT, int b, int tile) {
    Tile t;
    t.b = b;
    t.t0 = tile * TT;
    t.nt = min(TT, p.T - t.t0);
    t.tile_len = (t.nt - 1) * p.hop + NFFT;
    t.yb = p.y + (long long)t.b * p.ldy;
    t.src0 = t.t0 * p.hop - p.pad;
    const int a0 = int((reinterpret_cast<uintptr_t>(t.yb) >> 2) & 3);
    t.lead = (a0 + t.src0) & 3;
    const int e = t.src0 + t.tile_len;
    int lo = t.src0 - t.lead;                    // aligned at or below the first sample ...
    if (lo < 0) lo = (-a0) & 3;                  // ... or the first aligned sample of the clip
    int hi = e + ((-(a0 + e)) & 3);              // aligned at or above the end ...
    if (hi > p.L) hi = p.L - ((a0 + p.L) & 3);   // ... or the last aligned position inside the clip
    if (hi <= lo || (reinterpret_cast<uintptr_t>(t.yb) & 3)) {
        t.n_bulk = 0; t.bulk_lo = 0; t.nl = t.tile_len; t.nr = 0;
    } else {
        t.n_bulk = hi - lo;
        t.bulk_lo = lo - t.src0;
        t.nl = max(0, lo - t.src0);
        t.nr = max(0, e - hi);
    }
    return t;
}
// one thread: start the tile's bulk copy into staging buffer s_in
MLXA_D void tile_issue_bulk(const Tile& t, float* s_in, uint64_t* bar) {
    if (t.n_bulk > 0) {
        mbar_arrive_expect_tx(bar, t.n_bulk * 4);
        bulk_copy_g2s(s_in + t.lead + t.bulk_lo, t.yb + t.src0 + t.bulk_lo, t.n_bulk * 4, bar);
    }
}
// all threads: the samples the bulk copy does not cover (padding rules of pad_signal.metal:11-92)
template <int THREADS>
MLXA_D void tile_fill_edges(const FwdParams& p, const Tile& t, float* s_in) {
    const int n = t.nl + t.nr;
    for (int i = threadIdx.x; i < n; i += THREADS) {
        const int s = (i < t.nl) ? i : t.tile_len - t.nr + (i - t.nl);
        s_in[t.lead + s] = load_padded(t.yb, p.L, t.src0 + s, p.pad_mode);
    }
}

// A persistent CTA's walk over the (clip, tile) items blockIdx.x, blockIdx.x + gridDim.x, ...: one
// 32-bit division per kernel, then add-and-carry per step.
struct TileWalk {
    int b, tile, dq, dr, tpc;
    MLXA_D TileWalk(int tiles_per_clip) : tpc(tiles_per_clip) {
        b = int(blockIdx.x / unsigned(tpc));
        tile = int(blockIdx.x - unsigned(b) * unsigned(tpc));
        dq = int(gridDim.x / unsigned(tpc));
        dr = int(gridDim.x - unsigned(dq) * unsigned(tpc));
    }
    MLXA_D void advance() {
        b += dq;
        tile += dr;
        if (tile >= tpc) { tile -= tpc; ++b; }
    }
};

#include "fwd_mel_rows.cuh"
#include "fwd_mel_ws.cuh"

template <int EP, int PW>
// 16 warps per SM (register cap 128), except the 64-values-per-lane plan: one 8-warp CTA, 255 registers
__global__ void __launch_bounds__(threads_for(EP), (P::E > 32) ? 1 : 512 / threads_for(EP)) fwd_kernel(const FwdParams p) {
    constexpr int THREADS = threads_for(EP);
    constexpr int NG = THREADS / P::G;  // transforms in flight per CTA
    // Mel epilogue of the register-unpack plans: two of the three CTA barriers per tile are split into an arrival and
    // a wait with work between them (mbarriers counted per warp).  "Exchange buffers are free" is signalled right
    // after a warp's last read of its buffer and awaited only before the powers are parked -- the whole last pass and
    // the unpack lie between; "projection done" is signalled after a warp's bands and awaited before the NEXT tile's
    // first write into the exchange buffers -- its window/sample loads and pass-0 butterflies lie between.  Only the
    // barrier between parking the powers and projecting them (every warp reads every column) stays a barrier.
#ifdef MLXA_NO_SPLIT_SYNC
    constexpr bool SPLIT_SYNC = false;
#else
    constexpr bool SPLIT_SYNC = EP == EP_MEL && PF::MODE == MODE_PACK && P::NPASS == 2 && P::G <= 32 &&
                                (P::nb(P::NPASS - 1) % P::G == 0) && P::rounds(1) <= 2;  // = REG_UNPACK below
#endif
    // (the second of the two, "projection done", does not depend on how the spectrum is unpacked: every packed plan has it)
#ifdef MLXA_NO_SPLIT_SYNC
    constexpr bool SPLIT_C = false;
#else
    constexpr bool SPLIT_C = EP == EP_MEL && PF::MODE == MODE_PACK;
#endif
    // Store-through and feature epilogues: lane groups never share an exchange buffer, so the only CTA-wide hazard is
    // the staging buffer being refilled while a slow warp still reads the previous tile from it.  That is the classic
    // full / empty pair, with a ticket for "empty": after its last pass-0 load of a tile every warp draws a ticket from
    // a shared counter, and the warp that draws the tile's LAST ticket -- all warps are done with the buffer -- issues
    // the bulk copy that refills it (the next tile with one staging buffer, the tile after next with two).  Nobody
    // waits: no CTA barrier per tile, and no thread parked on an mbarrier until the slowest warp arrives.
#ifdef MLXA_NO_EMPTY_SYNC
    constexpr bool EMPTY_SYNC = false;
#else
    // Measured (profiles/r04u_*, r05e_*): against the per-tile CTA barrier STFT gains 3-20 % on every plan, the
    // Griffin-Lim projection 8 %, the feature kernels 8 %.  (A first form -- per-warp arrivals on an mbarrier, thread 0
    // waiting for them before issuing the copy -- lost 20 % on the feature kernels and 5 % at n_fft 4096: the issuing
    // warp falls behind by its wait every tile and becomes the slowest arrival of the next one.)
    constexpr bool EMPTY_SYNC = EP != EP_MEL;
#endif
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int TT = p.tile_frames;
    const int nbuf = p.n_in_buf;
    const SmemLayout lay = smem_layout(EP, NG, p.hop, TT, nbuf, p.n_bands, p.n_w4, p.bank_in_smem);

    float* s_in0 = reinterpret_cast<float*>(smem_raw);
    float* s_win = s_in0 + nbuf * lay.in_floats;
    float2* s_tw = reinterpret_cast<float2*>(s_win + NFFT);
    float2* s_buf = s_tw + lay.tw_f2;
    float* s_pw = reinterpret_cast<float*>(s_buf);  // EP_MEL: power tile [bins + 3][TT + 2] over the exchange buffers
    float* s_mel = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(s_buf) + lay.xch_bytes);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_mel + lay.mel_floats);  // [0],[1]: tile buffers, [2]: constants
    __shared__ float s_red[THREADS / 32];

    const int tiles_per_clip = (p.T + TT - 1) / TT;
    TileWalk cur(tiles_per_clip);
    if (cur.b >= p.B) return;

    // ---- once per CTA: barriers, constants, first tile -----------------------------------------
    const bool cbulk = p.const_bulk != 0;
    const uint32_t mel_bytes = (EP == EP_MEL) ? uint32_t(lay.mel_floats) * 4u : 0u;
    if (threadIdx.x == 0) {
        mbar_init(s_bar + 0, 1);
        mbar_init(s_bar + 1, 1);
        mbar_init(s_bar + 2, 1);
        if constexpr (SPLIT_SYNC) mbar_init(s_bar + 3, THREADS / 32);  // "this warp has read its exchange buffer for the last time" (this tile)
        if constexpr (SPLIT_C) mbar_init(s_bar + 4, THREADS / 32);     // "this warp has projected its share of the tile"
        if constexpr (EMPTY_SYNC) {  // tickets drawn so far, per staging buffer
            reinterpret_cast<int*>(s_bar + 3)[0] = 0;
            reinterpret_cast<int*>(s_bar + 3)[1] = 0;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(s_bar + 2, (TW_SMEM ? (TWP + TWU) * 8 : 0) + (cbulk ? NFFT * 4 + mel_bytes : 0));
        if constexpr (TW_SMEM) {
            if (TWP) bulk_copy_g2s(s_tw, p.tw_plan, TWP * 8, s_bar + 2);
            if (TWU) bulk_copy_g2s(s_tw + TWP, p.tw_unpack, TWU * 8, s_bar + 2);
        }
        if (cbulk) {
            bulk_copy_g2s(s_win, p.window, NFFT * 4, s_bar + 2);
            if (EP == EP_MEL && mel_bytes) bulk_copy_g2s(s_mel, p.bank, mel_bytes, s_bar + 2);
        }
        tile_issue_bulk(tile_at(p, TT, cur.b, cur.tile), s_in0, s_bar + 0);
    }
    if (!cbulk) {
        for (int i = threadIdx.x; i < NFFT; i += THREADS) s_win[i] = __ldg(p.window + i);
        if (EP == EP_MEL)
            for (int i = threadIdx.x; i < lay.mel_floats; i += THREADS) s_mel[i] = __ldg(p.bank + i);
    }
    const float2* tw_plan = TW_SMEM ? s_tw : p.tw_plan;
    const float2* tw_unpack = TW_SMEM ? s_tw + TWP : p.tw_unpack;
    // (two carvings: a pointer that may be shared OR global would make every weight read a generic load)
    [[maybe_unused]] RowBank rb_smem{}, rb_glob{};
    [[maybe_unused]] DbConst dbc{};
    if constexpr (EP == EP_MEL) {
        rb_smem = row_bank_carve(s_mel, p.n_w4);
        rb_glob = row_bank_carve(p.bank, p.n_w4);
        dbc = db_constants(p.db_coef, p.db_amin, p.db_ref);  // loop-invariant: one precise log2 per thread
    }
    __syncthreads();
    mbar_wait(s_bar + 2, 0);

    const int gi = threadIdx.x / P::G, g = threadIdx.x % P::G;
    float2* buf = s_buf + gi * P::BUF;
    const int PS = power_tile_stride(TT);
    // The lane's bins are k = g + Q*G: 0.5 exp(-i pi k / N) = (0.5 exp(-i pi g / N)) * exp(-i pi Q G / N) -- one table read
    // per kernel and a product with compile-time constants per bin instead of one table read per bin.  Taken where
    // it measured faster (profiles/r03a_*): the store-through epilogues (c1, c5) and the plans whose tables do not
    // fit shared memory (n_fft 4096: c4 -4.6 %); the mel / feature kernels with resident tables sit at their
    // register limit and lose to the two extra registers (c3 +0.8 %).
    constexpr bool UNPACK_IMM = PACK && ((EP != EP_MEL && EP != EP_FEAT) || !TW_SMEM);
    [[maybe_unused]] const float2 tw_base = UNPACK_IMM ? tw_unpack[g] : make_float2(0.f, 0.f);
    uint32_t ph0 = 0u, ph1 = 0u;  // parity of the next completion on each staging barrier
    [[maybe_unused]] uint32_t ph_free = 0u, ph_proj = 0u;  // ... and on the two split barriers
    float vmax = 0.f;

    for (int it = 0; cur.b < p.B; ++it, cur.advance()) {
        const int c = (nbuf == 2) ? (it & 1) : 0;
        float* s_in = s_in0 + c * lay.in_floats;
        const Tile ti = tile_at(p, TT, cur.b, cur.tile);

        // ---- stage: prefetch the next tile (two buffers) / fetch this one (one buffer) -------
        if (threadIdx.x == 0) {
            if (nbuf == 2) {
                TileWalk nxt = cur;
                nxt.advance();
                // (ticket form: only tile 1 is fetched from here; tile it + 2 is fetched by the warp that finishes tile it last)
                if (nxt.b < p.B && (!EMPTY_SYNC || it == 0))
                    tile_issue_bulk(tile_at(p, TT, nxt.b, nxt.tile), s_in0 + (c ^ 1) * lay.in_floats, s_bar + (c ^ 1));
            }  // one buffer: the copy was started as soon as the previous tile's samples had been read (below)
        }
        if (ti.nl + ti.nr) {  // CTA-uniform: a clip's first / last tile
            if constexpr (EMPTY_SYNC) __syncthreads();  // (no warp may still be reading an earlier tile from this buffer)
            tile_fill_edges<THREADS>(p, ti, s_in);
            __syncthreads();
        }
        if (ti.n_bulk > 0) {
            mbar_wait(s_bar + c, c ? ph1 : ph0);
            if (c) ph1 ^= 1u; else ph0 ^= 1u;
        }
        const int in_off = ti.lead;
        const bool even_off = (((p.hop | in_off) & 1) == 0);
        const float* tile = s_in + in_off;
        const int b = ti.b, t0 = ti.t0, nt = ti.nt;

        for (int base = 0; base < nt; base += NG * FPT) {
            const int f0 = base + gi * FPT;
            const bool va = f0 < nt;  // frames beyond the tile: finite dummy input, results unused
            float2 v[P::E];
            [[maybe_unused]] bool pair_zero_a = false, pair_zero_b = false;
            if constexpr (EP == EP_GL) {  // the projection's target magnitudes: pull the frame's row(s) towards the SM now
                if (va) {
                    const char* mp = reinterpret_cast<const char*>(p.mag + ((long long)b * p.T + t0 + f0) * p.F);
                    for (int off = g * 128; off < FPT * p.F * 4 + 128; off += P::G * 128) prefetch_l2(mp + off);
                }
            }

            // ---- pass 0: windowed samples straight from the staged tile --------------------
            if constexpr (PACK) {
                const float* src = tile + (va ? f0 * p.hop : 0);
                const bool zero = (EP == EP_GL) && (t0 + f0 >= p.T_valid);  // Griffin-Lim frame padding
                if (even_off) {
                    pass_load_fn<P, 0>(g, v, [&](int n) {
                        const float2 x = *reinterpret_cast<const float2*>(src + 2 * n);
                        const float2 w = *reinterpret_cast<const float2*>(s_win + 2 * n);
                        if constexpr (EP == EP_GL) return zero ? make_float2(0.f, 0.f) : pmul(x, w.x, w.y);
                        else return pmul(x, w.x, w.y);
                    });
                } else {
                    pass_load_fn<P, 0>(g, v, [&](int n) {
                        const float2 w = *reinterpret_cast<const float2*>(s_win + 2 * n);
                        if constexpr (EP == EP_GL) return zero ? make_float2(0.f, 0.f) : pmul(make_float2(src[2 * n], src[2 * n + 1]), w.x, w.y);
                        else return pmul(make_float2(src[2 * n], src[2 * n + 1]), w.x, w.y);
                    });
                }
            } else {
                // an absent second frame rides as a copy of the first: finite, and never stored
                const bool vb = f0 + 1 < nt;
                const float* sa = tile + (va ? f0 * p.hop : 0);
                const float* sb = tile + (vb ? (f0 + 1) * p.hop : (va ? f0 * p.hop : 0));
                const bool za = (EP == EP_GL) && (t0 + f0 >= p.T_valid), zb = (EP == EP_GL) && (t0 + f0 + 1 >= p.T_valid);
                pass_load_fn<P, 0>(g, v, [&](int n) {
                    const float w = s_win[n];
                    if constexpr (EP == EP_GL) return make_float2(za ? 0.f : sa[n] * w, zb ? 0.f : sb[n] * w);
                    else return cscale(make_float2(sa[n], sb[n]), w);
                });
                // A silent frame riding with a loud one would pick up ~1e-7 of its partner's spectrum through
                // the pair unpack; its true spectrum is exactly zero (as the reference returns), so remember
                // which of the two frames is all-zero after windowing and store zeros for it.
                unsigned ora = 0u, orb = 0u;
                static_for<P::regs(0)>([&](auto i) {
                    constexpr int I = decltype(i)::value;
                    ora |= __float_as_uint(v[I].x);
                    orb |= __float_as_uint(v[I].y);
                });
                const unsigned gm = (P::G == 32) ? 0xffffffffu : (((1u << (P::G & 31)) - 1u) << (threadIdx.x & 31 & ~(P::G - 1)));
                pair_zero_a = __all_sync(gm, (ora << 1) == 0u);
                pair_zero_b = __all_sync(gm, (orb << 1) == 0u);
            }
            pass_compute<P, 0>(g, v, tw_plan);
            if constexpr (SPLIT_C) {
                if (it > 0 && base == 0) {  // the previous tile's power tile (over the exchange buffers) has been projected by every warp
                    mbar_wait(s_bar + 4, ph_proj);
                    ph_proj ^= 1u;
                }
            }
            pass_store_buf<P, 0>(g, v, buf);
            if constexpr (EMPTY_SYNC) {
                if (base + NG * FPT >= nt) {  // last round of the tile: this warp is done with the staged samples
                    __syncwarp();
                    if ((threadIdx.x & 31) == 0) {
                        __threadfence_block();
                        // One counter per buffer; the k-th tile staged in a buffer owns tickets (THREADS / 32) * k ... + THREADS / 32 - 1.
                        // The (k + 1)-th tile is only fetched once those are gone, so draws of successive tiles never mix.
                        const int use = (nbuf == 2) ? (it >> 1) : it;
                        const int tk = atomicAdd(reinterpret_cast<int*>(s_bar + 3) + c, 1);
                        if (tk == (THREADS / 32) * (use + 1) - 1) {
                            TileWalk nxt = cur;
                            nxt.advance();
                            if (nbuf == 2) nxt.advance();
                            if (nxt.b < p.B) tile_issue_bulk(tile_at(p, TT, nxt.b, nxt.tile), s_in, s_bar + c);
                        }
                    }
                }
                group_sync<P::G>(gi);
            } else if (nbuf == 1 && EP != EP_MEL && base + NG * FPT >= nt) {
                // (the CTA-barrier form of the above)
                __syncthreads();
                if (threadIdx.x == 0) {
                    TileWalk nxt = cur;
                    nxt.advance();
                    if (nxt.b < p.B) tile_issue_bulk(tile_at(p, TT, nxt.b, nxt.tile), s_in0, s_bar + 0);
                }
            } else {
                group_sync<P::G>(gi);
            }
            pass_load_buf<P, 1>(g, v, buf);
            group_sync<P::G>(gi);
            if constexpr (SPLIT_SYNC) {
                if ((threadIdx.x & 31) == 0) mbar_arrive_one(s_bar + 3);
            }
            // Griffin-Lim: the first GL_EARLY target magnitudes of the lane go out before the last pass' butterflies, so
            // their (L2) latency hides under the arithmetic; the rest follow in the epilogue (registers: E + NQ would
            // not fit beside the transform)
            constexpr int NQ_ = ceil_div(NBINS, P::G);
            constexpr int GL_EARLY = (EP == EP_GL && PACK && P::NPASS == 2) ? (NQ_ / 2) : 0;
            [[maybe_unused]] float mg[EP == EP_GL ? NQ_ : 1];
            if constexpr (GL_EARLY > 0) {
                const float* mrow = p.mag + ((long long)b * p.T + t0 + (va ? f0 : 0)) * p.F + g;
                static_for<GL_EARLY>([&](auto q) { mg[decltype(q)::value] = __ldg(mrow + decltype(q)::value * P::G); });
            }
            pass_compute<P, 1>(g, v, tw_plan);
            if constexpr (P::NPASS == 3) {
                pass_store_buf<P, 1>(g, v, buf);
                group_sync<P::G>(gi);
                pass_load_buf<P, 2>(g, v, buf);
                group_sync<P::G>(gi);
                pass_compute<P, 2>(g, v, tw_plan);
            }
            // ---- unpack the real spectrum, feed the epilogue --------------------------------
            // packed: X[k] = 0.5*(E + w^k O), E = Z[k] + conj Z[N-k], O = -i (Z[k] - conj Z[N-k]);
            //         the table holds 0.5*w^k.   pair: Xa = (Z[k] + conj Z[N-k])/2, Xb = -i (Z[k] - conj Z[N-k])/2.
            // REG_UNPACK (two-pass packed plans): Z never goes back to shared memory.  After the last pass
            // (radix R1 over NB = R0 butterflies, lane g runs b = g + G*rd) register [rd][pos(kk)] holds
            // Z[b + R0*kk]; its Hermitian partner Z[N - b - R0*kk] is leg R1-1-kk of butterfly R0 - b, i.e. of
            // lane G - g, round RD-1-rd: one warp shuffle per component instead of a natural-order store and
            // two loads (a shuffle costs one shared-memory wavefront, tools/probes/shfl_probe.cu; the round
            // trip cost 4 + conflicts).  Lane 0 is its own mirror (leg (R1-kk)%R1 in round 0, R1-1-kk after).
            // Only the mel epilogue takes this path: it is bound by shared-memory wavefronts; the store-through
            // epilogues are bound by their global stores and lose more to the longer register lifetimes
            // (Griffin-Lim c5: 23.6 ms -> 26.0 ms with it).
            constexpr bool REG_UNPACK = (EP == EP_MEL || EP == EP_FEAT) && PACK && P::NPASS == 2 && P::G <= 32 &&
                                        (P::nb(P::NPASS - 1) % P::G == 0) && P::rounds(1) <= 2;  // (lane 0's self-mirror rule covers <= 2 rounds)
            if constexpr (!REG_UNPACK) {
                pass_store_natural<P, P::NPASS - 1>(g, v, buf);  // Z[k] at buf[k]
                group_sync<P::G>(gi);
            }
            constexpr int NQ = ceil_div(NBINS, P::G);
            const unsigned gmask = (P::G == 32) ? 0xffffffffu : (((1u << (P::G & 31)) - 1u) << (threadIdx.x & 31 & ~(P::G - 1)));
            auto bin = [&](auto q, int k) {
                constexpr int Q = decltype(q)::value;
                if constexpr (REG_UNPACK) {
                    constexpr int RD = P::rounds(1), R1 = P::radix(1), N = P::N;
                    float2 zk, zm;
                    if constexpr (Q * P::G >= N) {  // the Nyquist slot k = N (lane 0 only): Z[N] = Z[0]
                        zk = zm = v[dft_pos(R1, 0)];
                    } else {
                        constexpr int rd = Q % RD, kk = Q / RD;
                        zk = v[rd * R1 + dft_pos(R1, kk)];
                        const float2 vs = v[(RD - 1 - rd) * R1 + dft_pos(R1, R1 - 1 - kk)];
                        const int src = (P::G - g) & (P::G - 1);
                        zm.x = __shfl_sync(gmask, vs.x, src, P::G);
                        zm.y = __shfl_sync(gmask, vs.y, src, P::G);
                        const float2 own = v[rd * R1 + dft_pos(R1, rd == 0 ? (R1 - kk) % R1 : R1 - 1 - kk)];
                        if (g == 0) zm = own;
                    }
                    float2 w;
                    if constexpr (UNPACK_IMM) w = mul_tw<Q * P::G, 2 * P::N>(tw_base);
                    else w = tw_unpack[k];
                    const float2 E = cadd_conj(zk, zm), D = csub_conj(zk, zm);  // O = -i D
                    return caxpy(0.5f, E, cmul(mul_neg_i(D), w));
                } else if constexpr (PACK) {
                    constexpr int N = P::N;
                    const float2 zk = buf[(Q + 1 == NQ && k == N) ? 0 : k];
                    const float2 zm = buf[(Q == 0 && k == 0) ? 0 : N - k];
                    float2 w;
                    if constexpr (UNPACK_IMM) w = mul_tw<Q * P::G, 2 * P::N>(tw_base);
                    else w = tw_unpack[k];
                    const float2 E = cadd_conj(zk, zm), D = csub_conj(zk, zm);  // O = -i D
                    return caxpy(0.5f, E, cmul(mul_neg_i(D), w));
                } else {
                    return make_float2(0.f, 0.f);
                }
            };
            if constexpr (EP == EP_MEL) {
                // |X|^p of every bin into registers, then parked in the group's own exchange buffer
                // (Z is dead by then) where the band-sparse projection reads it.
                float pw[NQ * FPT];
                static_for<NQ>([&](auto q) {
                    constexpr int Q = decltype(q)::value;
                    const int k = g + Q * P::G;
                    if (Q + 1 < NQ || k < NBINS) {
                        if constexpr (PACK) {
                            pw[Q] = spectral_power<PW>(bin(q, k), p.power);
                        } else {
                            constexpr int N = P::N;
                            const float2 zk = buf[k];
                            const float2 zm = buf[(Q == 0 && k == 0) ? 0 : N - k];
                            pw[2 * Q] = spectral_power<PW>(make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y)), p.power);
                            pw[2 * Q + 1] = spectral_power<PW>(make_float2(0.5f * (zk.y + zm.y), 0.5f * (zm.x - zk.x)), p.power);
                        }
                    }
                });
                // every group is past its unpack reads: the power tile may overwrite the exchange buffers
                if constexpr (SPLIT_SYNC) {
                    mbar_wait(s_bar + 3, ph_free);
                    ph_free ^= 1u;
                } else {
                    __syncthreads();
                }
                // ... and past its reads of the staged samples: with a single staging buffer the next tile's
                // bulk copy starts now and lands under the projection instead of stalling the next transforms
                if (nbuf == 1 && threadIdx.x == 0) {
                    TileWalk nxt = cur;
                    nxt.advance();
                    if (nxt.b < p.B) tile_issue_bulk(tile_at(p, TT, nxt.b, nxt.tile), s_in0, s_bar + 0);
                }
                if (va) {  // frame f of the tile lives in column power_tile_col(f): lanes of the projection own frames fl, fl + TT/2
                    float* col = s_pw + power_tile_col(f0, TT >> 1);
                    [[maybe_unused]] float* col1 = s_pw + power_tile_col(f0 + (FPT - 1), TT >> 1);
                    static_for<NQ>([&](auto q) {
                        constexpr int Q = decltype(q)::value;
                        const int k = g + Q * P::G;
                        if (Q + 1 < NQ || k < NBINS) {
                            if constexpr (PACK) col[k * PS] = pw[Q];
                            else { col[k * PS] = pw[2 * Q]; col1[k * PS] = pw[2 * Q + 1]; }
                        }
                    });
                }
                for (int i = threadIdx.x; i < 3 * PS; i += THREADS) s_pw[NBINS * PS + i] = 0.f;  // rows padded quads touch
                __syncthreads();
                if (p.bank_in_smem) project_power_tile<THREADS / 32, 0, false>(p, rb_smem, dbc, s_pw, TT, b, t0, nt, 1.f, threadIdx.x >> 5, vmax);
                else project_power_tile<THREADS / 32, 0, false>(p, rb_glob, dbc, s_pw, TT, b, t0, nt, 1.f, threadIdx.x >> 5, vmax);
                if constexpr (SPLIT_C) {
                    __syncwarp();
                    if ((threadIdx.x & 31) == 0) mbar_arrive_one(s_bar + 4);
                }
            } else if constexpr (EP == EP_FEAT && P::G > 32) {
                // (not reachable: the launcher refuses EP_FEAT for two-warp groups -- the reductions are warp shuffles)
            } else if constexpr (EP == EP_FEAT) {
                // |X| (|X|^power for flatness) of the lane's bins stays in registers; the statistic of the frame
                // is reduced inside the lane group and one float per frame is written
                constexpr int KIND = PW;  // EP_FEAT instantiates the kernel per statistic (the PW slot carries STAT_*)
                constexpr bool flat = KIND == STAT_FLATNESS;
                const bool flat_sq = p.feat_p1 == 2.0f, flat_abs = p.feat_p1 == 1.0f;
                auto value = [&](float2 X) {
                    const float sq = fmaf(X.x, X.x, X.y * X.y);
                    if constexpr (flat) {
                        if (flat_sq) return sq;
                    }
                    const float m = sq > 0.f ? sq * rsqrtf(sq) : 0.f;  // |X| to 2 ulp on the SFU (sqrtf is ~8 instructions per bin)
                    if constexpr (flat) return flat_abs ? m : powf(m, p.feat_p1);
                    else return m;
                };
                float sv[NQ * FPT];
                static_for<NQ>([&](auto q) {
                    constexpr int Q = decltype(q)::value;
                    const int k = g + Q * P::G;
                    const bool ok = (Q + 1 < NQ || k < NBINS);
                    if constexpr (PACK) {
                        sv[Q] = ok ? value(bin(q, ok ? k : 0)) : 0.f;
                    } else {
                        constexpr int N = P::N;
                        const float2 zk = buf[ok ? k : 0];
                        const float2 zm = buf[(!ok || k == 0) ? 0 : N - k];
                        sv[2 * Q] = (ok && !pair_zero_a) ? value(make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y))) : 0.f;
                        sv[2 * Q + 1] = (ok && !pair_zero_b) ? value(make_float2(0.5f * (zk.y + zm.y), 0.5f * (zm.x - zk.x))) : 0.f;
                    }
                });
                static_for<FPT>([&](auto j) {
                    constexpr int J = decltype(j)::value;
                    const long long frame = (long long)b * p.T + t0 + f0 + J;
                    const float r = group_spectral_stat<KIND, P::G, NQ, FPT>(p, sv + J, g, gmask, NBINS, (f0 + J < nt) ? frame : 0);
                    if (g == 0 && f0 + J < nt) p.feat_out[frame] = r;
                });
            } else if (va) {
                const long long obase = ((long long)b * p.T + t0 + f0) * p.F;
                if constexpr (PACK) {
                    if constexpr (EP == EP_GL) {  // the remaining target magnitudes of the lane in flight before the first bin needs one
                        static_for<NQ>([&](auto q) {
                            constexpr int Q = decltype(q)::value;
                            const int k = g + Q * P::G;
                            if constexpr (Q >= GL_EARLY) mg[Q] = (Q + 1 < NQ || k < NBINS) ? __ldg(p.mag + obase + k) : 0.f;
                        });
                    }
                    static_for<NQ>([&](auto q) {
                        constexpr int Q = decltype(q)::value;
                        const int k = g + Q * P::G;
                        if (Q + 1 < NQ || k < NBINS) epilogue_bin_global<EP>(p, obase + k, bin(q, k), EP == EP_GL ? mg[Q] : 0.f);
                    });
                } else {
                    constexpr int N = P::N;  // n_fft
                    const bool fb = f0 + 1 < nt;
                    static_for<NQ>([&](auto q) {
                        constexpr int Q = decltype(q)::value;
                        const int k = g + Q * P::G;
                        if (Q + 1 < NQ || k < NBINS) {
                            const float2 zk = buf[k];
                            const float2 zm = buf[(Q == 0 && k == 0) ? 0 : N - k];
                            const float2 zero = make_float2(0.f, 0.f);
                            epilogue_bin_global<EP>(p, obase + k, pair_zero_a ? zero : make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y)));
                            if (fb) epilogue_bin_global<EP>(p, obase + p.F + k, pair_zero_b ? zero : make_float2(0.5f * (zk.y + zm.y), 0.5f * (zm.x - zk.x)));
                        }
                    });
                }
            }
            group_sync<P::G>(gi);
        }

        if constexpr (!SPLIT_C && !EMPTY_SYNC) __syncthreads();  // tile done: its staging buffer and the power tile may be overwritten
    }
    if constexpr (EP == EP_MEL) {
        if (p.gmax != nullptr) block_max_to_global<THREADS>(vmax, p.gmax, s_red, p.xchg);
    }
}

template <int EP, int PW>
cudaError_t launch_one(FwdParams& p, size_t smem, cudaStream_t s) {
    constexpr int THREADS = threads_for(EP);
    cudaError_t e = cudaFuncSetAttribute(fwd_kernel<EP, PW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int dev = 0, n_sm = 0, per_sm = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fwd_kernel<EP, PW>, THREADS, smem)) != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorInvalidConfiguration;
    const long long total = (long long)p.B * ((p.T + p.tile_frames - 1) / p.tile_frames);
    const long long slots = (long long)n_sm * per_sm;
    fwd_kernel<EP, PW><<<(unsigned)(total < slots ? total : slots), THREADS, smem, s>>>(p);
    return cudaGetLastError();
}

// mel epilogue of PAIR-mode plans: the mirror-paired kernel of fwd_mel_rows.cuh (row-format bank)
template <class PL, bool PAIR>
struct MelRowsLaunch {
    static cudaError_t run(FwdParams&, cudaStream_t) { return cudaErrorInvalidConfiguration; }
};
template <class PL, int THREADS>
struct MelRowsVariant {
    using C = MelRows<PL, THREADS>;
    template <int PW>
    static cudaError_t go(FwdParams& p, size_t smem, cudaStream_t s) {
        return p.bank_in_smem ? go2<PW, true>(p, smem, s) : go2<PW, false>(p, smem, s);
    }
    template <int PW, bool BS>
    static cudaError_t go2(FwdParams& p, size_t smem, cudaStream_t s) {
        return p.n_in_buf == 0 ? go3<PW, BS, true>(p, smem, s) : go3<PW, BS, false>(p, smem, s);
    }
    template <int PW, bool BS, bool SEP>
    static cudaError_t go3(FwdParams& p, size_t smem, cudaStream_t s) {
        if (SEP) p.n_in_buf = 1;
        auto kern = mel_rows_kernel<PL, C::THREADS, PW, BS, SEP>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int dev = 0, n_sm = 0, per_sm = 0;
        if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
        if ((e = cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, C::THREADS, smem)) != cudaSuccess) return e;
        if (per_sm < 1) return cudaErrorInvalidConfiguration;
        const long long total = (long long)p.B * ((p.T + C::TT - 1) / C::TT), slots = (long long)n_sm * per_sm;
        kern<<<(unsigned)(total < slots ? total : slots), C::THREADS, smem, s>>>(p);
        return cudaGetLastError();
    }
    static cudaError_t run(FwdParams& p, cudaStream_t s) {
        constexpr size_t kMaxSmem = 227 * 1024 - 256;
        constexpr size_t kShare = (228 * 1024) / C::CTAS_PER_SM - 1024 - 64;  // an SM's shared memory split over the resident CTAs
        long long bw = packed_bank_words(p.n_bands, p.n_w4, -1);
        p.bank_in_smem = 1;
        if (C::smem_bytes(p.hop, 1, bw) > kMaxSmem) { p.bank_in_smem = 0; bw = 0; }
        if (C::smem_bytes(p.hop, 1, bw) > kMaxSmem) return cudaErrorInvalidConfiguration;
        p.tile_frames = C::TT;
        // preferred: the power tile in shared memory of its own + one staging buffer, when that still leaves the SM
        // its full set of resident CTAs (MLXA_MEL_NO_SEP=1: the aliased layout, for A/B runs)
        static const bool no_sep = getenv("MLXA_MEL_NO_SEP") != nullptr;
        if (!no_sep && C::smem_bytes_sep(p.hop, bw) <= kShare) {
            p.n_in_buf = 0;  // (marks the SEP instantiation for go2)
            const size_t smem = C::smem_bytes_sep(p.hop, bw);
            if (p.power_mode == POW_SQUARE) return go<POW_SQUARE>(p, smem, s);
            if (p.power_mode == POW_ABS) return go<POW_ABS>(p, smem, s);
            return go<POW_GENERAL>(p, smem, s);
        }
        // double-buffer the staging when it does not cost a resident CTA
        const bool fits1 = C::smem_bytes(p.hop, 1, bw) <= kShare;
        const size_t lim = fits1 ? kShare : kMaxSmem;
        p.n_in_buf = (C::smem_bytes(p.hop, 2, bw) <= lim) ? 2 : 1;
        const size_t smem = C::smem_bytes(p.hop, p.n_in_buf, bw);
        if (p.power_mode == POW_SQUARE) return go<POW_SQUARE>(p, smem, s);
        if (p.power_mode == POW_ABS) return go<POW_ABS>(p, smem, s);
        return go<POW_GENERAL>(p, smem, s);
    }
};
template <class PL>
struct MelRowsLaunch<PL, true> {
    // two barrier-phased 8-warp CTAs per SM (fwd_mel_rows.cuh); MLXA_MEL_WS=1 selects the warp-specialised kernel
    // (fwd_mel_ws.cuh) where its single-CTA shared-memory layout fits (A/B runs: it measures within 2 % of the default)
    template <int PW, bool BS>
    static cudaError_t ws_go(FwdParams& p, size_t smem, cudaStream_t s) {
        using C = MelWs<PL>;
        auto kern = mel_ws_kernel<PL, PW, BS>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int dev = 0, n_sm = 0;
        if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
        if ((e = cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
        const long long total = (long long)p.B * ((p.T + C::TT - 1) / C::TT);
        kern<<<(unsigned)(total < n_sm ? total : n_sm), C::THREADS, smem, s>>>(p);
        return cudaGetLastError();
    }
    static cudaError_t run(FwdParams& p, cudaStream_t s) {
        using C = MelWs<PL>;
        static const bool no_ws = getenv("MLXA_MEL_WS") == nullptr;
        constexpr size_t kMaxSmem = 227 * 1024 - 256;
        long long bw = packed_bank_words(p.n_bands, p.n_w4, -1);
        bool bank_smem = true;
        if (C::smem_bytes(p.hop, bw) > kMaxSmem) { bank_smem = false; bw = 0; }
        if (no_ws || C::smem_bytes(p.hop, bw) > kMaxSmem) return MelRowsVariant<PL, 256>::run(p, s);
        p.bank_in_smem = bank_smem;
        p.n_in_buf = 1;
        p.tile_frames = C::TT;
        const size_t smem = C::smem_bytes(p.hop, bw);
        auto go = [&](auto pw) {
            constexpr int PW = decltype(pw)::value;
            return bank_smem ? ws_go<PW, true>(p, smem, s) : ws_go<PW, false>(p, smem, s);
        };
        if (p.power_mode == POW_SQUARE) return go(std::integral_constant<int, POW_SQUARE>{});
        if (p.power_mode == POW_ABS) return go(std::integral_constant<int, POW_ABS>{});
        return go(std::integral_constant<int, POW_GENERAL>{});
    }
};

}  // namespace

#define MLXA_CAT2(a, b) a##b
#define MLXA_CAT(a, b) MLXA_CAT2(a, b)

cudaError_t MLXA_CAT(launch_fwd_, MLXA_NFFT)(int ep, FwdParams& p, cudaStream_t s) {
    constexpr size_t kMaxSmem = 227 * 1024 - 256;  // opt-in maximum minus the kernel's static shared memory
    if (!PACK && ep == EP_MEL) return MelRowsLaunch<P, !PACK>::run(p, s);
    const int NG = threads_for(ep) / P::G;
    p.bank_in_smem = 1;
    if (ep == EP_MEL && smem_layout(ep, NG, p.hop, 1, 1, p.n_bands, p.n_w4, 1).bytes > kMaxSmem) p.bank_in_smem = 0;
    auto bytes = [&](int TT, int nb) { return smem_layout(ep, NG, p.hop, TT, nb, p.n_bands, p.n_w4, p.bank_in_smem).bytes; };
    // one round of transforms per tile for the mel epilogue (its staging tile is [n_bands][TT+1]),
    // two rounds for the store-through epilogues
    int TT = NG * FPT * (ep == EP_MEL ? 1 : 2);
    while (TT > 1 && bytes(TT, 1) > kMaxSmem) TT >>= 1;  // huge hops: fewer frames per tile than transform slots
    if (bytes(TT, 1) > kMaxSmem) return cudaErrorInvalidConfiguration;
    // double-buffer the staging when it does not cost a resident CTA
    const int ctas1 = (int)(kMaxSmem / bytes(TT, 1));
    const int want = ctas1 > 2 ? 2 : ctas1;
    const int nbuf = (bytes(TT, 2) <= kMaxSmem && (int)(kMaxSmem / bytes(TT, 2)) >= want) ? 2 : 1;
    p.tile_frames = TT;
    p.n_in_buf = nbuf;
    int lg = 0;
    while ((1 << lg) < TT) ++lg;
    p.log2_tile = lg;
    if (ep == EP_MEL && ((1 << lg) != TT || TT < 2 || TT > kMinBlockFrames)) return cudaErrorInvalidConfiguration;
    const size_t smem = bytes(TT, nbuf);
    if (ep == EP_STFT) return launch_one<EP_STFT, POW_SQUARE>(p, smem, s);
    if (ep == EP_GL) return launch_one<EP_GL, POW_SQUARE>(p, smem, s);
    if (ep == EP_FEAT) {
        if (P::G > 32) return cudaErrorNotSupported;
        switch (p.feat_kind) {
            case STAT_CENTROID: return launch_one<EP_FEAT, STAT_CENTROID>(p, smem, s);
            case STAT_BANDWIDTH: return launch_one<EP_FEAT, STAT_BANDWIDTH>(p, smem, s);
            case STAT_ROLLOFF: return launch_one<EP_FEAT, STAT_ROLLOFF>(p, smem, s);
            default: return launch_one<EP_FEAT, STAT_FLATNESS>(p, smem, s);
        }
    }
    if constexpr (PACK) {  // (pair-mode plans never get here: their mel epilogue is mel_rows_kernel)
        if (p.power_mode == POW_SQUARE) return launch_one<EP_MEL, POW_SQUARE>(p, smem, s);
        if (p.power_mode == POW_ABS) return launch_one<EP_MEL, POW_ABS>(p, smem, s);
        return launch_one<EP_MEL, POW_GENERAL>(p, smem, s);
    } else {
        return cudaErrorInvalidConfiguration;
    }
}

// host tables: plan twiddles and the real-unpack twiddle 0.5*exp(-i*pi*k/N)
// how the mel kernel of this n_fft wants its filterbank packed: -GP = row-pair format whose GP adjacent bands (the
// bands one warp step of the projection covers: 64 / tile frames) share one pair count
template <class PL, bool PAIR>
struct BankGroup {  // packed plans: the bands one warp step of fwd_kernel<EP_MEL> covers
    static constexpr int TT = threads_for(EP_MEL) / PL::G;
    static constexpr int value = TT >= 64 ? 1 : 64 / TT;
};
template <class PL>
struct BankGroup<PL, true> {  // pair-mode plans: mel_rows_kernel, 32-frame tiles, two bands per warp step
    static constexpr int value = 2;
};
int MLXA_CAT(plan_group_, MLXA_NFFT)() { return -BankGroup<P, !PACK>::value; }
// the fused per-frame statistics reduce with warp shuffles: plans whose groups fit a warp
int MLXA_CAT(plan_fused_feature_, MLXA_NFFT)() { return P::G <= 32 ? 1 : 0; }

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
            tw_unpack_host[k] = make_float2(float(0.5 * __builtin_cos(a)), float(0.5 * __builtin_sin(a)));
        }
}

}  // namespace mlxa
