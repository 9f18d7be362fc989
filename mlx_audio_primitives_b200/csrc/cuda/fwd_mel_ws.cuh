// Warp-specialised mel-spectrogram kernel for PAIR-mode plans (n_fft = 400): the same mirror-paired transform
// and frame-major projection as fwd_mel_rows.cuh, but the two kinds of work run on DIFFERENT warps of one
// persistent CTA per SM and meet only through mbarriers:
//
//   * 16 transform warps (32 lane groups -> 64-frame tiles): wait for the tile's samples, pass 0, last pass +
//     powers in registers, park the powers in the [bin][frame] tile.  They never execute a CTA-wide barrier:
//     a group's exchange buffer is private to its half-warp, the power tile has its own shared memory (it no
//     longer aliases the exchange buffers), and the only waits are "samples landed" and "power tile free".
//     Left alone, the warps drift apart, so the load bursts (LSU-bound) of some overlap the butterflies
//     (FMA-pipe-bound) of others instead of all warps hammering one unit at a time (ncu r02c: FMA pipe 46 %,
//     shared-memory wavefronts 60 %, issue 50 % with two barrier-locked CTAs per SM).
//   * 4 projection warps, one per scheduler: wait for "power tile full", run the band-sparse projection of the
//     previous tile (a latency-bound chain of shared-memory round trips and MUFUs: it wants few warps with many
//     chains in flight, not every warp of the CTA at once) under the transforms of the next one, signal "free".
//   * the TMA producer is whichever transform warp is LAST to finish reading the staged samples of tile k (a
//     shared-memory ticket): it starts the bulk copy of tile k + 1 into the single staging buffer at once, so
//     the copy has the whole last pass of tile k to land.
//
// Included by fwd_inst.cu INSIDE its per-translation-unit namespace, after fwd_mel_rows.cuh.

template <class P>
struct MelWs {
    static constexpr int N = P::N, G = P::G, R0 = P::R0, R1 = P::R1;
    static_assert(G == 16 && Mirror<P>::OWNERS + 1 <= G, "two groups per warp");
    static constexpr int T_WARPS = 16, P_WARPS = 4, THREADS = 32 * (T_WARPS + P_WARPS), T_THREADS = 32 * T_WARPS;
    static constexpr int NG = T_THREADS / G, TT = 2 * NG;  // 64-frame tiles
    static_assert(TT == kMinBlockFrames, "a tile is one block of minima");
    static constexpr int NBINS = N / 2 + 1;
    static constexpr int PS = power_tile_stride(TT), PROWS = power_tile_rows(NBINS);
    static constexpr int TWP = (P::TW + 1) & ~1;
    static constexpr int XCH_BYTES = NG * P::BUF * 8, PT_BYTES = (PROWS * PS * 4 + 15) & ~15;
    static constexpr int in_floats(int hop) { return ((TT - 1) * hop + N + 8 + 3) & ~3; }
    static constexpr size_t smem_bytes(int hop, long long bank_words) {
        return size_t(in_floats(hop)) * 4 + size_t(N) * 4 + size_t(TWP) * 8 + size_t(XCH_BYTES) + size_t(PT_BYTES) + size_t(bank_words) * 4 + 64;
    }
};

MLXA_D void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// for warps that wait long (the projection warps): back off between polls so the spin does not eat issue slots
MLXA_D void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins) {
        if (spins > (1u << 22)) __trap();
        __nanosleep(32);
    }
}
MLXA_D void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

template <class P, int PW, bool BANK_SMEM>
__global__ void __launch_bounds__(MelWs<P>::THREADS, 1) mel_ws_kernel(const FwdParams p) {
    using C = MelWs<P>;
    constexpr int G = C::G, TT = C::TT, N = C::N, R0 = C::R0, R1 = C::R1, PS = C::PS, NBINS = C::NBINS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int in_floats = C::in_floats(p.hop);
    const long long bank_words = BANK_SMEM ? packed_bank_words(p.n_bands, p.n_w4, -1) : 0;

    float* s_in = reinterpret_cast<float*>(smem_raw);
    float* s_win = s_in + in_floats;
    float2* s_tw = reinterpret_cast<float2*>(s_win + N);
    unsigned char* s_x = reinterpret_cast<unsigned char*>(s_tw + C::TWP);
    float* s_pw = reinterpret_cast<float*>(s_x + C::XCH_BYTES);  // power tile [PROWS][PS]
    float* s_bank = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(s_pw) + C::PT_BYTES);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_bank + bank_words);
    uint64_t *bar_in = s_bar + 0, *bar_infree = s_bar + 1, *bar_ptfull = s_bar + 2, *bar_ptfree = s_bar + 3, *bar_const = s_bar + 4;
    __shared__ float s_red[C::THREADS / 32];
    __shared__ unsigned s_ticket;  // transform warps done with the staged samples of the current tile

    const int tiles_per_clip = (p.T + TT - 1) / TT;
    TileWalk cur(tiles_per_clip);
    if (cur.b >= p.B) return;

    const bool cbulk = p.const_bulk != 0;
    if (threadIdx.x == 0) {
        mbar_init(bar_in, 1);
        mbar_init(bar_infree, C::T_WARPS);
        mbar_init(bar_ptfull, C::T_WARPS);
        mbar_init(bar_ptfree, C::P_WARPS);
        mbar_init(bar_const, 1);
        s_ticket = 0u;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(bar_const, C::TWP * 8 + (cbulk ? N * 4 + uint32_t(bank_words) * 4u : 0u));
        bulk_copy_g2s(s_tw, p.tw_plan, C::TWP * 8, bar_const);
        if (cbulk) {
            bulk_copy_g2s(s_win, p.window, N * 4, bar_const);
            if (bank_words) bulk_copy_g2s(s_bank, p.bank, uint32_t(bank_words) * 4u, bar_const);
        }
        tile_issue_bulk(tile_at(p, TT, cur.b, cur.tile), s_in, bar_in);
    }
    if (!cbulk) {
        for (int i = threadIdx.x; i < N; i += C::THREADS) s_win[i] = __ldg(p.window + i);
        for (int i = threadIdx.x; i < (int)bank_words; i += C::THREADS) s_bank[i] = __ldg(p.bank + i);
    }
    for (int i = threadIdx.x; i < 3 * PS; i += C::THREADS) s_pw[NBINS * PS + i] = 0.f;  // rows zero-padded weight runs may touch
    const RowBank rb = row_bank_carve(BANK_SMEM ? s_bank : p.bank, p.n_w4);
    const DbConst dbc = db_constants(p.db_coef, p.db_amin, p.db_ref);
    const float pscale = (PW == POW_SQUARE) ? 0.25f : (PW == POW_ABS ? 0.5f : exp2f(-p.power));
    __syncthreads();
    mbar_wait(bar_const, 0);
    if constexpr (BANK_SMEM) {  // the 1/4 (1/2, 2^-p) of the pair transform rides on the staged weights
        for (int i = threadIdx.x; i < (int)p.n_w4; i += C::THREADS) s_bank[i] *= pscale;
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float vmax = 0.f;
    uint32_t ph_in = 0u;  // parity of the next completion of bar_in (tiles without a bulk part do not use it)

    if (warp < C::T_WARPS) {
        // ================================ transform warps ================================================
        const int gi = threadIdx.x / G, g = threadIdx.x % G;
        float2* buf = reinterpret_cast<float2*>(s_x) + gi * P::BUF;
        const int sh = (gi & 1) ? G : 0;                    // odd group of a warp: frames rotated by 16 samples
        const int o_last = (gi & 1) ? -G : G * (R0 - 1);
        MirrorTwiddles<P> mtw;  // the lane's last-pass twiddle roots, resident for the whole kernel
        mtw.load(g, s_tw);
        for (int it = 0; cur.b < p.B; ++it, cur.advance()) {
            const Tile ti = tile_at(p, TT, cur.b, cur.tile);
            if (ti.n_bulk > 0) {
                mbar_wait(bar_in, ph_in);  // landed -- and issued only after every warp had read the previous tile
                ph_in ^= 1u;
            } else if (it > 0) {
                mbar_wait(bar_infree, (it - 1) & 1);  // no bulk part: wait for the readers of the previous tile here
            }
            if (ti.nl + ti.nr) {  // CTA-uniform: a clip's first / last tile
                for (int i = threadIdx.x; i < ti.nl + ti.nr; i += C::T_THREADS) {
                    const int s = (i < ti.nl) ? i : ti.tile_len - ti.nr + (i - ti.nl);
                    s_in[ti.lead + s] = load_padded(ti.yb, p.L, ti.src0 + s, p.pad_mode);
                }
                named_bar_sync(1, C::T_THREADS);
            }
            const float* tile = s_in + ti.lead;
            const int nt = ti.nt;
            float2 pp[R1];  // (|.|^p of frame gi, of frame gi + NG) for the lane's R1 bins, times 1/pscale
            {
                const int fa = gi, fb = gi + C::NG;
                const bool va = fa < nt, vb = fb < nt;  // absent frames ride as copies: finite, never stored
                const float* sa = tile + (va ? fa * p.hop : 0) + g + sh;
                const float* sb = tile + (vb ? fb * p.hop : (va ? fa * p.hop : 0)) + g + sh;
                const float* wp = s_win + g + sh;
                mirror_pass0<P>(g, [&](auto r_) {
                    constexpr int r = decltype(r_)::value;
                    const int o = (r == R0 - 1) ? o_last : G * r;
                    return cscale(make_float2(sa[o], sb[o]), wp[o]);
                }, buf);
            }
            __syncwarp();
            if (lane == 0) {  // this warp is done with the staged samples; the last one to get here fetches the next tile
                mbar_arrive(bar_infree);
                if ((atomicAdd(&s_ticket, 1u) % C::T_WARPS) == C::T_WARPS - 1) {  // (the counter just runs on: no reset to race with)
                    TileWalk nxt = cur;
                    nxt.advance();
                    if (nxt.b < p.B) tile_issue_bulk(tile_at(p, TT, nxt.b, nxt.tile), s_in, bar_in);
                }
            }
            mirror_last_pass_powers<P, PW>(g, buf, mtw, p.power, pp);
            if (it > 0) mbar_wait(bar_ptfree, (it - 1) & 1);  // the projection of the previous tile has read the power tile
            if (g <= R0 / 2) {
                float2* lo = reinterpret_cast<float2*>(s_pw) + g * (PS / 2) + gi;
                float2* hi = reinterpret_cast<float2*>(s_pw) + (R0 - g) * (PS / 2) + gi;
                static_for<R1>([&](auto k_) {
                    constexpr int k = decltype(k_)::value;
                    if constexpr (k < R1 / 2) lo[R0 * k * (PS / 2)] = pp[k];
                    else hi[R0 * (R1 - 1 - k) * (PS / 2)] = pp[k];
                });
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_ptfull);
        }
    } else {
        // ================================ projection warps ===============================================
        const int pw = warp - C::T_WARPS;
        for (int it = 0; cur.b < p.B; ++it, cur.advance()) {
            const Tile ti = tile_at(p, TT, cur.b, cur.tile);
            mbar_wait_sleep(bar_ptfull, it & 1);
            project_power_tile<C::P_WARPS, TT, !BANK_SMEM, true>(p, rb, dbc, s_pw, TT, ti.b, ti.t0, ti.nt, pscale, pw, vmax);
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_ptfree);
        }
    }
    if (p.gmax != nullptr) block_max_to_global<C::THREADS>(vmax, p.gmax, s_red, p.xchg);
}
