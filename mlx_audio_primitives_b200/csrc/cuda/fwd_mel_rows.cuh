// Mel-spectrogram kernel for PAIR-mode plans (n_fft = 400 = 25 x 16): two real frames ride one
// N-point complex transform, and -- because only |X|^p is wanted -- the real-spectrum unpack never
// touches shared memory:
//
//   * pass 0 (radix R0, one butterfly per lane) reads the windowed samples from the staged tile.  The
//     second lane group of a warp reads its frames cyclically rotated by 16 samples: a rotation only
//     multiplies the spectrum by a unit-modulus ramp, |X| is unchanged, and the two groups then hit
//     disjoint halves of the 32 banks (every frame starts on the same bank when hop % 32 == 0).
//   * last pass (radix R1 over NB = R0 butterflies): lane g runs butterfly g AND its mirror R0 - g.
//     Output leg k of butterfly g is Z[g + R0*k]; its Hermitian partner Z[N - g - R0*k] is leg
//     R1-1-k of the mirror butterfly, so both members of every (k, N-k) pair sit in the same lane.
//     The mirror's twiddles are conj(t) * W_R1^r; the W_R1^r factor is a one-leg rotation of the DFT
//     output (free register renaming), so one table read serves both butterflies.
//   * |Z[k] +- conj Z[N-k]|^2 = 4 |Xa[k]|^2, 4 |Xb[k]|^2: the 1/4 rides on the projected sums.
//   * the powers of a whole 64-frame tile are laid out [bin][frame] over the (dead) exchange buffers;
//     the band-sparse projection then runs with lanes along FRAMES: conflict-free 64-bit reads,
//     warp-uniform weights fetched four at a time, no padding to a lane group's longest band, and the
//     (band, 64 frames) row goes straight to global memory (coalesced), with the running max and the
//     optional dB fused.  (Replaces reference mel.py:309-352.)
//
// Included by fwd_inst.cu INSIDE its per-translation-unit namespace, after Tile / tile_info.

template <class P, int THREADS_>
struct MelRows {
    static constexpr int N = P::N, G = P::G, R0 = P::R0, R1 = P::R1;
    static_assert(G == 16 && Mirror<P>::OWNERS + 1 <= G, "two groups per warp; one idle lane zeroes the pad rows");
    static constexpr int THREADS = THREADS_, NG = THREADS / G, TT = 2 * NG;  // frames per tile: 64 (512 threads) or 32
    static_assert(kMinBlockFrames % TT == 0 && TT >= 16, "tiles never straddle a block of minima");
    static constexpr int CTAS_PER_SM = 512 / THREADS;  // 16 warps per SM either way (the register file allows no more)
    static constexpr int NBINS = N / 2 + 1;
    static constexpr int PS = power_tile_stride(TT), PROWS = power_tile_rows(NBINS);
    static constexpr int TWP = (P::TW + 1) & ~1;
    static constexpr int XCH_BYTES = (NG * P::BUF * 8 > PROWS * PS * 4) ? NG * P::BUF * 8 : PROWS * PS * 4;
    static constexpr int in_floats(int hop) { return ((TT - 1) * hop + N + 8 + 3) & ~3; }
    static constexpr size_t smem_bytes(int hop, int nbuf, long long bank_words) {
        return size_t(nbuf) * in_floats(hop) * 4 + size_t(N) * 4 + size_t(TWP) * 8 + size_t(XCH_BYTES) + size_t(bank_words) * 4 + 32;
    }
    // SEP layout: one staging buffer, and the power tile in shared memory of its own (not over the exchange buffers)
    static constexpr int XCH_ONLY = NG * P::BUF * 8, PT_BYTES = (PROWS * PS * 4 + 15) & ~15;
    static constexpr size_t smem_bytes_sep(int hop, long long bank_words) {
        return size_t(in_floats(hop)) * 4 + size_t(N) * 4 + size_t(TWP) * 8 + size_t(XCH_ONLY) + size_t(PT_BYTES) + size_t(bank_words) * 4 + 32;
    }
};

// SEP: the power tile has shared memory of its own and the samples a single staging buffer.  The transforms then need
// no CTA barrier before they park their powers (the barrier that waited for the slowest warp of the tile: 8 % of
// the warp time, ncu r03d) -- a warp writes as soon as its own last pass is done -- and the next tile's bulk copy
// is issued right behind the barrier that follows the writes, so it lands under the projection.
template <class P, int THREADS_, int PW, bool BANK_SMEM, bool SEP = false>
__global__ void __launch_bounds__(THREADS_, 512 / THREADS_) mel_rows_kernel(const FwdParams p) {
    using C = MelRows<P, THREADS_>;
    constexpr int THREADS = C::THREADS, G = C::G, TT = C::TT, N = C::N, R0 = C::R0, R1 = C::R1, PS = C::PS;
    constexpr int NBINS = C::NBINS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nbuf = SEP ? 1 : p.n_in_buf;
    const int in_floats = C::in_floats(p.hop);
    const long long bank_words = BANK_SMEM ? packed_bank_words(p.n_bands, p.n_w4, -1) : 0;

    float* s_in0 = reinterpret_cast<float*>(smem_raw);
    float* s_win = s_in0 + nbuf * in_floats;
    float2* s_tw = reinterpret_cast<float2*>(s_win + N);
    unsigned char* s_x = reinterpret_cast<unsigned char*>(s_tw + C::TWP);
    // power tile [PROWS][PS]: over the exchange buffers, or (SEP) behind them
    float* s_pw = reinterpret_cast<float*>(SEP ? s_x + C::XCH_ONLY : s_x);
    float* s_bank = reinterpret_cast<float*>(s_x + (SEP ? C::XCH_ONLY + C::PT_BYTES : C::XCH_BYTES));
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_bank + bank_words);
    __shared__ float s_red[THREADS / 32];

    const int tiles_per_clip = (p.T + TT - 1) / TT;
    TileWalk cur(tiles_per_clip);
    if (cur.b >= p.B) return;

    const bool cbulk = p.const_bulk != 0;
    if (threadIdx.x == 0) {
        mbar_init(s_bar + 0, 1);
        mbar_init(s_bar + 1, 1);
        mbar_init(s_bar + 2, 1);
        if constexpr (SEP) mbar_init(s_bar + 3, THREADS / 32);  // "every warp has projected its share of the tile"
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(s_bar + 2, C::TWP * 8 + (cbulk ? N * 4 + uint32_t(bank_words) * 4u : 0u));
        bulk_copy_g2s(s_tw, p.tw_plan, C::TWP * 8, s_bar + 2);
        if (cbulk) {
            bulk_copy_g2s(s_win, p.window, N * 4, s_bar + 2);
            if (bank_words) bulk_copy_g2s(s_bank, p.bank, uint32_t(bank_words) * 4u, s_bar + 2);
        }
        tile_issue_bulk(tile_at(p, TT, cur.b, cur.tile), s_in0, s_bar + 0);
    }
    if (!cbulk) {
        for (int i = threadIdx.x; i < N; i += THREADS) s_win[i] = __ldg(p.window + i);
        for (int i = threadIdx.x; i < (int)bank_words; i += THREADS) s_bank[i] = __ldg(p.bank + i);
    }
    const RowBank rb = row_bank_carve(BANK_SMEM ? s_bank : p.bank, p.n_w4);
    const DbConst dbc = db_constants(p.db_coef, p.db_amin, p.db_ref);  // loop-invariant: one precise log2 per thread
    const float pscale = (PW == POW_SQUARE) ? 0.25f : (PW == POW_ABS ? 0.5f : exp2f(-p.power));
    __syncthreads();
    mbar_wait(s_bar + 2, 0);
    if constexpr (BANK_SMEM) {  // the 1/4 (1/2, 2^-p) of the pair transform rides on the staged weights
        for (int i = threadIdx.x; i < (int)p.n_w4; i += THREADS) s_bank[i] *= pscale;
        __syncthreads();
    }

    const int gi = threadIdx.x / G, g = threadIdx.x % G;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float2* buf = reinterpret_cast<float2*>(s_x) + gi * P::BUF;
    // pass-0 sample offsets: lane g reads element g + 16*r (+16, cyclically, in the odd group of a warp)
    const int sh = (gi & 1) ? G : 0;
    const int o_last = (gi & 1) ? -G : G * (R0 - 1);
    uint32_t ph0 = 0u, ph1 = 0u;
    float vmax = 0.f;
    MirrorTwiddles<P> mtw;  // the lane's last-pass twiddle roots, resident for the whole kernel
    mtw.load(g, s_tw);

    for (int it = 0; cur.b < p.B; ++it, cur.advance()) {
        const int c = (nbuf == 2) ? (it & 1) : 0;
        float* s_in = s_in0 + c * in_floats;
        const Tile ti = tile_at(p, TT, cur.b, cur.tile);

        if (threadIdx.x == 0) {
            if (nbuf == 2) {
                TileWalk nxt = cur;
                nxt.advance();
                if (nxt.b < p.B) tile_issue_bulk(tile_at(p, TT, nxt.b, nxt.tile), s_in0 + (c ^ 1) * in_floats, s_bar + (c ^ 1));
            } else if (it > 0 && !SEP) {
                tile_issue_bulk(ti, s_in0, s_bar + 0);
            }
        }
        if (ti.nl + ti.nr) {  // CTA-uniform: a clip's first / last tile
            tile_fill_edges<THREADS>(p, ti, s_in);
            __syncthreads();
        }
        if (ti.n_bulk > 0) {
            mbar_wait(s_bar + c, c ? ph1 : ph0);
            if (c) ph1 ^= 1u; else ph0 ^= 1u;
        }
        const float* tile = s_in + ti.lead;
        const int nt = ti.nt;

        // ---- transform of the group's frame pair (gi, gi + NG): columns 2*gi, 2*gi + 1 of the power tile ----
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
        mirror_last_pass_powers<P, PW>(g, buf, mtw, p.power, pp);
        if constexpr (!SEP) __syncthreads();  // every exchange buffer is dead: the power tile may overwrite them
        else if (it > 0) mbar_wait(s_bar + 3, (it - 1) & 1);  // every warp has read the previous tile's powers (long since)

        // ---- powers -> tile [bin][frame]; bin of leg k: g + R0*k (k < R1/2) or its mirror ---------
        if (g <= R0 / 2) {
            float2* lo = reinterpret_cast<float2*>(s_pw) + g * (PS / 2) + gi;
            float2* hi = reinterpret_cast<float2*>(s_pw) + (R0 - g) * (PS / 2) + gi;
            static_for<R1>([&](auto k_) {
                constexpr int k = decltype(k_)::value;
                if constexpr (k < R1 / 2) lo[R0 * k * (PS / 2)] = pp[k];
                else hi[R0 * (R1 - 1 - k) * (PS / 2)] = pp[k];
            });
        } else if (g == R0 / 2 + 1 && (!SEP || it == 0)) {  // (a tile of its own keeps its zero rows)
            float2* z = reinterpret_cast<float2*>(s_pw) + NBINS * (PS / 2) + gi;
            z[0] = z[PS / 2] = z[2 * (PS / 2)] = make_float2(0.f, 0.f);
        }
        __syncthreads();
        if constexpr (SEP) {  // every warp is past its reads of the staged samples: fetch the next tile under the projection
            if (threadIdx.x == 0) {
                TileWalk nxt = cur;
                nxt.advance();
                if (nxt.b < p.B) tile_issue_bulk(tile_at(p, TT, nxt.b, nxt.tile), s_in0, s_bar + 0);
            }
        }

        // ---- band-sparse projection, lanes along frames (mel_project.cuh) -----------------------------
        project_power_tile<THREADS / 32, TT, !BANK_SMEM, true>(p, rb, dbc, s_pw, TT, ti.b, ti.t0, nt, pscale, warp, vmax);
        if constexpr (SEP) {
            // no CTA barrier here: a warp goes straight on to the next tile's transforms (its exchange buffers are its
            // own) and only checks, before it parks the next powers, that every warp has arrived here
            __syncwarp();
            if (lane == 0) mbar_arrive_one(s_bar + 3);
        } else {
            __syncthreads();  // power tile consumed: the next tile's transforms may reuse the buffers
        }
    }
    if (p.gmax != nullptr) block_max_to_global<THREADS>(vmax, p.gmax, s_red, p.xchg);
}
