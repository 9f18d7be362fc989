// Band-sparse filterbank projection of a tile of power spectra laid out [bin][column] in shared memory
// (replaces the dense matmul of reference mel.py:344): lanes run along FRAMES.  A band row of the tile is
// TT/2 lanes x 2 frames -- the lane's float2 holds frames fl and fl + TT/2 of the tile (the transforms park
// frame f in column 2*(f mod TT/2) + f / (TT/2)), so both stores of a band row are contiguous runs of TT/2
// floats.  A warp covers 64/TT ADJACENT bands per step; the ROW-PAIR bank pads the bands of such a group to
// one common count of bin quads, so the accumulation loop has a warp-uniform trip count (no divergence, no
// per-lane loop bookkeeping); a quad's four weights arrive in one 128-bit shared-memory read and enter the packed
// FFMA2 as broadcast scalar operands (no register moves).  Every (band, frame)
// value goes from registers to global memory with the running max, the per-block minimum (3-input FMNMX) and
// the optional dB conversion (one clamp, one MUFU.LG2, one FFMA) fused.
#pragma once
#include "fwd_epilogue.cuh"

namespace mlxa {

// bank packed in ROW-PAIR format (mlxa_plan_group = -GP < 0, see include/mlxa_cuda.h): per band 1 + nq entries of
// four weights ({w0, w1, -, -}, then quads of bins), then one int4 {first bin, nq, first entry, support length}
// per band, bands padded to a multiple of 32
struct RowBank {
    const float4* wt4;
    const int4* desc;
};
MLXA_D RowBank row_bank_carve(const float* base, long long n_wt) {
    RowBank r;
    r.wt4 = reinterpret_cast<const float4*>(base);
    r.desc = reinterpret_cast<const int4*>(base + n_wt);
    return r;
}

// power-tile geometry for TT frames: row stride (floats; even, == 2 mod 4 so the 64-bit frame pairs of
// consecutive rows spread over the banks) and row count (+3: rows the zero-padded weight pairs may touch)
constexpr int power_tile_stride(int TT) { return TT + 2; }
constexpr int power_tile_rows(int n_bins) { return n_bins + 3; }
// column of the power tile that holds frame f of a TT-frame tile, and the frame a column holds
MLXA_HD int power_tile_col(int f, int half) { return f < half ? 2 * f : 2 * (f - half) + 1; }

#if defined(__CUDACC__)
MLXA_D float max3(float a, float b, float c) { float r; asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
MLXA_D float min3(float a, float b, float c) { float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
#endif

// DBM: 0 raw values, 1 dB with amin / ref in the normal range (flush-to-zero MUFU.LG2, no denormal fix-up),
// 2 dB for any amin (same bits as 1 wherever 1 is valid).
template <int DBM>
MLXA_D float project_db(float v, float amin, float c1, float c0) {
    if constexpr (DBM == 0) return v;
    const float x = fmaxf(v, amin);
    float l;
    if constexpr (DBM == 1) asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(x));
    else l = __log2f(x);
    return fmaf(l, c1, c0);
}

// TTC: tile frames when known at compile time, 0 -> tt (a power of two >= 2).  SCALE: multiply the sums by
// pscale (pair transforms deliver 4|X|^2 and the bank could not be pre-scaled).
// NW warps share the tile's bands; `warp` is this warp's index among them.
// One band step per warp at a time: the form the barrier-phased kernels use (every warp of the CTA projects at
// once, long bands: the quad loop dominates and the serial chain per step is hidden by the other warps).
template <int NW, int TTC, bool SCALE, int DBM, bool FULL>
MLXA_D void project_power_tile_impl(const FwdParams& p, const RowBank& rb, const DbConst& dbc, const float* s_pw, int tt, int b,
                                    int t0, int nt, float pscale, int warp, float& vmax, float& tmin) {
    const int TT = TTC ? TTC : tt;
    const int PS = power_tile_stride(TT);
    const int lane = threadIdx.x & 31;
    const int LB = TT / 2, NBW = LB >= 32 ? 1 : 32 / LB;  // lanes per band row, bands per warp step
    const int fl = lane & (LB - 1), sub = lane / LB;
    const int per_step = NW * NBW;                 // bands the CTA covers per step
    const int n_pad = (p.n_bands + NBW - 1) / NBW * NBW;  // the bank carries descriptors (zero weights) up to here
    char* ob = reinterpret_cast<char*>(p.mel + (long long)b * p.n_bands * p.T + t0 + fl);
    const unsigned row_bytes = unsigned(p.T) * 4u;
    const float2* q_lane = reinterpret_cast<const float2*>(s_pw) + fl;
    const int rs = PS / 2;  // row stride in float2
    const bool ok0 = FULL || fl < nt, ok1 = FULL || fl + LB < nt;
    // band groups go round the warps boustrophedon, so every warp gets short and long bands alike
    const int m_even = warp * NBW + sub, m_odd = (NW - 1 - warp) * NBW + sub;
#pragma unroll 1
    for (int m0 = 0, odd = 0; m0 < n_pad; m0 += per_step, odd ^= 1) {
        const int m = m0 + (odd ? m_odd : m_even);
        if (m - sub >= n_pad) continue;  // warp-uniform: no band group left for this warp
        const int4 d = rb.desc[m];  // {first bin, quads after the first pair, first weight entry, -}
        const float4* w = rb.wt4 + d.z;
        const float2* q = q_lane + d.x * rs;
        // the first pair of bins is always there (entry 0 = {w0, w1, -, -}) ...
        const float4 w0 = w[0];
        mlxa_u64 a0 = mul2(pk2(q[0]), pk2(w0.x, w0.x));
        mlxa_u64 a1 = mul2(pk2(q[rs]), pk2(w0.y, w0.y));
        // ... then nq quads, nq the same for all bands of the warp step
        const float4* wend = w + d.y;
#pragma unroll 1
        for (; w != wend; ++w, q += 4 * rs) {
            const float4 wa = w[1];
            a0 = fma2(pk2(q[2 * rs]), pk2(wa.x, wa.x), a0);
            a1 = fma2(pk2(q[3 * rs]), pk2(wa.y, wa.y), a1);
            a0 = fma2(pk2(q[4 * rs]), pk2(wa.z, wa.z), a0);
            a1 = fma2(pk2(q[5 * rs]), pk2(wa.w, wa.w), a1);
        }
        float2 v = up2(add2(a0, a1));
        if constexpr (SCALE) v = cscale(v, pscale);
        if (m < p.n_bands) {
            float* o = reinterpret_cast<float*>(ob + (unsigned long long)unsigned(m) * row_bytes);
            if (FULL || ok1) {  // frames fill a tile from column 0 up: ok1 implies ok0
                vmax = max3(vmax, v.x, v.y);
                tmin = min3(tmin, v.x, v.y);
                o[0] = project_db<DBM>(v.x, dbc.amin, dbc.c1, dbc.c0);
                o[LB] = project_db<DBM>(v.y, dbc.amin, dbc.c1, dbc.c0);
            } else if (ok0) {
                vmax = fmaxf(vmax, v.x);
                tmin = fminf(tmin, v.x);
                o[0] = project_db<DBM>(v.x, dbc.amin, dbc.c1, dbc.c0);
            }
        }
    }
}

// The form the n_fft = 400 kernels use (short bands: one or two quads, so a band step is all fixed cost -- a
// serial chain of four shared-memory round trips and a MUFU, one instruction every ~7 cycles per warp when taken
// one after the other, ncu r02b).  A warp takes its band steps kBatch at a time as independent chains: all
// descriptors are read first, then the first weights and power pairs of every chain, then the chains' quad
// loops, then all epilogues with their MUFUs in flight together.
constexpr int kBatch = 5;
template <int NW, int TTC, bool SCALE, int DBM, bool FULL>
MLXA_D void project_power_tile_batched(const FwdParams& p, const RowBank& rb, const DbConst& dbc, const float* s_pw, int b, int t0,
                                       int nt, float pscale, int warp, float& vmax, float& tmin) {
    static_assert(TTC >= 2, "compile-time tile size");
    constexpr int TT = TTC, PS = power_tile_stride(TT), rs = PS / 2, CH = kBatch;
    constexpr int LB = TT / 2, NBW = LB >= 32 ? 1 : 32 / LB;  // lanes per band row, bands per warp step
    const int lane = threadIdx.x & 31;
    const int fl = lane & (LB - 1), sub = lane / LB;
    constexpr int per_step = NW * NBW;             // bands the CTA covers per step
    const int n_pad = (p.n_bands + NBW - 1) / NBW * NBW;  // the bank carries descriptors (zero weights) up to here
    char* ob = reinterpret_cast<char*>(p.mel + (long long)b * p.n_bands * p.T + t0 + fl);
    const unsigned row_bytes = unsigned(p.T) * 4u;
    const float2* q_lane = reinterpret_cast<const float2*>(s_pw) + fl;
    const bool ok0 = FULL || fl < nt, ok1 = FULL || fl + LB < nt;
    // band groups go round the warps boustrophedon, so every warp gets short and long bands alike
    const int m_even = warp * NBW + sub, m_odd = (NW - 1 - warp) * NBW + sub;
#pragma unroll 1
    for (int m0 = 0; m0 < n_pad; m0 += CH * per_step) {
        int m[CH];
        bool act[CH];  // warp-uniform: the warp has a band group in this slot
        int4 d[CH];    // {first bin, quads after the first pair, first weight entry, -}
        static_for<CH>([&](auto c_) {
            constexpr int c = decltype(c_)::value;
            m[c] = m0 + c * per_step + ((c & 1) ? m_odd : m_even);
            act[c] = m[c] - sub < n_pad;
            d[c] = rb.desc[act[c] ? m[c] : sub];
        });
        const float4* w[CH];
        const float2* q[CH];
        float4 w0[CH];
        float2 q0[CH], q1[CH];
        static_for<CH>([&](auto c_) {  // the pair of bins every run starts with (entry 0 = {w0, w1, -, -})
            constexpr int c = decltype(c_)::value;
            w[c] = rb.wt4 + d[c].z;
            q[c] = q_lane + d[c].x * rs;
            w0[c] = w[c][0];
            q0[c] = q[c][0];
            q1[c] = q[c][rs];
        });
        mlxa_u64 a0[CH], a1[CH];
        static_for<CH>([&](auto c_) {
            constexpr int c = decltype(c_)::value;
            a0[c] = mul2(pk2(q0[c]), pk2(w0[c].x, w0[c].x));
            a1[c] = mul2(pk2(q1[c]), pk2(w0[c].y, w0[c].y));
        });
        static_for<CH>([&](auto c_) {  // then nq quads of bins, nq the same for all bands of the warp step
            constexpr int c = decltype(c_)::value;
            const float4* wp = w[c];
            const float4* wend = wp + d[c].y;
            const float2* qp = q[c];
#pragma unroll 1
            for (; wp != wend; ++wp, qp += 4 * rs) {
                const float4 wa = wp[1];
                a0[c] = fma2(pk2(qp[2 * rs]), pk2(wa.x, wa.x), a0[c]);
                a1[c] = fma2(pk2(qp[3 * rs]), pk2(wa.y, wa.y), a1[c]);
                a0[c] = fma2(pk2(qp[4 * rs]), pk2(wa.z, wa.z), a0[c]);
                a1[c] = fma2(pk2(qp[5 * rs]), pk2(wa.w, wa.w), a1[c]);
            }
        });
        static_for<CH>([&](auto c_) {
            constexpr int c = decltype(c_)::value;
            float2 v = up2(add2(a0[c], a1[c]));
            if constexpr (SCALE) v = cscale(v, pscale);
            if (act[c] && m[c] < p.n_bands) {
                float* o = reinterpret_cast<float*>(ob + (unsigned long long)unsigned(m[c]) * row_bytes);
                if (FULL || ok1) {  // frames fill a tile from column 0 up: ok1 implies ok0
                    vmax = max3(vmax, v.x, v.y);
                    tmin = min3(tmin, v.x, v.y);
                    o[0] = project_db<DBM>(v.x, dbc.amin, dbc.c1, dbc.c0);
                    o[LB] = project_db<DBM>(v.y, dbc.amin, dbc.c1, dbc.c0);
                } else if (ok0) {
                    vmax = fmaxf(vmax, v.x);
                    tmin = fminf(tmin, v.x);
                    o[0] = project_db<DBM>(v.x, dbc.amin, dbc.c1, dbc.c0);
                }
            }
        });
    }
}

template <int NW, int TTC, bool SCALE, bool BATCHED = false>
MLXA_D void project_power_tile(const FwdParams& p, const RowBank& rb, const DbConst& dbc, const float* s_pw, int tt, int b, int t0,
                               int nt, float pscale, int warp, float& vmax) {
    const int TT = TTC ? TTC : tt;
    float tmin = INFINITY;
    // db_mode: 0 raw, 1 dB (fast form valid), 2 dB with a denormal amin / ref quotient (set by the host)
    auto run = [&](auto dbm) {
        constexpr int DBM = decltype(dbm)::value;
        if constexpr (BATCHED) {
            if (nt == TT) project_power_tile_batched<NW, TTC, SCALE, DBM, true>(p, rb, dbc, s_pw, b, t0, nt, pscale, warp, vmax, tmin);
            else project_power_tile_batched<NW, TTC, SCALE, DBM, false>(p, rb, dbc, s_pw, b, t0, nt, pscale, warp, vmax, tmin);
        } else {
            if (nt == TT) project_power_tile_impl<NW, TTC, SCALE, DBM, true>(p, rb, dbc, s_pw, tt, b, t0, nt, pscale, warp, vmax, tmin);
            else project_power_tile_impl<NW, TTC, SCALE, DBM, false>(p, rb, dbc, s_pw, tt, b, t0, nt, pscale, warp, vmax, tmin);
        }
    };
    if (p.db_mode == 0) run(std::integral_constant<int, 0>{});
    else if (p.db_mode == 1) run(std::integral_constant<int, 1>{});
    else run(std::integral_constant<int, 2>{});
    if (p.block_min != nullptr) block_min_to_global(p, b, t0, tmin);
}

}  // namespace mlxa
