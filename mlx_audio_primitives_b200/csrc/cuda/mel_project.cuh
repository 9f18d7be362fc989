// Band-sparse filterbank projection of a tile of power spectra laid out [bin][frame] in shared memory
// (replaces the dense matmul of reference mel.py:344): lanes run along FRAMES.  A band row of the tile is
// TT/2 lanes x 2 frames (64-bit conflict-free reads, packed FFMA2), a warp covers 64/TT rows at once, the
// weights are warp-uniform float4 quads of the ROW-format bank, and every (band, frame pair) goes straight to
// global memory with the running max, the per-block minima and the optional dB fused.
#pragma once
#include "fwd_epilogue.cuh"

namespace mlxa {

// bank packed in ROW format (mlxa_plan_group == 1, see include/mlxa_cuda.h): quad-padded weight runs, then
// one int4 {start, n4, off4, len} per band
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

// power-tile geometry for TT frames: row stride (floats; even, == 2 mod 4 so frame pairs of consecutive
// rows spread over the banks) and row count (+3: rows the zero-padded weight quads may touch)
constexpr int power_tile_stride(int TT) { return TT + 2; }
constexpr int power_tile_rows(int n_bins) { return n_bins + 3; }

// TTC: tile frames when known at compile time, 0 -> tt (a power of two >= 2).  SCALE: multiply the sums by
// pscale (pair transforms deliver 4|X|^2 and the bank could not be pre-scaled).
template <int THREADS, int TTC, bool SCALE>
MLXA_D void project_power_tile(const FwdParams& p, const RowBank& rb, const float* s_pw, int tt, int b, int t0, int nt,
                               float pscale, float& vmax) {
    const int TT = TTC ? TTC : tt;
    const int PS = power_tile_stride(TT);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int LB = TT / 2, per_warp = 32 / LB, SLOTS = (THREADS / 32) * per_warp;  // band slots of the CTA per step
    const int fl = lane & (LB - 1), slot = warp * per_warp + lane / LB;
    const float db_ref = fmaxf(p.db_ref, p.db_amin);
    char* ob = reinterpret_cast<char*>(p.mel + (long long)b * p.n_bands * p.T + t0 + 2 * fl);
    const float2* q_lane = reinterpret_cast<const float2*>(s_pw) + fl;
    const unsigned row_bytes = unsigned(p.T) * 4u;
    const int rs = PS / 2;  // row stride in float2
    // Bands are taken two at a time per lane (two independent accumulation chains and epilogues in flight:
    // the phase is latency-bound otherwise); band m -> the lane's frames 2*fl, 2*fl + 1 of row m.  FULL:
    // every frame of the tile exists.
    float tmin = INFINITY;
    auto quad = [&](const float4* w4, const float2* q, float2 acc) {
        const float4 w = *w4;
        const float2 q0 = q[0], q1 = q[rs], q2 = q[2 * rs], q3 = q[3 * rs];
        return caxpy(w.w, q3, caxpy(w.z, q2, caxpy(w.y, q1, caxpy(w.x, q0, acc))));
    };
    auto band_pair = [&](auto full_, int mA, int mB) {
        constexpr bool FULL = decltype(full_)::value;
        const bool hasA = mA < p.n_bands, hasB = mB < p.n_bands;
        const int4 dA = hasA ? rb.desc[mA] : make_int4(0, 0, 0, 0);  // start, quads, first quad
        const int4 dB = hasB ? rb.desc[mB] : make_int4(0, 0, 0, 0);
        const float4 *wA = rb.wt4 + dA.z, *wB = rb.wt4 + dB.z;
        const float2 *qA = q_lane + dA.x * rs, *qB = q_lane + dB.x * rs;
        float2 accA = make_float2(0.f, 0.f), accB = make_float2(0.f, 0.f);
        int nA = dA.y, nB = dB.y;
#pragma unroll 1
        for (; nA > 0 && nB > 0; --nA, --nB, ++wA, ++wB, qA += 4 * rs, qB += 4 * rs) {
            accA = quad(wA, qA, accA);
            accB = quad(wB, qB, accB);
        }
#pragma unroll 1
        for (; nA > 0; --nA, ++wA, qA += 4 * rs) accA = quad(wA, qA, accA);
#pragma unroll 1
        for (; nB > 0; --nB, ++wB, qB += 4 * rs) accB = quad(wB, qB, accB);
        float v[4] = {accA.x, accA.y, accB.x, accB.y};
        if constexpr (SCALE) {
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] *= pscale;
        }
        const bool ok0 = FULL || 2 * fl < nt, ok1 = FULL || 2 * fl + 1 < nt;
        const bool st[4] = {ok0 && hasA, ok1 && hasA, ok0 && hasB, ok1 && hasB};
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (st[i]) { vmax = fmaxf(vmax, v[i]); tmin = fminf(tmin, v[i]); }
        if (p.db_mode) {
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = to_db_one(v[i], p.db_coef, p.db_amin, db_ref);
        }
        float* oA = reinterpret_cast<float*>(ob + (unsigned long long)unsigned(mA) * row_bytes);
        float* oB = reinterpret_cast<float*>(ob + (unsigned long long)unsigned(mB) * row_bytes);
        if (st[0]) oA[0] = v[0];
        if (st[1]) oA[1] = v[1];
        if (st[2]) oB[0] = v[2];
        if (st[3]) oB[1] = v[3];
    };
    // boustrophedon over the CTA's band slots: long and short bands mix
    if (nt == TT) {
#pragma unroll 1
        for (int m0 = 0; m0 < p.n_bands; m0 += 2 * SLOTS) band_pair(std::true_type{}, m0 + slot, m0 + 2 * SLOTS - 1 - slot);
    } else {
#pragma unroll 1
        for (int m0 = 0; m0 < p.n_bands; m0 += 2 * SLOTS) band_pair(std::false_type{}, m0 + slot, m0 + 2 * SLOTS - 1 - slot);
    }
    if (p.block_min != nullptr) block_min_to_global(p, b, t0, tmin);
}

}  // namespace mlxa
