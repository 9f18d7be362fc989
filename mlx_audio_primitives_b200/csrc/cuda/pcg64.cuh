// NumPy-compatible PCG64 (XSL-RR 128/64) stream generation with O(log n) jump-ahead, so that the
// Griffin-Lim random phase init (reference griffinlim.py:112-115: default_rng(seed).uniform(-pi, pi)
// drawn on the HOST, ~0.25 s for BASELINE config 5) can be produced on the device with the SAME
// bits: value i is low + range * ((next_uint64_i >> 11) * 2^-53), rounded once to float32.
#pragma once
#include "common.cuh"

namespace mlxa {

typedef unsigned __int128 u128;

MLXA_HD u128 pcg_mult() { return ((u128)0x2360ED051FC65DA4ULL << 64) | (u128)0x4385DF649FCCF645ULL; }

// state after `delta` steps of  s <- s*mult + inc  (Brown's LCG jump-ahead)
MLXA_HD u128 pcg_advance(u128 state, u128 inc, unsigned long long delta) {
    u128 acc_mult = 1, acc_plus = 0, cur_mult = pcg_mult(), cur_plus = inc;
    while (delta > 0) {
        if (delta & 1) {
            acc_mult *= cur_mult;
            acc_plus = acc_plus * cur_mult + cur_plus;
        }
        cur_plus = (cur_mult + 1) * cur_plus;
        cur_mult *= cur_mult;
        delta >>= 1;
    }
    return acc_mult * state + acc_plus;
}

// NumPy's pcg64_next64: step, then XSL-RR output of the NEW state
MLXA_HD unsigned long long pcg_next64(u128& state, u128 inc) {
    state = state * pcg_mult() + inc;
    const unsigned long long hi = (unsigned long long)(state >> 64), lo = (unsigned long long)state;
    const unsigned long long x = hi ^ lo;
    const unsigned rot = (unsigned)(hi >> 58);
    return (x >> rot) | (x << ((64 - rot) & 63));
}

MLXA_HD float pcg_uniform_f32(u128& state, u128 inc, double low, double range) {
    const double d = double(pcg_next64(state, inc) >> 11) * (1.0 / 9007199254740992.0);
    return float(low + range * d);
}

}  // namespace mlxa
