// The compiled FFT plans.  MODE_PACK: one real frame of n_fft = 2N samples is packed as N
// complex points (even samples -> re, odd -> im) and unpacked with one extra twiddle.
// MODE_PAIR: two real frames of n_fft = N samples ride one N-point complex transform
// (frame a -> re, frame b -> im); used where N/2 has no balanced factorisation (n_fft = 400).
#pragma once
#include "fft_plan.cuh"
#include "fft_sizes.cuh"

namespace mlxa {

enum : int { MODE_PACK = 0, MODE_PAIR = 1 };

template <int NFFT> struct PlanFor;  // ::Plan, ::MODE
#define MLXA_PLAN(NFFT, MODE_, N, G, ...)                 \
    template <> struct PlanFor<NFFT> {                    \
        using Plan = FftPlan<N, G, __VA_ARGS__>;          \
        static constexpr int MODE = MODE_;                \
        static constexpr int n_fft = NFFT;                \
    };
MLXA_PLAN(32, MODE_PACK, 16, 4, 4, 4)
MLXA_PLAN(64, MODE_PACK, 32, 4, 8, 4)
MLXA_PLAN(128, MODE_PACK, 64, 8, 8, 8)
MLXA_PLAN(256, MODE_PACK, 128, 8, 16, 8)
MLXA_PLAN(400, MODE_PAIR, 400, 16, 25, 16)
MLXA_PLAN(480, MODE_PACK, 240, 16, 16, 15)       // 15 of 16 lanes in pass 0
MLXA_PLAN(512, MODE_PACK, 256, 16, 16, 16)
MLXA_PLAN(600, MODE_PACK, 300, 16, 20, 15)       // 15 lanes in pass 0, 20 butterflies in two rounds after it
MLXA_PLAN(800, MODE_PACK, 400, 16, 25, 16)       // the n_fft = 400 pair plan, here on one packed frame
MLXA_PLAN(1000, MODE_PACK, 500, 32, 20, 25)
MLXA_PLAN(1200, MODE_PACK, 600, 32, 24, 25)
MLXA_PLAN(1600, MODE_PACK, 800, 32, 32, 25)
MLXA_PLAN(2000, MODE_PACK, 1000, 32, 40, 25)     // 50 values per lane in the last pass: the one-CTA, 255-register class of n_fft = 4096
MLXA_PLAN(3072, MODE_PACK, 1536, 32, 48, 32)     // 64 values per lane, as n_fft = 4096
MLXA_PLAN(1024, MODE_PACK, 512, 16, 32, 16)
MLXA_PLAN(2048, MODE_PACK, 1024, 32, 32, 32)
#ifndef MLXA_PLAN_4096_TWO_WARPS
MLXA_PLAN(4096, MODE_PACK, 2048, 32, 64, 32)     // one warp, 64 values per lane: 255 registers, 8 warps per SM
#else
// Two warps per transform (group_sync = named barrier), 32 values per lane, 16 warps per SM.  Parity-green but
// slower (MFCC c4 1.93 ms vs 1.80 ms): three passes and an unpack through shared memory cost more wavefronts
// than the doubled occupancy hides.  Kept for experiments (-DMLXA_PLAN_4096_TWO_WARPS).
MLXA_PLAN(4096, MODE_PACK, 2048, 64, 32, 8, 8)
#endif
MLXA_PLAN(8192, MODE_PACK, 4096, 64, 64, 64)     // two warps per transform (named barrier), 64 values per lane
#undef MLXA_PLAN

}  // namespace mlxa
