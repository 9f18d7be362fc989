// The compiled FFT plans.  MODE_PACK: one real frame of n_fft = 2N samples is packed as N
// complex points (even samples -> re, odd -> im) and unpacked with one extra twiddle.
// MODE_PAIR: two real frames of n_fft = N samples ride one N-point complex transform
// (frame a -> re, frame b -> im); used where N/2 has no balanced factorisation (n_fft = 400).
#pragma once
#include "fft_plan.cuh"

namespace mlxa {

enum : int { MODE_PACK = 0, MODE_PAIR = 1 };

template <int NFFT> struct PlanFor;  // ::Plan, ::MODE
#define MLXA_PLAN(NFFT, MODE_, N, G, ...)                 \
    template <> struct PlanFor<NFFT> {                    \
        using Plan = FftPlan<N, G, __VA_ARGS__>;          \
        static constexpr int MODE = MODE_;                \
        static constexpr int n_fft = NFFT;                \
    };
MLXA_PLAN(64, MODE_PACK, 32, 4, 8, 4)
MLXA_PLAN(128, MODE_PACK, 64, 8, 8, 8)
MLXA_PLAN(256, MODE_PACK, 128, 8, 16, 8)
MLXA_PLAN(400, MODE_PAIR, 400, 16, 25, 16)
MLXA_PLAN(512, MODE_PACK, 256, 16, 16, 16)
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
#undef MLXA_PLAN

// X(macro) over every planned n_fft
#define MLXA_FOR_EACH_NFFT(X) X(64) X(128) X(256) X(400) X(512) X(1024) X(2048) X(4096)

}  // namespace mlxa
