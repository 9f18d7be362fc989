// Parameter blocks shared by the per-n_fft translation units and the C-ABI dispatcher.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace mlxa {

enum : int { EP_STFT = 0, EP_MEL = 1, EP_GL = 2, EP_FEAT = 3 };
enum : int { STAT_CENTROID = 0, STAT_BANDWIDTH = 1, STAT_ROLLOFF = 2, STAT_FLATNESS = 3 };
enum : int { POW_SQUARE = 0, POW_ABS = 1, POW_GENERAL = 2 };

// One-float MAX exchange over peer memory (NVLink / NVSwitch), replacing the all-reduce between the mel
// kernel and the dB kernel when clips are sharded over GPUs.  Every rank owns `2 * world` 64-bit slots in a
// peer-mapped buffer: slot [epoch & 1][r] holds (epoch << 32 | float bits of rank r's peak).  The LAST CTA of
// the producer kernel (ticket counter) stores this rank's peak into its slot on every peer; consumers spin
// on their own, local slots until all `world` entries carry the current epoch.  Alternating the two slot
// sets by epoch parity keeps a fast rank from overwriting a value a slow rank has not read yet.
struct PeakExchange {
    unsigned long long* const* peer_slots;  // device array: base of every rank's slots (nullptr: disabled)
    int rank, world;
    unsigned epoch;
    unsigned* ticket;  // local counter of finished CTAs, left at 0
};

struct FwdParams {
    // input clips
    const float* y;
    long long ldy;
    int L, B;
    // framing
    int T;        // frames produced per clip
    int T_valid;  // frames >= T_valid see all-zero input (Griffin-Lim frame padding)
    int F;        // n_fft/2 + 1
    int n_fft, hop, pad, pad_mode;
    int tile_frames;  // frames per tile (a power of two for the mel epilogue)
    int log2_tile;    // log2(tile_frames) when it is a power of two
    int n_in_buf;     // 1 or 2 staging buffers for the tile samples (2 = next tile prefetched by TMA)
    const float* window;      // n_fft
    const float2* tw_plan;    // inter-pass twiddles of the plan (or the full table for the naive DFT)
    const float2* tw_unpack;  // exp(-i*pi*k/N), k in [0, N]  (packed real transform)
    // EP_STFT
    float2* spec;  // (B, T, F)
    // EP_MEL
    int power_mode;
    float power;
    const float* bank;   // packed band-sparse filterbank (device), see fwd_epilogue.cuh
    int n_bands;
    long long n_w4;      // transposed weight words (n_wt) in the packed bank
    int const_bulk;      // window / bank pointers are 16-byte aligned -> bulk async copies
    int bank_in_smem;    // 0: the packed bank is too large for shared memory and is read from global
    float* mel;   // (B, n_bands, T)
    float* gmax;  // optional running max
    PeakExchange xchg;    // optional: publish the final gmax to every peer
    float* block_min;     // optional (B, blocks_per_clip): min of the raw values per 64-frame block of a clip
    int blocks_per_clip;  // ceil(T / 64)
    int db_mode;          // 0 raw values, 1 dB, 2 dB with amin below the normal range (no flush-to-zero logarithm)
    float db_coef, db_amin, db_ref;
    // EP_FEAT: one per-frame spectral statistic (STAT_*), the spectrum never leaves the SM
    int feat_kind, feat_norm;
    float feat_p1, feat_p2;        // p | roll_percent | power, amin
    const float* feat_freq;        // F bin frequencies
    float feat_freq_step;          // > 0: the frequencies are k * step (default linspace): no table reads
    const float* feat_centroid;    // optional (B, T): bandwidth around a given centroid
    float* feat_out;               // (B, T)
    // EP_GL
    const float* mag;  // (B, T, F)
    float2* rebuilt;   // (B, T, F), out: mag * X/|X|
};

// 32-bit words of a packed band-sparse filterbank (layout: fwd_epilogue.cuh / mlxa_cuda.h)
// group = -GP < 0 is the ROW-PAIR format (mel_project.cuh): [wt][int4 {start, nq, off, len} per band, bands padded to a multiple of 32]
__host__ __device__ inline long long packed_bank_words(int n_bands, long long n_wt, int group) {
    if (group < 0) return n_wt + 4LL * ((n_bands + 31) / 32 * 32);
    return (n_wt + 2LL * n_bands + 2LL * ((n_bands + group - 1) / group) + 3) & ~3LL;
}

struct InvParams {
    const float2* spec;       // (B, T, F_in)
    const float* u_prev;      // optional (B, out_len), stride ldy: y = u + momentum * (u - u_prev), u = istft(spec)
    float* u_out;             // optional (B, out_len), stride ldy: receives u
    float momentum;
    int const_bulk;           // window pointer is 16-byte aligned -> bulk async copy
    int B, T, F_in;
    int n_fft, hop;
    int tile_hops;       // output tile = tile_hops * hop samples
    int spaced;          // frames of a round are r = ceil(n_fft / hop) apart: they never overlap, one barrier per round
    const float* window; // n_fft
    const float* wss;    // ola_len
    const float2* tw_plan;
    const float2* tw_unpack;
    long long ola_len, trim, out_len, ldy;
    float* y;            // (B, out_len)
};

struct AcfParams {  // acf_inst.cu: autocorrelation pitch detector (pitch.py:118-260)
    const float* y;
    long long ldy, T;
    int B, L;
    int frame_length, hop, pad;
    int min_lag, max_lag;
    float threshold, sr;
    const float2* tw_plan;
    const float2* tw_unpack;
    float* f0;             // (B, T)
    unsigned char* voiced; // (B, T)
};

// one launcher per compiled n_fft (fwd_inst.cu / inv_inst.cu built with -DMLXA_NFFT=...)
#define MLXA_DECL_LAUNCHERS(NF)                                                            \
    cudaError_t launch_fwd_##NF(int ep, FwdParams& p, cudaStream_t s);                     \
    cudaError_t launch_inv_##NF(InvParams& p, cudaStream_t s);                             \
    int plan_group_##NF();                                                                  \
    int plan_fused_feature_##NF();                                                          \
    void plan_tables_##NF(float2* tw_plan_host, int* n_plan, float2* tw_unpack_host, int* n_unpack);

// naive O(n^2) DFT fallback for any other n_fft
cudaError_t launch_fwd_naive(int ep, FwdParams& p, cudaStream_t s);
cudaError_t launch_irdft_naive(const float2* spec, long long rows, int F_in, int n_fft,
                               const float2* tw_full, float* frames, cudaStream_t s);

}  // namespace mlxa
