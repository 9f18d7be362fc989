/*
 * mlxa_cuda.h -- C ABI of libmlxaudio_cuda.so, the sm_100a replacement for the reference's
 * native extension `mlx_audio_primitives._ext` (reference csrc/bindings.cpp:11) on the spectral
 * hot path.  Plain pointers and sizes only; every pointer is a DEVICE pointer unless the name
 * says `host`.  Every entry point is asynchronous on `stream` (a cudaStream_t passed as void*),
 * never allocates or frees caller-visible memory, and returns 0 on success, a negative
 * MLXA_E_* code for an invalid argument or a positive cudaError_t.  `mlxa_last_error()` gives
 * the message of the last failure on the calling thread.  Nothing throws across this boundary
 * (the reference throws std::invalid_argument -> ValueError; the Python host layer raises the
 * same ValueErrors before calling in).
 *
 * Layouts (row-major, contiguous unless a leading dimension is given):
 *   clips      y      (B, L)        float32, row stride ldy elements
 *   spectrum   spec   (B, T, F)     complex64 {re, im}, F = n_fft/2 + 1.  This is the physical
 *                                   buffer behind the reference's logical (B, F, T) result,
 *                                   which is itself a transposed view of (B, T, F)
 *                                   (reference stft.py:216, :292).
 *   mel / mfcc        (B, n_mels|n_mfcc, T)   float32 (reference mel.py:344, mfcc.py:271)
 */
#ifndef MLXA_CUDA_H
#define MLXA_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MLXA_ABI_VERSION 6

#define MLXA_E_INVALID (-1)   /* bad size / null pointer / unknown mode            */
#define MLXA_E_UNSUPPORTED (-2)

/* pad modes (reference stft.py:434-468, csrc/primitives/pad_signal.cpp:36-48) */
#define MLXA_PAD_CONSTANT 0
#define MLXA_PAD_REFLECT 1
#define MLXA_PAD_EDGE 2

typedef struct { float re, im; } mlxa_c64;

/* One-float MAX exchange over peer memory (NVLink / NVSwitch) for a batch sharded over `world` GPUs: what the
 * reference gets from taking ref=max / top_db over the whole batch (convert.py:42-58).  peer_slots is a
 * DEVICE array of `world` pointers, entry r the base of rank r's slot buffer of 2*world uint64 (zeroed once,
 * peer-mapped on every rank, e.g. torch symmetric memory); ticket a zeroed device uint32 owned by this rank;
 * epoch a counter the caller raises by one per producer launch, identical on all ranks, starting at 1.  The
 * last CTA of the producer (mlxa_melspec_f32) stores (epoch, peak) into slot [epoch & 1][rank] of every rank;
 * consumers given the same descriptor (mlxa_to_db_f32, mlxa_db_floor_blocks_f32) spin on their own slots
 * until all ranks have published this epoch and use the maximum in place of *gmax_dev (and of a ref_dev that
 * aliases gmax_dev).  Pass NULL for single-GPU use. */
typedef struct {
    uint64_t* const* peer_slots;
    int32_t rank, world;
    uint32_t epoch;
    uint32_t* ticket;
} mlxa_peak_exchange;

int mlxa_abi_version(void);
const char* mlxa_last_error(void);
/* 1 when n_fft runs on a compiled Stockham plan, 0 when it falls to the O(n^2) DFT kernel. */
int mlxa_has_fast_plan(int n_fft);

/* ---- reference-boundary entry points (same granularity as the nanobind module) ---------- */

/* replaces _ext.pad_signal (csrc/bindings.cpp:77, primitives/pad_signal.cpp:133):
 * out (B, L + 2*pad).  reflect requires pad <= L - 1. */
int mlxa_pad_signal_f32(const float* x, int64_t B, int64_t L, int64_t pad, int mode,
                        float* out, void* stream);

/* replaces _ext.frame_signal (csrc/bindings.cpp:49, primitives/frame_signal.cpp:119):
 * out (B, T, frame_length), T = 1 + (L - frame_length) / hop. */
int mlxa_frame_signal_f32(const float* x, int64_t B, int64_t L, int frame_length, int hop,
                          float* out, void* stream);

/* replaces _ext.overlap_add (csrc/bindings.cpp:16, primitives/overlap_add.cpp:195,
 * metal/overlap_add.metal:16): gather OLA of raw frames (B, T, n_fft) with window applied
 * inside and division by max(sum w^2, 1e-8); out (B, out_len). */
int mlxa_overlap_add_f32(const float* frames, const float* window, int64_t B, int64_t T,
                         int n_fft, int hop, int64_t out_len, float* out, void* stream);

/* sum_f w[i - f*hop]^2 over the T existing frames, ascending f (overlap_add.metal:36-50);
 * wss (out_len).  The fused ISTFT takes this envelope as an input. */
int mlxa_window_sumsquare_f32(const float* window, int n_fft, int hop, int64_t T,
                              int64_t out_len, float* wss, void* stream);

/* ---- fused hot-path kernels -------------------------------------------------------------- */

/* pad -> frame -> window -> rFFT in one kernel (replaces the chain stft.py:118-130 ->
 * _ext.pad_signal -> _ext.frame_signal -> mx.fft.rfft).  window: n_fft floats, already
 * centre-padded (stft.py:88-107).  center != 0 pads n_fft/2 per side with pad_mode.
 * spec (B, T, F), T = 1 + (L + 2*pad - n_fft) / hop. */
int mlxa_stft_f32(const float* y, int64_t B, int64_t L, int64_t ldy, const float* window,
                  int n_fft, int hop, int center, int pad_mode, mlxa_c64* spec, void* stream);

/* 1 when mlxa_spectral_feature_f32 serves n_fft (a compiled plan whose lane groups fit one warp). */
int mlxa_has_fused_feature(int n_fft);

/* How the mel kernel that serves n_fft wants its filterbank packed: -GP < 0 = ROW-PAIR format whose GP
 * adjacent bands (the bands one warp step of the projection covers) share a trip count -- every compiled plan;
 * 32 = lane-group format of the O(n^2) kernels that serve sizes without a compiled plan. */
int mlxa_plan_group(int n_fft);

/* Packed band-sparse filterbank ("bank") for a lane group of `group` = mlxa_plan_group(n_fft):
 * a 16-byte aligned array of 32-bit words
 *     float   wt[n_wt]          bands are taken `group` at a time; inside group j the weights are
 *                               transposed: wt[goff[j] + i*group + g] = weight of band j*group + g
 *                               at bin start[band] + i, zero-padded to glen[j] rows
 *     int32   start[n_bands]    first frequency bin of each row's contiguous support
 *     int32   len[n_bands]      support length in bins
 *     int32   goff[n_groups], glen[n_groups]      n_groups = ceil(n_bands / group)
 * padded to mlxa_packed_bank_words(n_bands, n_wt, group) words (a multiple of 4).
 * ROW-PAIR format (group = -GP < 0; the projection runs with lanes along frames, a lane holding two frames):
 *     float   wt[n_wt]          per band 1 + nq entries of 4 words: entry 0 = {w[0], w[1], 0, 0} -- the pair of
 *                               bins that is always there -- then nq quads of bins {w[2+4i] .. w[5+4i]};
 *                               zero-padded; the GP adjacent bands m0*GP .. m0*GP + GP - 1 share nq
 *                               (warp-uniform loops)
 *     int32   desc[n_pad][4]    {first bin of the run (moved down where the run would leave the F + 3 rows of
 *                               the kernel's power tile), nq, first entry of the run (wt + 4*entry), support
 *                               length in bins}; n_pad = n_bands rounded up to a multiple of 32, the padding
 *                               bands carry zero weights
 * The kernels
 * bulk-copy this blob into shared memory once per CTA.  mlxa_pack_filterbank builds it on the
 * HOST from a dense (n_bands, F) row-major matrix (rows must have contiguous support, which
 * triangular mel / linear / Bark banks have); returns n_wt through *n_wt_out.  Call it with
 * packed_host == NULL to size the buffer. */
int64_t mlxa_packed_bank_words(int n_bands, int64_t n_wt, int group);
int mlxa_pack_filterbank(const float* dense_host, int n_bands, int F, int group, float* packed_host,
                         int64_t capacity_words, int64_t* n_wt_out);

/* STFT with the |X|^power + band-sparse filterbank epilogue (replaces mel.py:309-352 =
 * stft -> abs -> power -> dense matmul); the spectrum never reaches HBM.
 * bank: packed filterbank on the DEVICE, packed for mlxa_plan_group(n_fft); n_wt as returned by the packer.
 * mel (B, n_bands, T).  gmax (optional, may be NULL): device float, atomically raised to
 * max(mel) -- the producer side of power_to_db(ref=max / top_db) (convert.py:42-58).
 * db_mode != 0 writes db_coef*log10(max(v, db_amin)/max(db_ref, db_amin)) instead of v
 * (the no-global-max form of convert.py:48-52).
 * block_min (optional, may be NULL): (B, ceil(T / MLXA_MIN_BLOCK_FRAMES)) device floats the caller
 * initialised to +inf; each slot is atomically lowered to the smallest raw mel value of that 64-frame
 * block of the clip -- what mlxa_db_floor_blocks_f32 needs to skip blocks the top_db floor cannot touch. */
#define MLXA_MIN_BLOCK_FRAMES 64
int mlxa_melspec_f32(const float* y, int64_t B, int64_t L, int64_t ldy, const float* window,
                     int n_fft, int hop, int center, int pad_mode, float power,
                     const float* bank, int n_bands, int64_t n_wt,
                     float* mel, float* gmax, int db_mode, float db_coef, float db_amin,
                     float db_ref, float* block_min, const mlxa_peak_exchange* xchg, void* stream);

/* irFFT -> window -> gather overlap-add -> / max(sum w^2, 1e-8) -> trim in one kernel
 * (replaces stft.py:292-338 = mx.fft.irfft -> _ext.overlap_add -> slices).
 * spec (B, T, F_in) complex64; bins k >= F_in are zero, bins beyond n_fft/2 ignored
 * (irfft(n=n_fft) semantics, stft.py:295).  OLA runs over ola_len samples; output sample j
 * is OLA sample j + trim for j < min(out_len, ola_len - trim), zero after that
 * (stft.py:315-338).  wss: ola_len floats from mlxa_window_sumsquare_f32. */
int mlxa_istft_f32(const mlxa_c64* spec, int64_t B, int64_t T, int F_in, const float* window,
                   const float* wss, int n_fft, int hop, int64_t ola_len, int64_t trim,
                   int64_t out_len, float* y, int64_t ldy, void* stream);

/* Same kernel with Griffin-Lim's momentum step (griffinlim.py:176-178: rebuilt = new + m*(new - tprev),
 * then istft(rebuilt)) applied in the SIGNAL domain -- the inverse STFT is linear:
 *   u = istft(spec);  y = u + momentum*(u - u_prev)   (y = u when u_prev is NULL or momentum == 0).
 * u_out (optional) receives u, the u_prev of the next iteration; u_prev, u_out and y are distinct (B, out_len)
 * buffers of row stride ldy.  The previous projection is never read again (8*F*T bytes per clip saved for
 * 8*L bytes of signal traffic).  Works for every n_fft (the O(n^2) path adds one small kernel). */
int mlxa_istft_momentum_f32(const mlxa_c64* spec, const float* u_prev, float momentum, float* u_out,
                            int64_t B, int64_t T, int F_in, const float* window, const float* wss,
                            int n_fft, int hop, int64_t ola_len, int64_t trim, int64_t out_len,
                            float* y, int64_t ldy, void* stream);

/* One Griffin-Lim projection (griffinlim.py:143-169) fused into the STFT epilogue:
 *   X = stft(y); projected = mag * X/|X|  (mag + 0j where X == 0, i.e. angle 0).
 * mag and projected are (B, T, F); frames t >= T_valid see X = 0 (zero-padded frames,
 * griffinlim.py:159-165).  The momentum step rebuilt = new + m*(new - prev) is NOT materialised:
 * mlxa_istft_momentum_f32 applies it to the inverse transforms of the projections. */
int mlxa_griffinlim_project_f32(const float* y, int64_t B, int64_t L, int64_t ldy,
                                const float* window, int n_fft, int hop, int center,
                                int pad_mode, int64_t T, int64_t T_valid, const float* mag,
                                mlxa_c64* projected, void* stream);

/* out = x + momentum * (x - x_prev) over n floats (complex arrays: 2n): the Griffin-Lim momentum step
 * (griffinlim.py:176-178) as a standalone elementwise kernel, for callers that drive single iterations. */
int mlxa_momentum_f32(const float* x, const float* x_prev, float momentum, int64_t n, float* out, void* stream);

/* rebuilt = mag * exp(i*angles) elementwise over n values (griffinlim.py:123) */
int mlxa_polar_f32(const float* mag, const float* angles, int64_t n, mlxa_c64* out, void* stream);

/* out[i] = float32(low + (high - low) * u_i), u_i the i-th double of NumPy's PCG64 stream whose
 * 128-bit state / increment are given as two 64-bit halves each -- bit-identical to
 * np.random.Generator(PCG64).uniform(low, high, n).astype(float32).  Replaces the host-side draw of
 * the Griffin-Lim phase init (griffinlim.py:112-115) while keeping seeds reproducible. */
int mlxa_pcg64_uniform_f32(uint64_t state_hi, uint64_t state_lo, uint64_t inc_hi, uint64_t inc_lo,
                           double low, double high, int64_t n, float* out, void* stream);

/* Griffin-Lim's random start in one pass (griffinlim.py:112-123): out[b,t,f] = mag[b,t,f] * exp(i * a), a the
 * ((b*F + f)*T + t)-th value of the stream above -- the phases are drawn in the reference's logical (B, F, T) order
 * while magnitudes and spectrum live in the physical (B, T, F) layout; the angle tensor never exists in HBM.
 * Bit-identical to mlxa_pcg64_uniform_f32 + a transpose + mlxa_polar_f32. */
int mlxa_pcg64_polar_f32(uint64_t state_hi, uint64_t state_lo, uint64_t inc_hi, uint64_t inc_lo, double low, double high,
                         const float* mag, int64_t B, int64_t F, int64_t T, mlxa_c64* out, void* stream);

/* |z| and atan2(im, re) over n complex values (stft.py:347-379) */
int mlxa_magnitude_f32(const mlxa_c64* z, int64_t n, float* out, void* stream);
int mlxa_phase_f32(const mlxa_c64* z, int64_t n, float* out, void* stream);
/* strided (B,T,F) -> contiguous logical (B,F,T) transposing variants for the Python layer */
int mlxa_transpose_f32(const float* in, int64_t B, int64_t R, int64_t C, float* out, void* stream);
int mlxa_transpose_c64(const mlxa_c64* in, int64_t B, int64_t R, int64_t C, mlxa_c64* out, void* stream);

/* max over n floats, atomically folded into *gmax (caller zero-initialises; values >= 0 or
 * any sign: the kernel uses an order-preserving integer key). */
int mlxa_max_f32(const float* x, int64_t n, float* gmax, void* stream);
int mlxa_fill_f32(float* x, int64_t n, float value, void* stream);

/* convert.py:14-60.  out = coef*log10(max(x, amin)/max(ref, amin)); ref is *ref_dev when
 * ref_dev != NULL (a device scalar, e.g. the fused max for ref=max) else ref_host.
 * use_top_db: clamp at coef*log10(max(*gmax_dev, amin)/max(ref, amin)) - top_db, the global
 * max of the output by monotonicity.  In-place (out == x) is allowed.  reset_next (optional): a
 * device float set to 0 by this launch -- the peak slot the NEXT producer call will atomically raise,
 * so a pipeline alternating two slots needs no separate fill launch. */
int mlxa_to_db_f32(const float* x, int64_t n, float coef, float amin, float ref_host,
                   const float* ref_dev, int use_top_db, float top_db, const float* gmax_dev,
                   float* out, float* reset_next, const mlxa_peak_exchange* xchg, void* stream);
/* Second half of a fused log-mel (convert.py:55-58 alone): x_db already holds coef*log10(max(S, amin)/
 * max(ref, amin)) -- mlxa_melspec_f32 with db_mode != 0 and gmax -- and is raised in place to
 * max(x_db) - top_db, max(x_db) derived from the peak *gmax_dev of S.  Read-mostly: only values below the
 * floor are written.  Bit-identical to mlxa_to_db_f32 on the raw mel values.  reset_next as above. */
int mlxa_db_floor_f32(float* x_db, int64_t n, float coef, float amin, float ref, float top_db,
                      const float* gmax_dev, float* reset_next, void* stream);
/* The same floor, block-wise: x_db (B, n_bands, T) with the producer's block_min.  A 64-frame block of a
 * clip whose minimum already clears the floor is skipped entirely, so the pass costs B*ceil(T/64) reads
 * unless the material really spans top_db of dynamic range.  Consumed block_min slots are reset to +inf.
 * raised (optional, 1 + B*ceil(T/64) device int32, raised[0] zeroed by the caller): raised[0] counts the
 * rewritten blocks, raised[1..] lists them as b*ceil(T/64) + block (any order).  Same bits as
 * mlxa_to_db_f32 / mlxa_db_floor_f32. */
int mlxa_db_floor_blocks_f32(float* x_db, int64_t B, int n_bands, int64_t T, float coef, float amin,
                             float ref, float top_db, const float* gmax_dev, float* block_min,
                             float* reset_next, int32_t* raised, const mlxa_peak_exchange* xchg, void* stream);
/* convert.py:100-129,169-198: out = ref * 10^(x/div) */
int mlxa_from_db_f32(const float* x, int64_t n, float ref, float div, float* out, void* stream);

/* mfcc.py:69-140 / csrc/primitives/dct.cpp:103: out (rows, n_out) = x (rows, n_in) @ D^T,
 * D (n_out, n_in) row-major. */
int mlxa_dct_f32(const float* x, int64_t rows, int n_in, const float* D, int n_out, float* out,
                 void* stream);

/* mfcc.py:257-282 in one pass over a (B, n_mels, T) mel tensor: dB (ref, amin, top_db against
 * *gmax_dev), DCT over the mel axis with D (n_mfcc, n_mels), optional lifter (n_mfcc) ->
 * out (B, n_mfcc, T).  apply_db == 0 skips the dB step (caller-supplied log-mel, mfcc.py:229). */
int mlxa_mfcc_tail_f32(const float* mel, int64_t B, int n_mels, int64_t T, const float* D,
                       int n_mfcc, const float* lifter, int apply_db, float amin, float ref,
                       int use_top_db, float top_db, const float* gmax_dev, float* out,
                       void* stream);

/* ---- the callers either side of the path (reference features.py / framing.py) ----------- */
/* Per-frame spectral statistics over the F bins of each of the B*T frames (reference features.py:57-442; native
 * twins csrc/primitives/spectral.cpp:8-257).  S is the PHYSICAL (B, T, F) array -- complex64 as written by
 * mlxa_stft_f32 (is_complex != 0: |X| is formed on load) or float32.  freq: F bin frequencies.  kind:
 * 0 centroid; 1 bandwidth (p1 = p, norm, optional centroid_in (B*T)); 2 rolloff (p1 = roll_percent);
 * 3 flatness (p1 = power applied to |X| / S, p2 = amin).  out (B*T). */
int mlxa_spectral_stats_f32(const void* S, int is_complex, int64_t B, int64_t T, int F, const float* freq,
                            int kind, float p1, float p2, int norm, const float* centroid_in, float* out,
                            void* stream);
/* Spectral contrast (features.py:445-592; host NumPy in the reference): per frame and band, mean of the nq
 * largest minus mean of the nq smallest magnitudes of the band's bins, as a difference of 10*log10 values
 * (linear == 0) or plainly.  S as for mlxa_spectral_stats_f32; bands: n_out DEVICE triples {first bin, bin
 * count, nq} produced on the host by the reference's band-edge rules (a band with count 0 yields 0);
 * out (B, n_out, T). */
int mlxa_spectral_contrast_f32(const void* S, int is_complex, int64_t B, int64_t T, int F,
                               const int32_t* bands, int n_out, int linear, float* out, void* stream);
/* The same statistics FROM AUDIO in one kernel: pad -> frame -> window -> rFFT as mlxa_stft_f32, |X| kept in the
 * registers of the lane group that produced it, the statistic reduced by shuffles inside the group -- the
 * spectrum is never written.  n_fft must satisfy mlxa_has_fused_feature; other sizes use
 * mlxa_stft_f32 + mlxa_spectral_stats_f32.  freq_step > 0 declares freq[k] == k * freq_step (the default
 * linspace(0, sr/2, F)): centroid / bandwidth then skip the table reads (rolloff always reads its one entry).
 * out (B, T). */
int mlxa_spectral_feature_f32(const float* y, int64_t B, int64_t L, int64_t ldy, const float* window,
                              int n_fft, int hop, int center, int pad_mode, const float* freq,
                              float freq_step, int kind, float p1, float p2, int norm,
                              const float* centroid_in, float* out, void* stream);
/* Autocorrelation pitch detector (pitch.py:118-260, a per-frame NumPy loop on the host in the reference), one kernel:
 * per frame, r = irfft(|rfft(frame - mean, n_fft)|^2) / r[0] with n_fft the power of two >= 2*frame_length - 1
 * (frame_length 33..2048), then f0 = sr / lag of the first local maximum above `threshold` in [int(sr/fmax),
 * int(sr/fmin)], else of the range's global maximum if above it.  Centre padding is zeros.  f0 (B, T) float32,
 * voiced (B, T) uint8, T = 1 + (L + 2*(frame_length/2 if center) - frame_length) / hop. */
int mlxa_pitch_acf_f32(const float* y, int64_t B, int64_t L, int64_t ldy, int frame_length, int hop, int center,
                       float sr, float fmin, float fmax, float threshold, float* f0, uint8_t* voiced, void* stream);
/* Fourier-method resampling of whole signals to `num` samples (resample.py:84-139 -> scipy.signal.resample on the host
 * in the reference): rfft, keep / zero-extend the spectrum (unpaired Nyquist bin doubled when shrinking, halved when
 * growing), irfft, times num / n and `gain`.  Arbitrary lengths up to 2^24: both DFTs are Bluestein chirp transforms over
 * power-of-two Stockham FFTs in global memory.  work: device scratch of mlxa_resample_fft_work_bytes(B, n, num) bytes.
 * Chirps and their transforms are cached per length on the device (first call of a length builds them). */
int64_t mlxa_resample_fft_work_bytes(int64_t B, int64_t n, int64_t num);
int mlxa_resample_fft_f32(const float* x, int64_t B, int64_t n, int64_t ldx, int64_t num, float gain, float* out, int64_t ldo,
                          void* work, int64_t work_bytes, void* stream);
/* Rational resampling along the last axis (resample.py:215-300, scipy.signal.resample_poly on the host in the
 * reference): out[r, j] = sum_i x[r, i] * h[(j + pre_remove)*down - i*up].  h (len_h, DEVICE) is the zero-padded Kaiser
 * low-pass already multiplied by `up`, pre_remove / n_out as scipy derives them -- all computed on the host
 * (mlx_audio_primitives_b200/resample.py restates the design).  x (rows, n_in) -> out (rows, n_out). */
int mlxa_resample_poly_f32(const float* x, int64_t rows, int64_t n_in, const float* h, int len_h, int up, int down,
                           int64_t pre_remove, int64_t n_out, float* out, void* stream);
/* Linear-interpolation resampling (resample.py:142-212): positions linspace(0, n_in - 1, n_out), blend in float64 like
 * NumPy, optional gain (scale=True), rounded to float32. */
int mlxa_resample_linear_f32(const float* x, int64_t rows, int64_t n_in, int64_t n_out, double gain, int apply_gain,
                             float* out, void* stream);
/* Periodicity per frame (pitch.py:267-383): max of r / r[0] over lags [int(sr/fmax), int(sr/fmin)] of the frame's
 * autocorrelation (same engine passes as mlxa_pitch_acf_f32), 0 for silent frames or an empty range.  out (B, T). */
int mlxa_periodicity_f32(const float* y, int64_t B, int64_t L, int64_t ldy, int frame_length, int hop, int center, float sr,
                         float fmin, float fmax, float* out, void* stream);
/* De-emphasis out[n] = y[n] + coef * out[n-1] (framing.py:298-392; scipy.signal.lfilter on the host in the reference):
 * a block-wise scan of the recurrence, float64 inside.  zi: (B) initial states added to out[0], or NULL for zero;
 * librosa_zi != 0 ignores zi and applies the reference's default correction -corr * coef^n,
 * corr = ((2 - coef) y[0] - y[1]) / (3 - coef).  zf: (B) final states as scipy reports them, or NULL.  out != y. */
int mlxa_deemphasis_f32(const float* y, int64_t B, int64_t n, int64_t ldy, double coef, const float* zi, int librosa_zi,
                        float* out, int64_t ldo, float* zf, void* stream);
/* Whole-signal autocorrelation r[k] = sum_n y[n] y[n + k], k < max_lag <= n, of the (optionally mean-removed) clips,
 * optionally divided by max(r[0], 1e-10) (pitch.py:16-116; the reference takes it from one zero-padded FFT of the entire
 * signal).  Direct, deterministic sum, float32 inside 2048-sample chunks and float64 across them: O(n * max_lag).
 * out (B, max_lag); scratch: 2*B device floats. */
int mlxa_autocorrelation_f32(const float* y, int64_t B, int64_t n, int64_t ldy, int max_lag, int normalize, int center,
                             float* out, float* scratch, void* stream);
/* The same result through two global-memory transforms, r = FFT_M(|FFT_M(y - mean)|^2) / M with M the power of two
 * >= 2n - 1 (what pitch.py:16-116 does in one FFT call): O(M log M), the choice for more than ~1000 lags.  float32
 * transforms (error ~1e-6 of r[0]).  work: device scratch of mlxa_autocorrelation_fft_work_bytes(B, n) bytes. */
int64_t mlxa_autocorrelation_fft_work_bytes(int64_t B, int64_t n);
int mlxa_autocorrelation_fft_f32(const float* y, int64_t B, int64_t n, int64_t ldy, int max_lag, int normalize, int center,
                                 float* out, float* scratch, void* work, int64_t work_bytes, void* stream);
/* Savitzky-Golay filter along the last axis: the `delta` features of reference mfcc.py:290-371, which calls
 * scipy.signal.savgol_filter on the host.  x, out (rows, T); taps: `width` correlation taps (out[t] = sum_j taps[j] *
 * x[t - width/2 + j]); mode 0 interp (edge_left / edge_right: (width/2, width) operators applied to the first / last
 * `width` samples), 1 nearest, 2 mirror, 3 constant (cval), 4 wrap -- scipy's boundary rules. */
int mlxa_savgol_f32(const float* x, int64_t rows, int64_t T, const float* taps, int width, int mode, float cval,
                    const float* edge_left, const float* edge_right, float* out, void* stream);
/* Per-frame time-domain statistics with the framing done by index arithmetic (centre padding constant or
 * edge): kind 0 RMS (framing.py:81-151), kind 1 zero-crossing rate (features.py:594-720).
 * out (B, T), T = 1 + (L + 2*pad - frame_length) / hop. */
int mlxa_frame_stats_f32(const float* y, int64_t B, int64_t L, int64_t ldy, int frame_length, int hop,
                         int center, int pad_mode, int kind, float* out, void* stream);
/* y[n] - coef*y[n-1] with out[0] = y[0] + zi[b] (zi NULL: 2 y[0] - y[1]); zf (optional, B) = y[L-1]
 * (framing.py:154-295). */
int mlxa_preemphasis_f32(const float* y, int64_t B, int64_t L, int64_t ldy, float coef, const float* zi,
                         float* out, float* zf, void* stream);

/* ---- host-buffer convenience (the e2e path: pinned or pageable HOST pointers) ------------ */
/* log-mel of host clips with chunked H2D / compute / D2H overlap on internal streams.
 * y_host (B, L), out_host (B, n_bands, T), bank_host: packed filterbank in HOST memory.  The dB step is power_to_db(ref, amin, top_db)
 * with the max taken over the whole batch; ref_is_max != 0 means ref = max(mel).  xchg (optional): this
 * rank's clips are a shard of the batch -- the rank's peak is published once after the last chunk and the
 * dB kernels use the maximum over all ranks (epoch advanced by the caller, once per call, on every rank).
 * Synchronous: returns when out_host is complete. */
int mlxa_logmel_host_f32(const float* y_host, int64_t B, int64_t L, const float* window_host,
                         int n_fft, int hop, int center, int pad_mode, float power,
                         const float* bank_host, int n_bands, int64_t n_wt,
                         int apply_db, int ref_is_max, float ref, float amin,
                         int use_top_db, float top_db, const mlxa_peak_exchange* xchg, float* out_host);

/* Measurement aid: launches blocks x threads threads, each running 8 independent chains of
 * `iters` FFMAs (16*iters flops per thread).  bench.py times it to get the FP32 CUDA-core peak
 * of the box the roofline fractions are quoted against. */
int mlxa_ffma_probe(float* scratch, int blocks, int threads, int iters, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MLXA_CUDA_H */
