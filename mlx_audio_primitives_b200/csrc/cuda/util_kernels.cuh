// Host launch wrappers implemented in util_kernels.cu
#pragma once
#include "params.cuh"

namespace mlxa {
cudaError_t run_pad(const float* x, long long B, int L, int pad, int mode, float* out, cudaStream_t s);
cudaError_t run_frame(const float* x, long long B, long long L, int fl, int hop, long long T, float* out, cudaStream_t s);
cudaError_t run_wss(const float* w, int n_fft, int hop, long long T, long long out_len, float* wss, cudaStream_t s);
cudaError_t run_ola(const float* frames, const float* w, long long B, long long T, int n_fft, int hop,
                    long long ola_len, long long trim, long long out_len, long long ldy, float* y, cudaStream_t s);
cudaError_t run_magnitude(const float2* z, long long n, float* out, cudaStream_t s);
cudaError_t run_phase(const float2* z, long long n, float* out, cudaStream_t s);
cudaError_t run_polar(const float* mag, const float* ang, long long n, float2* out, cudaStream_t s);
cudaError_t run_fill(float* x, long long n, float v, cudaStream_t s);
cudaError_t run_momentum(const float* u, const float* u_prev, float m, long long B, long long n, long long ld, float* y,
                         cudaStream_t s);
cudaError_t run_transpose_f32(const float* in, long long B, long long R, long long C, float* out, cudaStream_t s);
cudaError_t run_transpose_c64(const float2* in, long long B, long long R, long long C, float2* out, cudaStream_t s);
cudaError_t run_max(const float* x, long long n, float* gmax, cudaStream_t s);
cudaError_t run_peak_publish(const PeakExchange& xchg, float* gmax, cudaStream_t s);
cudaError_t run_to_db(const float* x, long long n, float coef, float amin, float ref_host, const float* ref_dev,
                      int use_top, float top_db, const float* gmax, float* out, float* reset_next, const PeakExchange& xchg,
                      cudaStream_t s);
cudaError_t run_db_floor(float* x, long long n, float coef, float amin, float ref, float top_db, const float* gmax,
                         float* reset_next, cudaStream_t s);
cudaError_t run_db_floor_blocks(float* x, long long B, int n_bands, long long T, float coef, float amin, float ref,
                                float top_db, const float* gmax, float* block_min, float* reset_next, int* n_raised,
                                const PeakExchange& xchg, float* host_mirror, cudaStream_t s);
cudaError_t run_from_db(const float* x, long long n, float ref, float div, float* out, cudaStream_t s);
cudaError_t run_dct(const float* x, long long rows, int n_in, const float* D, int n_out, float* out, cudaStream_t s);
cudaError_t run_mfcc_tail(const float* mel, long long B, int n_mels, long long T, const float* D, int n_mfcc,
                          const float* lifter, int apply_db, float amin, float ref, int use_top, float top_db,
                          const float* gmax, float* out, cudaStream_t s);
cudaError_t run_pcg64_polar(unsigned long long s_hi, unsigned long long s_lo, unsigned long long i_hi, unsigned long long i_lo,
                            double low, double range, const float* mag, long long B, long long F, long long T, float2* out,
                            cudaStream_t s);
cudaError_t run_pcg64_uniform(unsigned long long s_hi, unsigned long long s_lo, unsigned long long i_hi,
                              unsigned long long i_lo, double low, double range, long long n, float* out, cudaStream_t s);
// feat_kernels.cu
cudaError_t run_spectral_stats(const void* S, int is_complex, long long rows, int F, const float* freq, int kind, float p1,
                               float p2, int norm, const float* centroid_in, float* out, cudaStream_t s);
cudaError_t run_spectral_contrast(const void* S, int is_complex, long long B, long long T, int F, const int* bands, int n_out,
                                  int linear, float* out, cudaStream_t s);
cudaError_t run_savgol(const float* x, long long rows, long long T, const float* taps, int width, int mode, float cval,
                       const float* edge_left, const float* edge_right, float* out, cudaStream_t s);
cudaError_t run_resample_poly(const float* x, long long rows, long long n_in, const float* h, int len_h, int up, int down,
                              long long pre_remove, long long n_out, float* out, cudaStream_t s);
cudaError_t run_resample_linear(const float* x, long long rows, long long n_in, long long n_out, double step, double gain,
                                int apply_gain, float* out, cudaStream_t s);
cudaError_t run_autocorrelation(const float* y, long long B, long long n, long long ldy, int max_lag, int normalize, int center,
                                float* out, float* scratch, cudaStream_t s);
long long autocorr_fft_work_bytes(long long B, long long n);
cudaError_t run_autocorr_fft(const float* y, long long B, long long n, long long ldy, long long max_lag, const float* mean, float* out,
                             void* work, cudaStream_t s);
cudaError_t run_autocorr_prologue(const float* y, long long B, long long n, long long ldy, float* mean, cudaStream_t s);
cudaError_t run_autocorr_epilogue(float* out, long long B, int max_lag, float* r0, cudaStream_t s);
long long resample_fft_work_bytes(long long B, long long n, long long num);
cudaError_t run_resample_fft(const float* x, long long B, long long n, long long ldx, long long num, float gain, float* out,
                             long long ldo, void* work, cudaStream_t s);
cudaError_t run_deemphasis(const float* y, long long B, long long n, long long ldy, double coef, const float* zi, int librosa_zi,
                           float* out, long long ldo, float* zf, cudaStream_t s);
cudaError_t run_frame_stats(const float* y, long long B, int L, long long ldy, int frame_length, int hop, int pad, int pad_mode,
                            long long T, int kind, float* out, cudaStream_t s);
cudaError_t run_preemphasis(const float* y, long long B, long long L, long long ldy, float coef, const float* zi, float* out,
                            float* zf, cudaStream_t s);
}  // namespace mlxa
