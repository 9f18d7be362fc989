// Per-frame reductions either side of the spectral hot path: spectral
// centroid / bandwidth / rolloff / flatness over the F bins of every frame (reference features.py:57-442),
// RMS and zero-crossing rate over the samples of every frame (framing.py:81, features.py:594-720), and
// pre-emphasis (framing.py:154-295).  One warp per frame; the spectral kernel reads the PHYSICAL (B, T, F)
// rows the fused STFT produces (complex: |X| is formed on load, so no magnitude pass and no transposed copy)
// or a real (B, T, F) array.
#include <cmath>
#include "fwd_epilogue.cuh"
#include "util_kernels.cuh"

namespace mlxa {
namespace {

constexpr unsigned kFull = 0xffffffffu;
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
// inclusive prefix sum over the 32 lanes
__device__ __forceinline__ float warp_scan(float v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_up_sync(kFull, v, o);
        if (lane >= o) v += t;
    }
    return v;
}


// S[row][k]: |X|^power of bin k (power applied only for flatness-from-audio; 1 elsewhere)
template <bool CPLX>
__device__ __forceinline__ float spec_value(const void* base, long long row, int F, int k, float power) {
    float m;
    if constexpr (CPLX) {
        const float2 z = __ldg(reinterpret_cast<const float2*>(base) + row * F + k);
        if (power == 2.0f) return fmaf(z.x, z.x, z.y * z.y);
        m = sqrtf(fmaf(z.x, z.x, z.y * z.y));
    } else {
        m = __ldg(reinterpret_cast<const float*>(base) + row * F + k);
        if (power == 2.0f) return m * m;
    }
    return power == 1.0f ? m : powf(m, power);
}

template <bool CPLX>
__global__ void __launch_bounds__(256) spectral_stats_kernel(const void* __restrict__ S, long long rows, int F,
                                                             const float* __restrict__ freq, int kind, float p1, float p2,
                                                             int norm, const float* __restrict__ centroid_in,
                                                             float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += warps) {
        float res = 0.f;
        if (kind == STAT_CENTROID || kind == STAT_BANDWIDTH) {
            float s0 = 0.f, s1 = 0.f;
            for (int k = lane; k < F; k += 32) {
                const float v = spec_value<CPLX>(S, row, F, k, 1.0f);
                s0 += v;
                s1 = fmaf(__ldg(freq + k), v, s1);
            }
            s0 = warp_sum(s0);
            s1 = warp_sum(s1);
            const float c = s1 / (s0 + 1e-10f);  // features.py:120-134
            if (kind == STAT_CENTROID) {
                res = c;
            } else {  // features.py:226-271: (sum S |f - c|^p / (sum S + 1e-10))^(1/p), or without the division
                const float cc = centroid_in ? __ldg(centroid_in + row) : c;
                float s2 = 0.f;
                for (int k = lane; k < F; k += 32) {
                    const float v = spec_value<CPLX>(S, row, F, k, 1.0f);
                    const float d = fabsf(__ldg(freq + k) - cc);
                    s2 = fmaf(v, (p1 == 2.0f) ? d * d : powf(d, p1), s2);
                }
                s2 = warp_sum(s2);
                const float q = norm ? s2 / (s0 + 1e-10f) : s2;
                res = (p1 == 2.0f) ? sqrtf(q) : powf(q, 1.0f / p1);
            }
        } else if (kind == STAT_ROLLOFF) {  // features.py:242-271: first bin whose cumulative sum reaches p1 * total
            float carry = 0.f;
            for (int k0 = 0; k0 < F; k0 += 32) {
                const int k = k0 + lane;
                const float cs = warp_scan(k < F ? spec_value<CPLX>(S, row, F, k, 1.0f) : 0.f, lane);
                carry += __shfl_sync(kFull, cs, 31);
            }
            const float thr = p1 * carry;  // the same blocked scan below reaches exactly `carry`
            int idx = F - 1;
            carry = 0.f;
            for (int k0 = 0; k0 < F; k0 += 32) {
                const int k = k0 + lane;
                const float cs = carry + warp_scan(k < F ? spec_value<CPLX>(S, row, F, k, 1.0f) : 0.f, lane);
                const unsigned hit = __ballot_sync(kFull, k < F && !(cs < thr));
                if (hit) { idx = k0 + __ffs(hit) - 1; break; }
                carry = __shfl_sync(kFull, cs, 31);
            }
            res = __ldg(freq + idx);
        } else {  // STAT_FLATNESS, features.py:430-442: exp(mean log max(S, amin)) / (mean max(S, amin) + 1e-10)
            float sl = 0.f, sa = 0.f;
            for (int k = lane; k < F; k += 32) {
                const float v = fmaxf(spec_value<CPLX>(S, row, F, k, p1), p2);
                sl += logf(v);
                sa += v;
            }
            sl = warp_sum(sl);
            sa = warp_sum(sa);
            res = expf(sl / float(F)) / (sa / float(F) + 1e-10f);
        }
        if (lane == 0) out[row] = res;
    }
}

// Spectral contrast (features.py:445-592, host NumPy in the reference): per frame and octave band, the mean of the
// nq smallest ("valley") and of the nq largest ("peak") magnitudes of the band's bins.  One warp per frame: the
// frame's magnitudes are parked in the warp's shared-memory slice once, then each band is scanned nq-at-most
// times -- round j extracts the next distinct value beyond the last one taken, together with its multiplicity
// (ties count like in a sort), so no sort and no per-lane arrays are needed and any band width works.
// bands: (lo, n, nq) triples as computed on the host by the reference's edge rules.
template <bool CPLX>
__global__ void __launch_bounds__(256) spectral_contrast_kernel(const void* __restrict__ S, long long B, long long T, int F,
                                                                const int* __restrict__ bands, int n_out, int linear,
                                                                float* __restrict__ out) {
    extern __shared__ float s_mag[];  // [warps][F]
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    float* m = s_mag + (size_t)wi * F;
    const long long warps = (long long)gridDim.x * (blockDim.x >> 5), rows = B * T;
    for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + wi; row < rows; row += warps) {
        __syncwarp();
        for (int k = lane; k < F; k += 32) m[k] = spec_value<CPLX>(S, row, F, k, 1.0f);
        __syncwarp();
        const long long b = row / T, t = row - b * T;
        for (int band = 0; band < n_out; ++band) {
            const int lo = __ldg(bands + 3 * band), n = __ldg(bands + 3 * band + 1), nq = __ldg(bands + 3 * band + 2);
            float res[2] = {0.f, 0.f};  // valley, peak
            if (n > 0) {
#pragma unroll
                for (int top = 0; top < 2; ++top) {
                    float last = top ? INFINITY : -INFINITY, sum = 0.f;
                    int taken = 0;
                    while (taken < nq) {  // warp-uniform
                        float best = top ? -INFINITY : INFINITY;
                        int cnt = 0;
                        for (int i = lane; i < n; i += 32) {
                            const float v = m[lo + i];
                            const bool beyond = top ? (v < last) : (v > last);
                            if (beyond) {
                                if (top ? (v > best) : (v < best)) { best = v; cnt = 1; }
                                else if (v == best) ++cnt;
                            }
                        }
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            const float ob = __shfl_xor_sync(kFull, best, o);
                            const int oc = __shfl_xor_sync(kFull, cnt, o);
                            if (top ? (ob > best) : (ob < best)) { best = ob; cnt = oc; }
                            else if (ob == best) cnt += oc;
                        }
                        if (cnt == 0) break;  // fewer than nq values (NaNs): stop with what there is
                        const int take = min(cnt, nq - taken);
                        sum = fmaf(float(take), best, sum);
                        taken += take;
                        last = best;
                    }
                    res[top] = sum / float(nq);
                }
            }
            float r;
            if (linear) r = res[1] - res[0];
            else r = 10.0f * log10f(fmaxf(res[1], 1e-10f)) - 10.0f * log10f(fmaxf(res[0], 1e-10f));
            if (lane == 0) out[(b * n_out + band) * T + t] = r;
        }
    }
}

// kind 0: sqrt(mean x^2) (framing.py:81-151); kind 1: mean of sign changes, the first sample of a frame never
// counts, sign = (x >= 0) (features.py:594-720).  Frames by index arithmetic on the centre-padded clip.
__global__ void __launch_bounds__(256) frame_stats_kernel(const float* __restrict__ y, long long B, int L, long long ldy,
                                                          int frame_length, int hop, int pad, int pad_mode, long long T,
                                                          int kind, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * (blockDim.x >> 5), rows = B * T;
    for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += warps) {
        const long long b = row / T, t = row - b * T;
        const float* yb = y + b * ldy;
        const int s0 = int(t * hop) - pad;
        float acc = 0.f;
        const bool inside = s0 >= 1 && (long long)s0 + frame_length <= L;  // no padding rule applies (and x[s0 - 1] exists)
        if (kind == 0) {
            if (inside) {  // interior frame: plain coalesced loads, four independent partial sums
                const float* x = yb + s0;
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                int i = lane;
                for (; i + 96 < frame_length; i += 128) {
                    const float x0 = __ldg(x + i), x1 = __ldg(x + i + 32), x2 = __ldg(x + i + 64), x3 = __ldg(x + i + 96);
                    a0 = fmaf(x0, x0, a0); a1 = fmaf(x1, x1, a1); a2 = fmaf(x2, x2, a2); a3 = fmaf(x3, x3, a3);
                }
                for (; i < frame_length; i += 32) {
                    const float x0 = __ldg(x + i);
                    a0 = fmaf(x0, x0, a0);
                }
                acc = (a0 + a1) + (a2 + a3);
            } else {
                for (int i = lane; i < frame_length; i += 32) {
                    const float x = load_padded(yb, L, s0 + i, pad_mode);
                    acc = fmaf(x, x, acc);
                }
            }
            acc = sqrtf(warp_sum(acc) / float(frame_length));
        } else {
            if (inside) {
                const float* x = yb + s0;
                int i = lane + 1;
                for (; i + 32 < frame_length; i += 64) {
                    const float c0 = __ldg(x + i), p0 = __ldg(x + i - 1), c1 = __ldg(x + i + 32), p1 = __ldg(x + i + 31);
                    acc += ((c0 >= 0.f) != (p0 >= 0.f)) ? 1.f : 0.f;
                    acc += ((c1 >= 0.f) != (p1 >= 0.f)) ? 1.f : 0.f;
                }
                for (; i < frame_length; i += 32) acc += ((__ldg(x + i) >= 0.f) != (__ldg(x + i - 1) >= 0.f)) ? 1.f : 0.f;
            } else {
                for (int i = lane + 1; i < frame_length; i += 32) {
                    const bool a = load_padded(yb, L, s0 + i, pad_mode) >= 0.f, c = load_padded(yb, L, s0 + i - 1, pad_mode) >= 0.f;
                    acc += (a != c) ? 1.f : 0.f;
                }
            }
            acc = warp_sum(acc) / float(frame_length);
        }
        if (lane == 0) out[row] = acc;
    }
}

// out[n] = y[n] - coef*y[n-1]; out[0] = y[0] + zi (zi = 2 y[0] - y[1] by default); zf = y[L-1]
__global__ void preemphasis_kernel(const float* __restrict__ y, long long B, long long L, long long ldy, float coef,
                                   const float* __restrict__ zi, float* __restrict__ out, float* __restrict__ zf) {
    // four samples per thread where the row layout allows 16-byte accesses (product rounded first, then the
    // subtraction, like the reference: bit-exact)
    const bool vec = (L % 4 == 0) && (ldy % 4 == 0) && (((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(out)) & 15) == 0);
    const long long L4 = vec ? L / 4 : 0, n4 = B * L4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / L4, j = (i - b * L4) * 4;
        const float* yb = y + b * ldy;
        const float4 c = __ldg(reinterpret_cast<const float4*>(yb + j));
        float4 o;
        if (j == 0) o.x = c.x + (zi ? __ldg(zi + b) : __fsub_rn(__fmul_rn(2.f, c.x), c.y));
        else o.x = __fsub_rn(c.x, __fmul_rn(coef, __ldg(yb + j - 1)));
        o.y = __fsub_rn(c.y, __fmul_rn(coef, c.x));
        o.z = __fsub_rn(c.z, __fmul_rn(coef, c.y));
        o.w = __fsub_rn(c.w, __fmul_rn(coef, c.z));
        *reinterpret_cast<float4*>(out + b * L + j) = o;
        if (zf != nullptr && j + 4 == L) zf[b] = c.w;
    }
    if (vec) return;
    const long long n = B * L;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / L, j = i - b * L;
        const float* yb = y + b * ldy;
        float v;
        if (j == 0) v = yb[0] + (zi ? __ldg(zi + b) : (L > 1 ? __fsub_rn(__fmul_rn(2.f, yb[0]), yb[1]) : yb[0]));
        else v = __fsub_rn(yb[j], __fmul_rn(coef, yb[j - 1]));  // product rounded first, like the reference (bit-exact)
        out[i] = v;
        if (zf != nullptr && j == L - 1) zf[b] = yb[j];
    }
}

// Savitzky-Golay filter along the last axis (reference mfcc.py:290-371 = scipy.signal.savgol_filter on the host):
// out[r, t] = sum_j taps[j] * x[r, t - h + j] with the boundary rule of `mode`; in mode 0 ("interp") the first and
// last h outputs come from the polynomial fitted to the first / last `width` samples, i.e. from the (h, width)
// operators edge_left / edge_right.  taps / operators are computed on the host in float64.
enum : int { SG_INTERP = 0, SG_NEAREST = 1, SG_MIRROR = 2, SG_CONSTANT = 3, SG_WRAP = 4 };
__global__ void savgol_kernel(const float* __restrict__ x, long long rows, long long T, const float* __restrict__ taps,
                              int width, int mode, float cval, const float* __restrict__ edge_left,
                              const float* __restrict__ edge_right, float* __restrict__ out) {
    extern __shared__ float s_taps[];
    for (int i = threadIdx.x; i < width; i += blockDim.x) s_taps[i] = __ldg(taps + i);
    __syncthreads();
    const int h = width / 2;
    const long long n = rows * T;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / T, t = i - r * T;
        const float* xr = x + r * T;
        float acc = 0.f;
        if (mode == SG_INTERP && (t < h || t >= T - h)) {
            const bool left = t < h;
            const float* op = left ? edge_left + t * width : edge_right + (t - (T - h)) * width;
            const float* seg = left ? xr : xr + (T - width);
            for (int j = 0; j < width; ++j) acc = fmaf(__ldg(op + j), seg[j], acc);
        } else if (t >= h && t < T - h) {
            const float* seg = xr + (t - h);
            for (int j = 0; j < width; ++j) acc = fmaf(s_taps[j], seg[j], acc);
        } else {
            for (int j = 0; j < width; ++j) {
                long long q = t - h + j;
                float v;
                if (q >= 0 && q < T) {
                    v = xr[q];
                } else if (mode == SG_CONSTANT) {
                    v = cval;
                } else {
                    if (mode == SG_NEAREST) q = q < 0 ? 0 : T - 1;
                    else if (mode == SG_WRAP) { q %= T; if (q < 0) q += T; }
                    else {  // mirror: d c b | a b c d | c b a (period 2T - 2, the edge sample is not repeated)
                        if (T == 1) q = 0;
                        else {
                            const long long per = 2 * T - 2;
                            q %= per; if (q < 0) q += per;
                            if (q >= T) q = per - q;
                        }
                    }
                    v = xr[q];
                }
                acc = fmaf(s_taps[j], v, acc);
            }
        }
        out[i] = acc;
    }
}

// Polyphase rational resampling (reference resample.py:215-300 = scipy.signal.resample_poly on the host):
// out[r, j] = sum_i x[r, i] * h[(j + pre_remove)*down - i*up], h the zero-padded, up-scaled Kaiser low-pass designed on
// the host; one thread per output sample, the taps it meets are `up` apart.
// I: the integer type of the tap index (j + pre_remove) * down -- 32-bit whenever it fits (64-bit divisions cost ~100
// instructions each); rows ride on grid.y.  (Taps re-laid per phase in shared memory measured slower: 0.37 vs 0.33 ms.)
template <typename I>
__global__ void __launch_bounds__(256) resample_poly_kernel(const float* __restrict__ x, long long rows, long long n_in,
                                                            const float* __restrict__ h, int len_h, int up, int down,
                                                            long long pre_remove, long long n_out, float* __restrict__ out) {
    for (long long r = blockIdx.y; r < rows; r += gridDim.y) {
        const float* xr = x + r * n_in;
        float* orow = out + r * n_out;
        for (long long j = blockIdx.x * 256LL + threadIdx.x; j < n_out; j += 256LL * gridDim.x) {
            const I m = I(j + pre_remove) * I(down);
            I i_hi = m / I(up);
            if (i_hi > I(n_in - 1)) i_hi = I(n_in - 1);
            const I lo_num = m - I(len_h) + 1;
            const I i_lo = lo_num <= 0 ? I(0) : (lo_num + I(up) - 1) / I(up);
            const float* hp = h + (m - i_lo * I(up));
            const float* xp = xr + i_lo;
            float acc = 0.f;
            for (int c = int(i_hi - i_lo); c >= 0; --c) {  // ascending i: the reference's (upfirdn's) order
                acc = fmaf(*xp, __ldg(hp), acc);
                ++xp;
                hp -= up;
            }
            orow[j] = acc;
        }
    }
}
// Linear-interpolation resampling (resample.py:142-212): positions linspace(0, n_in - 1, n_out) and the blend in
// float64 like NumPy, the result rounded to float32 (and scaled by `gain` in float64 first when asked).
__global__ void resample_linear_kernel(const float* __restrict__ x, long long rows, long long n_in, long long n_out, double step,
                                       double gain, int apply_gain, float* __restrict__ out) {
    const long long n = rows * n_out;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx / n_out, j = idx - r * n_out;
        const double t = (j == n_out - 1 && n_out > 1) ? double(n_in - 1) : double(j) * step;
        const long long lo = (long long)floor(t);
        const long long hi = lo + 1 < n_in ? lo + 1 : n_in - 1;
        const double frac = t - double(lo);
        const float* xr = x + r * n_in;
        double v = (1.0 - frac) * double(xr[lo]) + frac * double(xr[hi]);
        if (apply_gain) v *= gain;
        out[idx] = float(v);
    }
}

// Whole-signal autocorrelation r[k] = sum_n y[n] y[n + k], k < max_lag (reference pitch.py:16-116 gets it from ONE
// zero-padded FFT of the entire signal -- up to 2^20 points, not a shared-memory transform).  Here it is the direct sum:
// a CTA owns 128 lags of a clip, stages the centred signal chunk by chunk in shared memory and accumulates every
// chunk in float32 and the chunks in float64, in a fixed order (deterministic, and more accurate than a float32 FFT).
// O(n * max_lag): meant for the lag ranges pitch analysis uses; max_lag = n is served but quadratic.
constexpr int kAcLags = 128, kAcChunk = 2048;
__global__ void __launch_bounds__(256) row_mean_kernel(const float* __restrict__ y, long long n, long long ldy, float* __restrict__ mean) {
    __shared__ double s[256];
    const float* yb = y + (long long)blockIdx.x * ldy;
    double acc = 0.0;
    for (long long i = threadIdx.x; i < n; i += 256) acc += double(yb[i]);
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) mean[blockIdx.x] = float(s[0] / double(n));
}
__global__ void __launch_bounds__(256) autocorr_kernel(const float* __restrict__ y, long long n, long long ldy, int max_lag,
                                                       const float* __restrict__ mean, float* __restrict__ out,
                                                       float* __restrict__ r0) {
    __shared__ __align__(16) float s_y[kAcChunk], s_p[kAcChunk + kAcLags + 4];
    __shared__ double s_part[8][kAcLags];
    const long long b = blockIdx.y;
    // a thread owns 4 consecutive lags and one eighth of every chunk: per 4 samples one broadcast 128-bit read of the
    // chunk and one 128-bit read of the partners' window feed 16 FMAs (the window slides through registers)
    const int k0 = blockIdx.x * kAcLags, kl = 4 * (threadIdx.x & 31), part_i = threadIdx.x >> 5;
    const float* yb = y + b * ldy;
    const float mu = mean ? __ldg(mean + b) : 0.f;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    // lag k = k0 + kl + l pairs y[n] with y[n + k]: s_y holds the chunk, s_p the partners' window starting k0 later
    for (long long n0 = 0; n0 + k0 < n; n0 += kAcChunk) {
        __syncthreads();
        for (int i = threadIdx.x; i < kAcChunk + kAcLags + 4; i += 256) {
            const long long q = n0 + i, qp = q + k0;
            if (i < kAcChunk) s_y[i] = (q < n) ? yb[q] - mu : 0.f;
            s_p[i] = (qp < n) ? yb[qp] - mu : 0.f;
        }
        __syncthreads();
        const int i_lo = part_i * (kAcChunk / 8);
        float part[4] = {0.f, 0.f, 0.f, 0.f};
        float4 pc = *reinterpret_cast<const float4*>(s_p + i_lo + kl);
#pragma unroll 4
        for (int i = i_lo; i < i_lo + kAcChunk / 8; i += 4) {
            const float4 y4 = *reinterpret_cast<const float4*>(s_y + i);
            const float4 pn = *reinterpret_cast<const float4*>(s_p + i + kl + 4);
            const float P[8] = {pc.x, pc.y, pc.z, pc.w, pn.x, pn.y, pn.z, pn.w};
            const float Y[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int l = 0; l < 4; ++l) part[l] = fmaf(Y[a], P[a + l], part[l]);
            pc = pn;
        }
#pragma unroll
        for (int l = 0; l < 4; ++l) acc[l] += double(part[l]);
    }
#pragma unroll
    for (int l = 0; l < 4; ++l) s_part[part_i][kl + l] = acc[l];
    __syncthreads();
    if (threadIdx.x < kAcLags) {
        const int k = k0 + threadIdx.x;
        double r = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) r += s_part[q][threadIdx.x];
        if (k < max_lag) out[b * max_lag + k] = float(r);
        if (k == 0) r0[b] = float(r);
    }
}
__global__ void autocorr_normalize_kernel(float* __restrict__ out, long long B, int max_lag, const float* __restrict__ r0) {
    const long long n = B * max_lag;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = out[i] / fmaxf(__ldg(r0 + i / max_lag), 1e-10f);
}

// De-emphasis out[n] = y[n] + coef * out[n-1] (framing.py:298-392; scipy.signal.lfilter on the host in the reference): a
// first-order recurrence, so a scan of affine maps.  One CTA per clip walks the clip in blocks of 1024 x 8 samples: a
// thread runs its 8 samples from a zero state, the states are combined by a shuffle scan (factor coef^(8 d) at distance d)
// and the carried state z = coef * out[n-1] enters sample j of a segment as coef^j * z.  float64 inside, rounded once.
constexpr int kDeSeg = 8, kDeThreads = 1024, kDeBlock = kDeSeg * kDeThreads;
__global__ void __launch_bounds__(kDeThreads) deemphasis_kernel(const float* __restrict__ y, long long n, long long ldy, double coef,
                                                                const float* __restrict__ zi, int librosa_zi,
                                                                float* __restrict__ out, long long ldo, float* __restrict__ zf) {
    __shared__ float s_v[kDeBlock + kDeBlock / 32];
    __shared__ double s_w[32];
    __shared__ double s_carry;
    const long long b = blockIdx.x;
    const float* yb = y + b * ldy;
    float* ob = out + b * ldo;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    double cj[kDeSeg + 1];  // coef^j
    cj[0] = 1.0;
#pragma unroll
    for (int j = 1; j <= kDeSeg; ++j) cj[j] = cj[j - 1] * coef;
    double cd[5], cw[5];  // coef^(8 d) for d = 1, 2, 4, 8, 16 lanes; coef^(256 d) for d warps
    cd[0] = cj[kDeSeg];
#pragma unroll
    for (int i = 1; i < 5; ++i) cd[i] = cd[i - 1] * cd[i - 1];
    cw[0] = cd[4] * cd[4];
#pragma unroll
    for (int i = 1; i < 5; ++i) cw[i] = cw[i - 1] * cw[i - 1];
    // the state in front of sample 0: the caller's zi, or -corr for the reference's default (framing.py:368-381), which
    // subtracts corr * coef^n with corr = ((2 - c) y[0] - y[1]) / (3 - c) evaluated in float32 like NumPy does
    float corr = 0.f;
    if (librosa_zi) {
        corr = __fdiv_rn(__fsub_rn(__fmul_rn(float(2.0 - coef), yb[0]), yb[1]), float(3.0 - coef));
    }
    if (t == 0) s_carry = librosa_zi ? -double(corr) : (zi ? double(zi[b]) : 0.0);
    auto pad = [](int i) { return i + (i >> 5); };
    for (long long n0 = 0; n0 < n; n0 += kDeBlock) {
        __syncthreads();
        for (int i = t; i < kDeBlock; i += kDeThreads) s_v[pad(i)] = (n0 + i < n) ? yb[n0 + i] : 0.f;
        __syncthreads();
        double loc[kDeSeg];
        double z = 0.0;
#pragma unroll
        for (int j = 0; j < kDeSeg; ++j) {
            const double o = double(s_v[pad(t * kDeSeg + j)]) + z;
            loc[j] = o;
            z = coef * o;
        }
        // inclusive scan of the end states: Z_t = z_t + coef^8 Z_{t-1}
        double inc = z;
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const double up = __shfl_up_sync(0xffffffffu, inc, 1 << i);
            if (lane >= (1 << i)) inc = fma(cd[i], up, inc);
        }
        if (lane == 31) s_w[warp] = inc;
        const double carry_in = s_carry;
        __syncthreads();
        if (warp == 0) {
            double w = s_w[lane];
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                const double up = __shfl_up_sync(0xffffffffu, w, 1 << i);
                if (lane >= (1 << i)) w = fma(cw[i], up, w);
            }
            s_w[lane] = w;
        }
        __syncthreads();
        // state in front of this thread's segment: the lane before, the warps before, the blocks before
        double zin = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) zin = 0.0;
        double cl = 1.0;  // coef^(8 lane)
#pragma unroll
        for (int i = 0; i < 5; ++i) if (lane & (1 << i)) cl *= cd[i];
        double cwp = 1.0;  // coef^(256 warp)
#pragma unroll
        for (int i = 0; i < 5; ++i) if (warp & (1 << i)) cwp *= cw[i];
        if (warp > 0) zin = fma(cl, s_w[warp - 1], zin);
        zin = fma(cl * cwp, carry_in, zin);
#pragma unroll
        for (int j = 0; j < kDeSeg; ++j) s_v[pad(t * kDeSeg + j)] = float(fma(cj[j], zin, loc[j]));
        __syncthreads();
        for (int i = t; i < kDeBlock; i += kDeThreads) if (n0 + i < n) ob[n0 + i] = s_v[pad(i)];
        if (t == kDeThreads - 1) s_carry = fma(cj[kDeSeg], zin, z);  // state after the block (zero samples past the end carry it on scaled)
    }
    __syncthreads();
    if (t == 0 && zf != nullptr) {
        // after the last REAL sample: the carried state was advanced over the zero tail of the last block -> undo by
        // recomputing from the last output instead: z = coef * out[n-1]; the reference's default path reports the state
        // of the uncorrected filter: add back corr * coef^n
        double zl = coef * double(ob[n - 1]);
        if (librosa_zi) zl += double(corr) * pow(coef, double(n));
        zf[b] = float(zl);
    }
}

unsigned grid_for_rows(long long rows, int per_cta) {
    const long long g = (rows + per_cta - 1) / per_cta;
    return (unsigned)(g < 1 ? 1 : (g > 148LL * 32 ? 148LL * 32 : g));
}
}  // namespace

cudaError_t run_spectral_stats(const void* S, int is_complex, long long rows, int F, const float* freq, int kind, float p1,
                               float p2, int norm, const float* centroid_in, float* out, cudaStream_t s) {
    const unsigned grid = grid_for_rows(rows, 8);
    if (is_complex) spectral_stats_kernel<true><<<grid, 256, 0, s>>>(S, rows, F, freq, kind, p1, p2, norm, centroid_in, out);
    else spectral_stats_kernel<false><<<grid, 256, 0, s>>>(S, rows, F, freq, kind, p1, p2, norm, centroid_in, out);
    return cudaGetLastError();
}
cudaError_t run_spectral_contrast(const void* S, int is_complex, long long B, long long T, int F, const int* bands, int n_out,
                                  int linear, float* out, cudaStream_t s) {
    int warps = 8;
    while (warps > 1 && (size_t)warps * F * 4 > 96 * 1024) warps >>= 1;
    const size_t smem = (size_t)warps * F * 4;
    if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
    const unsigned grid = grid_for_rows(B * T, warps);
    cudaError_t e;
    if (is_complex) {
        e = cudaFuncSetAttribute(spectral_contrast_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        spectral_contrast_kernel<true><<<grid, warps * 32, smem, s>>>(S, B, T, F, bands, n_out, linear, out);
    } else {
        e = cudaFuncSetAttribute(spectral_contrast_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        spectral_contrast_kernel<false><<<grid, warps * 32, smem, s>>>(S, B, T, F, bands, n_out, linear, out);
    }
    return cudaGetLastError();
}
cudaError_t run_savgol(const float* x, long long rows, long long T, const float* taps, int width, int mode, float cval,
                       const float* edge_left, const float* edge_right, float* out, cudaStream_t s) {
    const long long n = rows * T, g = (n + 255) / 256;
    savgol_kernel<<<(unsigned)(g < 1 ? 1 : (g > 148LL * 64 ? 148LL * 64 : g)), 256, (size_t)width * 4, s>>>(
        x, rows, T, taps, width, mode, cval, edge_left, edge_right, out);
    return cudaGetLastError();
}
cudaError_t run_resample_poly(const float* x, long long rows, long long n_in, const float* h, int len_h, int up, int down,
                              long long pre_remove, long long n_out, float* out, cudaStream_t s) {
    const long long gx = (n_out + 255) / 256;
    dim3 grid((unsigned)(gx < 1 ? 1 : (gx > 148LL * 64 ? 148LL * 64 : gx)), (unsigned)(rows > 65535 ? 65535 : rows));
    const bool small = (n_out + pre_remove) * (long long)down < (1LL << 31) && n_in < (1LL << 31);
    if (small) resample_poly_kernel<int><<<grid, 256, 0, s>>>(x, rows, n_in, h, len_h, up, down, pre_remove, n_out, out);
    else resample_poly_kernel<long long><<<grid, 256, 0, s>>>(x, rows, n_in, h, len_h, up, down, pre_remove, n_out, out);
    return cudaGetLastError();
}
cudaError_t run_resample_linear(const float* x, long long rows, long long n_in, long long n_out, double step, double gain,
                                int apply_gain, float* out, cudaStream_t s) {
    const long long n = rows * n_out, g = (n + 255) / 256;
    resample_linear_kernel<<<(unsigned)(g < 1 ? 1 : (g > 148LL * 64 ? 148LL * 64 : g)), 256, 0, s>>>(x, rows, n_in, n_out, step, gain,
                                                                                                   apply_gain, out);
    return cudaGetLastError();
}
__global__ void autocorr_r0_kernel(const float* __restrict__ out, long long B, int max_lag, float* __restrict__ r0) {
    const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b < B) r0[b] = out[b * max_lag];
}
cudaError_t run_autocorr_prologue(const float* y, long long B, long long n, long long ldy, float* mean, cudaStream_t s) {
    row_mean_kernel<<<(unsigned)B, 256, 0, s>>>(y, n, ldy, mean);
    return cudaGetLastError();
}
cudaError_t run_autocorr_epilogue(float* out, long long B, int max_lag, float* r0, cudaStream_t s) {
    autocorr_r0_kernel<<<(unsigned)((B + 255) / 256), 256, 0, s>>>(out, B, max_lag, r0);
    const long long tot = B * max_lag, g = (tot + 255) / 256;
    autocorr_normalize_kernel<<<(unsigned)(g > 148LL * 32 ? 148LL * 32 : g), 256, 0, s>>>(out, B, max_lag, r0);
    return cudaGetLastError();
}
cudaError_t run_autocorrelation(const float* y, long long B, long long n, long long ldy, int max_lag, int normalize, int center,
                                float* out, float* scratch, cudaStream_t s) {
    float* mean = scratch;
    float* r0 = scratch + B;
    if (center) row_mean_kernel<<<(unsigned)B, 256, 0, s>>>(y, n, ldy, mean);
    dim3 grid((unsigned)((max_lag + kAcLags - 1) / kAcLags), (unsigned)B);
    autocorr_kernel<<<grid, 256, 0, s>>>(y, n, ldy, max_lag, center ? mean : nullptr, out, r0);
    if (normalize) {
        const long long tot = B * max_lag, g = (tot + 255) / 256;
        autocorr_normalize_kernel<<<(unsigned)(g > 148LL * 32 ? 148LL * 32 : g), 256, 0, s>>>(out, B, max_lag, r0);
    }
    return cudaGetLastError();
}
cudaError_t run_deemphasis(const float* y, long long B, long long n, long long ldy, double coef, const float* zi, int librosa_zi,
                           float* out, long long ldo, float* zf, cudaStream_t s) {
    deemphasis_kernel<<<(unsigned)B, kDeThreads, 0, s>>>(y, n, ldy, coef, zi, librosa_zi, out, ldo, zf);
    return cudaGetLastError();
}
cudaError_t run_frame_stats(const float* y, long long B, int L, long long ldy, int frame_length, int hop, int pad, int pad_mode,
                            long long T, int kind, float* out, cudaStream_t s) {
    frame_stats_kernel<<<grid_for_rows(B * T, 8), 256, 0, s>>>(y, B, L, ldy, frame_length, hop, pad, pad_mode, T, kind, out);
    return cudaGetLastError();
}
cudaError_t run_preemphasis(const float* y, long long B, long long L, long long ldy, float coef, const float* zi, float* out,
                            float* zf, cudaStream_t s) {
    const long long n = B * L;
    const long long g = (n + 255) / 256;
    preemphasis_kernel<<<(unsigned)(g > 148LL * 64 ? 148LL * 64 : g), 256, 0, s>>>(y, B, L, ldy, coef, zi, out, zf);
    return cudaGetLastError();
}

}  // namespace mlxa
