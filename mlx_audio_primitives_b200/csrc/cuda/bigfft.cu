// Fourier-method resampling of whole signals (reference resample.py:84-139 calls scipy.signal.resample on the host:
// X = rfft(x); keep / zero-extend to the new length; irfft).  The signal length n and the new length are arbitrary
// integers far beyond a shared-memory transform, so both DFTs are Bluestein chirp-z transforms over power-of-two
// FFTs that live in global memory:
//
//   DFT_n(x)[k]  = conj(w_n[k]) * sum_j (x[j] conj(w_n[j])) w_n[k - j],      w_n[j] = exp(i pi j^2 / n)
//   IDFT_m(Z)[j] = (1/m) w_m[j] * sum_k (Z[k] w_m[k]) conj(w_m[j - k])
//
// Each convolution is FFT_M -> pointwise product with the cached transform of the chirp -> FFT_M again (the inverse
// through conjugation, folded into the pointwise kernels), M the power of two >= 2 n - 1.  FFT_M is a Stockham
// autosort sequence of radix-16 passes over global memory (out of place, ping-pong; one final radix-2/4/8 pass when
// log2 M is not a multiple of 4), twiddles from a table evaluated in float64.  Chirp phases use j^2 mod 2n in
// integer arithmetic, so they stay exact at any length.  Everything is batched over clips (grid.y).
#include <cuda_runtime.h>

#include <cstdint>
#include <map>
#include <mutex>
#include <tuple>
#include <utility>

#include "common.cuh"
#include "fft_radix.cuh"
#include "util_kernels.cuh"

namespace mlxa {
namespace {

__global__ void root_table_kernel(float2* __restrict__ w, long long M) {  // W_M^k = exp(-2 pi i k / M)
    const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (k >= M) return;
    double s, c;
    sincospi(-2.0 * double(k) / double(M), &s, &c);
    w[k] = make_float2(float(c), float(s));
}
__global__ void chirp_table_kernel(float2* __restrict__ w, long long n) {  // w_n[j] = exp(i pi j^2 / n)
    const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j >= n) return;
    const unsigned long long q = (unsigned long long)j * (unsigned long long)j % (unsigned long long)(2 * n);
    double s, c;
    sincospi(double(q) / double(n), &s, &c);
    w[j] = make_float2(float(c), float(s));
}
// the chirp wrapped onto a circle of M points: b[i] = b[M - i] = w (or conj w) for i < n, zero between
__global__ void chirp_wrap_kernel(const float2* __restrict__ w, long long n, long long M, int conj, float2* __restrict__ b) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= M) return;
    const long long j = (i < n) ? i : ((M - i < n) ? M - i : -1);
    float2 v = make_float2(0.f, 0.f);
    if (j >= 0) {
        v = w[j];
        if (conj) v.y = -v.y;
    }
    b[i] = v;
}

// One Stockham pass: thread j of M / R combines in[j + r M/R], r < R, twiddled by W_{Ns R}^{k r} (k = j mod Ns), into
// out[(j - k) R + k + q Ns], q < R.  Twiddles: three table reads (powers 1, 4, 16 of the step) and at most two products
// each -- a table read per leg would double the load traffic, powers above 3 of one rounded root would triple its phase
// error.  MUL fuses the convolution's pointwise product into the transform's last pass: out = conj(out * bhat).
struct PassIO {
    const float* x = nullptr;          // IN_REAL: rows of n reals (stride ldx), multiplied by conj(chirp_in), zero up to M
    long long n = 0, ldx = 0;
    const float2* chirp_in = nullptr;  // (nullptr: the plain signal, minus mean[row] when mean is given)
    const float* mean = nullptr;
    const float2* bhat = nullptr;      // OUT_MUL
    float* y = nullptr;                // OUT_REAL: y[j] = Re(chirp_out[j] conj(result[j])) * gain for j < num (stride ldo)
    long long num = 0, ldo = 0;
    const float2* chirp_out = nullptr; // (nullptr: y[j] = Re(result[j]) * gain)
    float gain = 0.f;
};
enum { IN_PLAIN = 0, IN_REAL = 1, OUT_PLAIN = 0, OUT_MUL = 1, OUT_REAL = 2, OUT_POWER = 3 };

template <int R, int IN, int OUT>
__global__ void __launch_bounds__(256) bigfft_pass_kernel(const float2* __restrict__ in, float2* __restrict__ out, long long M,
                                                          long long Ns, const float2* __restrict__ W, const PassIO io) {
    const long long j = blockIdx.x * 256LL + threadIdx.x;
    const long long per = M / R;
    if (j >= per) return;
    in += blockIdx.y * M;
    out += blockIdx.y * M;
    const long long k = j & (Ns - 1);
    const long long tstep = k * (M / (Ns * R));
    float2 v[R];
    if constexpr (IN == IN_REAL) {  // the chirp-premultiplied, zero-extended real signal: the padding is never read
        const float* xr = io.x + blockIdx.y * io.ldx;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const long long idx = j + r * per;
            v[r] = make_float2(0.f, 0.f);
            if (idx < io.n) {
                if (io.chirp_in != nullptr) {
                    const float xv = xr[idx];
                    const float2 c = __ldg(io.chirp_in + idx);
                    v[r] = make_float2(xv * c.x, -xv * c.y);
                } else {
                    v[r].x = xr[idx] - (io.mean ? __ldg(io.mean + blockIdx.y) : 0.f);
                }
            }
        }
    } else {
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = in[j + r * per];
    }
    if (Ns > 1) {
        float2 p1[4], p4[4], p16[2];
        p1[1] = __ldg(W + tstep);
        p1[2] = cmul(p1[1], p1[1]);
        p1[3] = cmul(p1[2], p1[1]);
        if constexpr (R > 4) {
            p4[1] = __ldg(W + 4 * tstep);
            p4[2] = cmul(p4[1], p4[1]);
            p4[3] = cmul(p4[2], p4[1]);
        }
        if constexpr (R > 16) p16[1] = __ldg(W + 16 * tstep);
        static_for<R>([&](auto r_) {
            constexpr int r = decltype(r_)::value;
            constexpr int c = r & 3, b = (r >> 2) & 3, a = r >> 4;
            if constexpr (r > 0) {
                float2 t;
                if constexpr (c > 0) {
                    t = p1[c];
                    if constexpr (b > 0) t = cmul(t, p4[b]);
                    if constexpr (a > 0) t = cmul(t, p16[a]);
                } else if constexpr (b > 0) {
                    t = p4[b];
                    if constexpr (a > 0) t = cmul(t, p16[a]);
                } else {
                    t = p16[a];
                }
                v[r] = cmul(v[r], t);
            }
        });
    }
    DftInplace<R, 1, 0>::run(v);
    const long long o0 = (j - k) * R + k;
    static_for<R>([&](auto i_) {
        constexpr int i = decltype(i_)::value;
        const long long idx = o0 + dft_perm(R, i) * Ns;
        float2 r = v[i];
        if constexpr (OUT == OUT_REAL) {
            if (idx < io.num) {
                float val = r.x;
                if (io.chirp_out != nullptr) {
                    const float2 w = __ldg(io.chirp_out + idx);
                    val = w.x * r.x + w.y * r.y;
                }
                io.y[blockIdx.y * io.ldo + idx] = val * io.gain;
            }
        } else if constexpr (OUT == OUT_POWER) {
            out[idx] = make_float2(r.x * r.x + r.y * r.y, 0.f);
        } else {
            if constexpr (OUT == OUT_MUL) {
                r = cmul(r, __ldg(io.bhat + idx));
                r.y = -r.y;
            }
            out[idx] = r;
        }
    });
}

template <int R>
void bigfft_launch(dim3 grid, cudaStream_t s, const float2* src, float2* dst, long long M, long long Ns, const float2* W, int in_mode,
                   int out_mode, const PassIO& io) {
#define MLXA_PASS(I, O) bigfft_pass_kernel<R, I, O><<<grid, 256, 0, s>>>(src, dst, M, Ns, W, io)
    if (in_mode == IN_REAL) {
        if (out_mode == OUT_MUL) MLXA_PASS(IN_REAL, OUT_MUL);
        else if (out_mode == OUT_REAL) MLXA_PASS(IN_REAL, OUT_REAL);
        else if (out_mode == OUT_POWER) MLXA_PASS(IN_REAL, OUT_POWER);
        else MLXA_PASS(IN_REAL, OUT_PLAIN);
    } else {
        if (out_mode == OUT_MUL) MLXA_PASS(IN_PLAIN, OUT_MUL);
        else if (out_mode == OUT_REAL) MLXA_PASS(IN_PLAIN, OUT_REAL);
        else if (out_mode == OUT_POWER) MLXA_PASS(IN_PLAIN, OUT_POWER);
        else MLXA_PASS(IN_PLAIN, OUT_PLAIN);
    }
#undef MLXA_PASS
}

// FFT_M of B rows, ping-pong a -> tmp -> a ...; *result is the buffer holding the transform.  log2 M splits into
// ceil(log2 M / 5) passes of near-equal radix (<= 32).  in_mode applies to the first pass (IN_REAL: `a` is not read),
// out_mode to the last (OUT_MUL stores conj(result * bhat); OUT_REAL stores only the real output rows).
cudaError_t bigfft(float2* a, float2* tmp, long long M, long long B, const float2* W, int in_mode, int out_mode, const PassIO& io,
                   cudaStream_t s, float2** result) {
    float2* src = a;
    float2* dst = tmp;
    int bits = 0;
    while ((1LL << bits) < M) ++bits;
    const int passes = bits == 0 ? 1 : (bits + 4) / 5;
    long long Ns = 1;
    for (int p = 0; p < passes; ++p) {
        const int rb = bits == 0 ? 0 : bits / passes + (p < bits % passes ? 1 : 0);
        const int R = 1 << rb;
        const int im = (p == 0) ? in_mode : IN_PLAIN, om = (p == passes - 1) ? out_mode : OUT_PLAIN;
        dim3 grid((unsigned)((M / R + 255) / 256), (unsigned)B);
        switch (R) {
            case 32: bigfft_launch<32>(grid, s, src, dst, M, Ns, W, im, om, io); break;
            case 16: bigfft_launch<16>(grid, s, src, dst, M, Ns, W, im, om, io); break;
            case 8: bigfft_launch<8>(grid, s, src, dst, M, Ns, W, im, om, io); break;
            case 4: bigfft_launch<4>(grid, s, src, dst, M, Ns, W, im, om, io); break;
            case 2: bigfft_launch<2>(grid, s, src, dst, M, Ns, W, im, om, io); break;
            default: bigfft_launch<1>(grid, s, src, dst, M, Ns, W, im, om, io); break;
        }
        Ns *= R;
        float2* t = src; src = dst; dst = t;
    }
    *result = src;
    return cudaGetLastError();
}

// c holds conj(M1 * convolution); X[k] = conj(w_n[k]) conv[k].  Build the Hermitian spectrum Z of the new length num
// (scipy.signal.resample: bins below m2 = min(n, num) / 2 + 1 kept, the unpaired bin at m / 2 doubled when shrinking /
// halved when growing, irfft semantics: imaginary parts of bin 0 and of the Nyquist bin dropped) and write the second
// convolution's input a2[k] = Z[k] w_num[k], zero up to M2.
__global__ void rs_mid_kernel(const float2* __restrict__ c, long long M1, long long n, const float2* __restrict__ wn, long long num,
                              const float2* __restrict__ wm, long long M2, float2* __restrict__ a2) {
    const long long k = blockIdx.x * 256LL + threadIdx.x;
    if (k >= M2) return;
    const long long b = blockIdx.y;
    float2 v = make_float2(0.f, 0.f);
    if (k < num) {
        const long long kk = (2 * k <= num) ? k : num - k;  // the one-sided bin this entry mirrors
        const long long m = n < num ? n : num, m2 = m / 2 + 1;
        if (kk < m2) {
            const float2 cv = c[b * M1 + kk];
            const float2 w = __ldg(wn + kk);
            // conv = conj(cv) / M1; X = conj(w) * conv = conj(w * cv) / M1
            float2 X = cmul(w, cv);
            X.y = -X.y;
            float f = 1.0f / float(M1);
            if (m % 2 == 0 && num != n && kk == m / 2) f *= (num < n) ? 2.f : 0.5f;
            X.x *= f; X.y *= f;
            if (kk == 0 || 2 * kk == num) X.y = 0.f;
            if (kk != k) X.y = -X.y;  // negative frequency: the conjugate
            v = cmul(X, __ldg(wm + k));
        }
    }
    a2[b * M2 + k] = v;
}
// ---- per-device caches: root tables per M, chirps and chirp transforms per length ---------------------------------------
struct Key {
    int dev, kind;
    long long n;
    bool operator<(const Key& o) const { return std::tie(dev, kind, n) < std::tie(o.dev, o.kind, o.n); }
};
std::mutex g_mu;
std::map<Key, float2*> g_cache;
constexpr size_t kMaxEntries = 24;

long long pow2_at_least(long long v) {
    long long m = 1;
    while (m < v) m <<= 1;
    return m;
}

cudaError_t cache_get(int kind, long long n, cudaStream_t s, float2** out);

// Every table is complete (stream synchronised) before its pointer is published in the cache: another stream or
// thread that hits the entry later may read it without an event.  Built once per (device, kind, length).
cudaError_t cache_build(int kind, long long n, cudaStream_t s, float2** out) {
    cudaError_t e;
    *out = nullptr;
    auto fail = [&](cudaError_t err, float2* tmp) {
        if (*out) cudaFree(*out);
        if (tmp) cudaFree(tmp);
        *out = nullptr;
        return err;
    };
    if (kind == 0 || kind == 1) {  // roots of unity of order n / chirp of length n
        if ((e = cudaMalloc(out, size_t(n) * 8)) != cudaSuccess) return e;
        if (kind == 0) root_table_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(*out, n);
        else chirp_table_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(*out, n);
        if ((e = cudaGetLastError()) != cudaSuccess) return fail(e, nullptr);
        if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return fail(e, nullptr);
        return cudaSuccess;
    }
    // kind 2 / 3: FFT_M of the wrapped chirp / of its conjugate
    const long long M = pow2_at_least(2 * n - 1);
    float2 *w = nullptr, *W = nullptr, *tmp = nullptr, *res = nullptr;
    if ((e = cache_get(1, n, s, &w)) != cudaSuccess) return e;
    if ((e = cache_get(0, M, s, &W)) != cudaSuccess) return e;
    if ((e = cudaMalloc(out, size_t(M) * 8)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&tmp, size_t(M) * 8)) != cudaSuccess) return fail(e, nullptr);
    chirp_wrap_kernel<<<(unsigned)((M + 255) / 256), 256, 0, s>>>(w, n, M, kind == 3, *out);
    if ((e = bigfft(*out, tmp, M, 1, W, IN_PLAIN, OUT_PLAIN, PassIO{}, s, &res)) != cudaSuccess) return fail(e, tmp);
    if (res != *out) e = cudaMemcpyAsync(*out, res, size_t(M) * 8, cudaMemcpyDeviceToDevice, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return fail(e, tmp);
    cudaFree(tmp);
    return cudaSuccess;
}

cudaError_t cache_get(int kind, long long n, cudaStream_t s, float2** out) {  // g_mu held by the caller
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const Key key{dev, kind, n};
    auto it = g_cache.find(key);
    if (it != g_cache.end()) { *out = it->second; return cudaSuccess; }
    float2* p = nullptr;
    if ((e = cache_build(kind, n, s, &p)) != cudaSuccess) return e;
    g_cache[key] = p;
    *out = p;
    return cudaSuccess;
}

}  // namespace

long long resample_fft_work_bytes(long long B, long long n, long long num) {
    const long long M1 = pow2_at_least(2 * n - 1), M2 = pow2_at_least(2 * num - 1);
    return 2 * B * (M1 > M2 ? M1 : M2) * 8;
}

cudaError_t run_resample_fft(const float* x, long long B, long long n, long long ldx, long long num, float gain, float* out,
                             long long ldo, void* work, cudaStream_t s) {
    const long long M1 = pow2_at_least(2 * n - 1), M2 = pow2_at_least(2 * num - 1);
    const long long Mx = M1 > M2 ? M1 : M2;
    float2 *W1, *W2, *wn, *wm, *bh1, *bh2;
    // held until every launch that reads the cached tables is enqueued: a flush (below) synchronises the device first, so
    // tables can only be freed when nothing enqueued still reads them
    std::lock_guard<std::mutex> lk(g_mu);
    {
        cudaError_t e;
        if (g_cache.size() + 6 > kMaxEntries) {  // rare (many distinct lengths): drop everything once the device is idle
            if ((e = cudaDeviceSynchronize()) != cudaSuccess) return e;
            for (auto& kv : g_cache) cudaFree(kv.second);
            g_cache.clear();
        }
        if ((e = cache_get(0, M1, s, &W1)) != cudaSuccess) return e;
        if ((e = cache_get(0, M2, s, &W2)) != cudaSuccess) return e;
        if ((e = cache_get(1, n, s, &wn)) != cudaSuccess) return e;
        if ((e = cache_get(1, num, s, &wm)) != cudaSuccess) return e;
        if ((e = cache_get(2, n, s, &bh1)) != cudaSuccess) return e;
        if ((e = cache_get(3, num, s, &bh2)) != cudaSuccess) return e;
    }
    float2* a = static_cast<float2*>(work);
    float2* t = a + B * Mx;
    float2* r = nullptr;
    cudaError_t e;
    const unsigned by = (unsigned)B;
    PassIO io;
    io.x = x; io.n = n; io.ldx = ldx; io.chirp_in = wn; io.bhat = bh1;
    // conj(FFT(x conj(chirp)) * FFT(chirp)), then the transform back: conj(M1 * convolution)
    if ((e = bigfft(a, t, M1, B, W1, IN_REAL, OUT_MUL, io, s, &r)) != cudaSuccess) return e;
    float2* r2 = nullptr;
    if ((e = bigfft(r, (r == a) ? t : a, M1, B, W1, IN_PLAIN, OUT_PLAIN, io, s, &r2)) != cudaSuccess) return e;
    float2* o2 = (r2 == a) ? t : a;
    rs_mid_kernel<<<dim3((unsigned)((M2 + 255) / 256), by), 256, 0, s>>>(r2, M1, n, wn, num, wm, M2, o2);
    PassIO io2;
    io2.bhat = bh2; io2.y = out; io2.num = num; io2.ldo = ldo; io2.chirp_out = wm;
    // y[j] = Re(w_num[j] conv[j]) / n (irfft's 1 / num times scipy's num / n) times the caller's gain; the last transform
    // holds conj(M2 conv)
    io2.gain = gain / (float(M2) * float(n));
    float2* r3 = nullptr;
    if ((e = bigfft(o2, r2, M2, B, W2, IN_PLAIN, OUT_MUL, io2, s, &r3)) != cudaSuccess) return e;
    float2* r4 = nullptr;
    if ((e = bigfft(r3, (r3 == a) ? t : a, M2, B, W2, IN_PLAIN, OUT_REAL, io2, s, &r4)) != cudaSuccess) return e;
    return cudaGetLastError();
}

// Whole-signal autocorrelation through the transforms (pitch.py:16-116 is exactly this): r = FFT_M(|FFT_M(x - mean)|^2) / M
// (the power spectrum is real and even, so the forward transform is its inverse up to 1 / M), M the power of two >= 2n - 1.
long long autocorr_fft_work_bytes(long long B, long long n) { return 2 * B * pow2_at_least(2 * n - 1) * 8; }

cudaError_t run_autocorr_fft(const float* y, long long B, long long n, long long ldy, long long max_lag, const float* mean, float* out,
                             void* work, cudaStream_t s) {
    const long long M = pow2_at_least(2 * n - 1);
    float2* W;
    std::lock_guard<std::mutex> lk(g_mu);  // until the launches are enqueued (see run_resample_fft)
    {
        cudaError_t e;
        if (g_cache.size() + 1 > kMaxEntries) {
            if ((e = cudaDeviceSynchronize()) != cudaSuccess) return e;
            for (auto& kv : g_cache) cudaFree(kv.second);
            g_cache.clear();
        }
        if ((e = cache_get(0, M, s, &W)) != cudaSuccess) return e;
    }
    float2* a = static_cast<float2*>(work);
    float2* t = a + B * M;
    PassIO io;
    io.x = y; io.n = n; io.ldx = ldy; io.mean = mean;
    io.y = out; io.num = max_lag; io.ldo = max_lag; io.gain = 1.0f / float(M);
    float2 *r = nullptr, *r2 = nullptr;
    cudaError_t e;
    if ((e = bigfft(a, t, M, B, W, IN_REAL, OUT_POWER, io, s, &r)) != cudaSuccess) return e;
    return bigfft(r, (r == a) ? t : a, M, B, W, IN_PLAIN, OUT_REAL, io, s, &r2);
}

}  // namespace mlxa
