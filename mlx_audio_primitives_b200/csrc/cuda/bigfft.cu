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
// out[(j - k) R + k + q Ns], q < R.
template <int R>
__global__ void __launch_bounds__(256) bigfft_pass_kernel(const float2* __restrict__ in, float2* __restrict__ out, long long M,
                                                          long long Ns, const float2* __restrict__ W) {
    const long long j = blockIdx.x * 256LL + threadIdx.x;
    const long long per = M / R;
    if (j >= per) return;
    in += blockIdx.y * M;
    out += blockIdx.y * M;
    const long long k = j & (Ns - 1);
    const long long tstep = k * (M / (Ns * R));
    float2 v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = in[j + r * per];
    if (Ns > 1) {
#pragma unroll
        for (int r = 1; r < R; ++r) v[r] = cmul(v[r], __ldg(W + tstep * r));
    }
    DftInplace<R, 1, 0>::run(v);
    float2* o = out + (j - k) * R + k;
    static_for<R>([&](auto i_) {
        constexpr int i = decltype(i_)::value;
        o[dft_perm(R, i) * Ns] = v[i];
    });
}

cudaError_t bigfft(float2* a, float2* tmp, long long M, long long B, const float2* W, cudaStream_t s, float2** result) {
    // ping-pong a -> tmp -> a ...; *result is the buffer holding the transform
    float2* src = a;
    float2* dst = tmp;
    long long Ns = 1;
    while (Ns < M) {
        const long long left = M / Ns;
        const int R = left >= 16 ? 16 : (int)left;
        dim3 grid((unsigned)((M / R + 255) / 256), (unsigned)B);
        switch (R) {
            case 16: bigfft_pass_kernel<16><<<grid, 256, 0, s>>>(src, dst, M, Ns, W); break;
            case 8: bigfft_pass_kernel<8><<<grid, 256, 0, s>>>(src, dst, M, Ns, W); break;
            case 4: bigfft_pass_kernel<4><<<grid, 256, 0, s>>>(src, dst, M, Ns, W); break;
            default: bigfft_pass_kernel<2><<<grid, 256, 0, s>>>(src, dst, M, Ns, W); break;
        }
        Ns *= R;
        float2* t = src; src = dst; dst = t;
    }
    *result = src;
    return cudaGetLastError();
}

// a[j] = x[j] conj(w_n[j]) (j < n), 0 up to M
__global__ void rs_chirp_in_kernel(const float* __restrict__ x, long long n, long long ldx, const float2* __restrict__ w,
                                   long long M, float2* __restrict__ a) {
    const long long j = blockIdx.x * 256LL + threadIdx.x;
    if (j >= M) return;
    const long long b = blockIdx.y;
    float2 v = make_float2(0.f, 0.f);
    if (j < n) {
        const float xv = x[b * ldx + j];
        const float2 c = __ldg(w + j);
        v = make_float2(xv * c.x, -xv * c.y);
    }
    a[b * M + j] = v;
}
// a <- conj(a * bhat): the conjugate turns the following forward transform into the inverse one
__global__ void rs_pointwise_kernel(float2* __restrict__ a, const float2* __restrict__ bhat, long long M) {
    const long long j = blockIdx.x * 256LL + threadIdx.x;
    if (j >= M) return;
    float2* p = a + blockIdx.y * M + j;
    const float2 v = cmul(*p, __ldg(bhat + j));
    *p = make_float2(v.x, -v.y);
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
// c2 holds conj(M2 * convolution); y[j] = Re(w_num[j] conv[j]) / n (irfft's 1 / num times scipy's num / n), times scale
__global__ void rs_out_kernel(const float2* __restrict__ c2, long long M2, long long num, const float2* __restrict__ wm, float gain,
                              float* __restrict__ out, long long ldo) {
    const long long j = blockIdx.x * 256LL + threadIdx.x;
    if (j >= num) return;
    const long long b = blockIdx.y;
    const float2 cv = c2[b * M2 + j];
    const float2 w = __ldg(wm + j);
    // Re(w * conj(cv)) = w.x cv.x + w.y cv.y
    out[b * ldo + j] = (w.x * cv.x + w.y * cv.y) * gain;
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

cudaError_t cache_build(int kind, long long n, cudaStream_t s, float2** out) {
    cudaError_t e;
    if (kind == 0) {  // roots of unity of order n
        if ((e = cudaMalloc(out, size_t(n) * 8)) != cudaSuccess) return e;
        root_table_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(*out, n);
        return cudaGetLastError();
    }
    if (kind == 1) {  // chirp of length n
        if ((e = cudaMalloc(out, size_t(n) * 8)) != cudaSuccess) return e;
        chirp_table_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(*out, n);
        return cudaGetLastError();
    }
    // kind 2 / 3: FFT_M of the wrapped chirp / of its conjugate
    const long long M = pow2_at_least(2 * n - 1);
    float2 *w = nullptr, *W = nullptr, *tmp = nullptr, *res = nullptr;
    if ((e = cache_get(1, n, s, &w)) != cudaSuccess) return e;
    if ((e = cache_get(0, M, s, &W)) != cudaSuccess) return e;
    if ((e = cudaMalloc(out, size_t(M) * 8)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&tmp, size_t(M) * 8)) != cudaSuccess) return e;
    chirp_wrap_kernel<<<(unsigned)((M + 255) / 256), 256, 0, s>>>(w, n, M, kind == 3, *out);
    if ((e = bigfft(*out, tmp, M, 1, W, s, &res)) != cudaSuccess) return e;
    if (res != *out) e = cudaMemcpyAsync(*out, res, size_t(M) * 8, cudaMemcpyDeviceToDevice, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(tmp);
    return e;
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
    {
        std::lock_guard<std::mutex> lk(g_mu);
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
    rs_chirp_in_kernel<<<dim3((unsigned)((M1 + 255) / 256), by), 256, 0, s>>>(x, n, ldx, wn, M1, a);
    if ((e = bigfft(a, t, M1, B, W1, s, &r)) != cudaSuccess) return e;
    float2* o = (r == a) ? t : a;
    rs_pointwise_kernel<<<dim3((unsigned)((M1 + 255) / 256), by), 256, 0, s>>>(r, bh1, M1);
    float2* r2 = nullptr;
    if ((e = bigfft(r, o, M1, B, W1, s, &r2)) != cudaSuccess) return e;
    float2* o2 = (r2 == a) ? t : a;
    rs_mid_kernel<<<dim3((unsigned)((M2 + 255) / 256), by), 256, 0, s>>>(r2, M1, n, wn, num, wm, M2, o2);
    float2* r3 = nullptr;
    if ((e = bigfft(o2, r2, M2, B, W2, s, &r3)) != cudaSuccess) return e;
    float2* o3 = (r3 == a) ? t : a;
    rs_pointwise_kernel<<<dim3((unsigned)((M2 + 255) / 256), by), 256, 0, s>>>(r3, bh2, M2);
    float2* r4 = nullptr;
    if ((e = bigfft(r3, o3, M2, B, W2, s, &r4)) != cudaSuccess) return e;
    rs_out_kernel<<<dim3((unsigned)((num + 255) / 256), by), 256, 0, s>>>(r4, M2, num, wm, gain / (float(M2) * float(n)), out, ldo);
    return cudaGetLastError();
}

}  // namespace mlxa
