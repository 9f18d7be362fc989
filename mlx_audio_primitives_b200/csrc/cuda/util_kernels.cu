// Small kernels around the fused transforms: the reference-granularity pad / frame /
// overlap-add entry points, the dB family with its global-max reduction, the MFCC tail,
// elementwise complex helpers and the O(n^2) DFT used for n_fft values without a compiled plan.
#include <algorithm>
#include "fwd_epilogue.cuh"
#include <cstdlib>
#include "pcg64.cuh"
#include "util_kernels.cuh"

namespace mlxa {

constexpr int kThreads = 256;
static inline unsigned grid_for(long long n, int per_block, unsigned cap = 148u * 32u) {
    long long g = (n + per_block - 1) / per_block;
    if (g < 1) g = 1;
    return (unsigned)(g > cap ? cap : g);
}

// ---- pad / frame / overlap-add (reference granularity) -----------------------------------
__global__ void pad_kernel(const float* __restrict__ x, long long B, int L, int pad, int mode,
                           float* __restrict__ out) {
    const long long W = (long long)L + 2 * pad;
    const long long n = B * W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / W;
        const int q = int(i - b * W);
        out[i] = load_padded(x + b * L, L, q - pad, mode);
    }
}

__global__ void frame_kernel(const float* __restrict__ x, long long B, long long L, int fl, int hop,
                             long long T, float* __restrict__ out) {
    const long long n = B * T * fl;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int s = int(i % fl);
        const long long bt = i / fl;
        const long long t = bt % T, b = bt / T;
        out[i] = __ldg(x + b * L + t * hop + s);  // frames[b,t,s] = x[b, t*hop + s]
    }
}

// frame range covering output sample i: overlap_add.metal:36-37
__device__ __forceinline__ void ola_range(long long i, int n_fft, int hop, long long T, long long& f0, long long& f1) {
    f0 = (i < n_fft) ? 0 : (i - n_fft) / hop + 1;
    f1 = i / hop;
    if (f1 > T - 1) f1 = T - 1;
}

__global__ void wss_kernel(const float* __restrict__ w, int n_fft, int hop, long long T, long long out_len,
                           float* __restrict__ wss) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < out_len; i += (long long)gridDim.x * blockDim.x) {
        long long f0, f1;
        ola_range(i, n_fft, hop, T, f0, f1);
        float s = 0.f;
        for (long long f = f0; f <= f1; ++f) {
            const float v = __ldg(w + (i - f * hop));
            s += v * v;
        }
        wss[i] = s;
    }
}

__global__ void ola_kernel(const float* __restrict__ frames, const float* __restrict__ w, long long B, long long T,
                           int n_fft, int hop, long long ola_len, long long trim, long long out_len, long long ldy,
                           float* __restrict__ y) {
    const long long n = B * out_len;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        const long long b = idx / out_len, j = idx - b * out_len;
        const long long i = j + trim;
        float val = 0.f;
        if (i < ola_len) {
            long long f0, f1;
            ola_range(i, n_fft, hop, T, f0, f1);
            float s = 0.f, ws = 0.f;
            const float* fb = frames + b * T * n_fft;
            for (long long f = f0; f <= f1; ++f) {
                const int k = int(i - f * hop);
                const float wv = __ldg(w + k);
                s += wv * __ldg(fb + f * n_fft + k);
                ws += wv * wv;
            }
            val = s / fmaxf(ws, 1e-8f);
        }
        y[b * ldy + j] = val;
    }
}

// ---- elementwise ------------------------------------------------------------------------
__global__ void magnitude_kernel(const float2* __restrict__ z, long long n, float* __restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float2 v = z[i];
        out[i] = hypotf(v.x, v.y);
    }
}
__global__ void phase_kernel(const float2* __restrict__ z, long long n, float* __restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float2 v = z[i];
        out[i] = atan2f(v.y, v.x);
    }
}
__global__ void polar_kernel(const float* __restrict__ mag, const float* __restrict__ ang, long long n,
                             float2* __restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float s, c;
        sincosf(ang[i], &s, &c);
        const float m = mag[i];
        out[i] = make_float2(m * c, m * s);
    }
}
// y = u + m * (u - u_prev) over (B, n) rows of stride ld (u_prev == nullptr: y = u): the signal-domain form of
// Griffin-Lim's momentum step for transforms that do not fuse it into their store loop (the O(n^2) fallback)
__global__ void momentum_kernel(const float* __restrict__ u, const float* __restrict__ u_prev, float m, long long B,
                                long long n, long long ld, float* __restrict__ y) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < B * n; i += (long long)gridDim.x * blockDim.x) {
        const long long o = (i / n) * ld + i % n;
        const float v = u[o];
        y[o] = u_prev ? fmaf(m, v - u_prev[o], v) : v;
    }
}
__global__ void fill_kernel(float* x, long long n, float v) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) x[i] = v;
}

// batched (R, C) -> (C, R) through a padded smem tile (coalesced both ways)
template <class T>
__global__ void transpose_kernel(const T* __restrict__ in, long long R, long long C, T* __restrict__ out) {
    __shared__ T tile[32][33];
    const long long b = blockIdx.z;
    const T* ib = in + b * R * C;
    T* ob = out + b * R * C;
    const long long r0 = (long long)blockIdx.y * 32, c0 = (long long)blockIdx.x * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const long long r = r0 + j, c = c0 + threadIdx.x;
        if (r < R && c < C) tile[j][threadIdx.x] = ib[r * C + c];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const long long c = c0 + j, r = r0 + threadIdx.x;
        if (r < R && c < C) ob[c * R + r] = tile[threadIdx.x][j];
    }
}

// ---- global max ---------------------------------------------------------------------------
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
    if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned*>(addr), __float_as_uint(v));
}
__global__ void max_kernel(const float* __restrict__ x, long long n, float* gmax) {
    float m = -INFINITY;
    const long long n4 = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) ? n / 4 : 0;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = x4[i];
        m = fmaxf(fmaxf(m, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
    }
    for (long long i = n4 * 4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        m = fmaxf(m, x[i]);
    __shared__ float s[kThreads / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < kThreads / 32; ++i) m = fmaxf(m, s[i]);
        if (m > -INFINITY) atomic_max_float(gmax, m);
    }
}

// one thread: this rank's finished peak to every peer (the host-buffer path publishes once, after its last chunk)
__global__ void peak_publish_kernel(PeakExchange xchg, float* gmax) { peak_publish(xchg, gmax, 1u); }

// ---- dB family (convert.py:14-60) ---------------------------------------------------------
__global__ void to_db_kernel(const float* __restrict__ x, long long n, float coef, float amin, float ref_host,
                             const float* __restrict__ ref_dev, int use_top, float top_db,
                             const float* __restrict__ gmax, float* __restrict__ out, float* reset_next,
                             PeakExchange xchg) {
    if (reset_next != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *reset_next = 0.f;  // peak slot of the NEXT call
    __shared__ float s_peak;
    // with an exchange, gmax (and a ref_dev that aliases it, i.e. ref = max) mean the maximum over all ranks
    const float peak = (use_top || xchg.peer_slots != nullptr) ? resolve_peak(gmax, xchg, &s_peak) : 0.f;
    const float ref = ref_dev ? ((xchg.peer_slots != nullptr && ref_dev == gmax) ? peak : __ldg(ref_dev)) : ref_host;
    const float refc = fmaxf(ref, amin);
    float floor_db = -INFINITY;
    if (use_top) floor_db = to_db_one(peak, coef, amin, refc) - top_db;
    const bool al = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    const long long n4 = al ? n / 4 : 0;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float4* o4 = reinterpret_cast<float4*>(out);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = x4[i];
        v.x = fmaxf(to_db_one(v.x, coef, amin, refc), floor_db);
        v.y = fmaxf(to_db_one(v.y, coef, amin, refc), floor_db);
        v.z = fmaxf(to_db_one(v.z, coef, amin, refc), floor_db);
        v.w = fmaxf(to_db_one(v.w, coef, amin, refc), floor_db);
        o4[i] = v;
    }
    for (long long i = n4 * 4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = fmaxf(to_db_one(x[i], coef, amin, refc), floor_db);
}
// Second half of a fused log-mel: x already holds coef*log10(max(S, amin)/refc) (written by the mel
// kernel's epilogue); raise everything below max(x) - top_db to that floor.  max(x) follows from the peak
// of S by monotonicity.  Only vectors that change are written back, so the pass is read-mostly and, right
// behind its producer, served from L2.
__global__ void db_floor_kernel(float* __restrict__ x, long long n, float coef, float amin, float ref, float top_db,
                                const float* __restrict__ gmax, float* reset_next) {
    if (reset_next != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *reset_next = 0.f;
    const float floor_db = to_db_one(__ldg(gmax), coef, amin, fmaxf(ref, amin)) - top_db;
    const long long n4 = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) ? n / 4 : 0;
    float4* x4 = reinterpret_cast<float4*>(x);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = x4[i];
        if (fminf(fminf(v.x, v.y), fminf(v.z, v.w)) < floor_db) {
            v.x = fmaxf(v.x, floor_db); v.y = fmaxf(v.y, floor_db); v.z = fmaxf(v.z, floor_db); v.w = fmaxf(v.w, floor_db);
            x4[i] = v;
        }
    }
    for (long long i = n4 * 4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        if (x[i] < floor_db) x[i] = floor_db;
}
// The same floor with the producer's per-block minima (mlxa_melspec_f32 block_min): a 64-frame block of a
// clip whose smallest value is already at or above the floor is not touched at all, so on material without
// 80 dB of dynamic range the pass reads B*ceil(T/64) floats.  Consumed slots are re-armed to +inf.
// host_mirror (optional): a device-accessible alias of a pinned HOST copy of x that already holds the values
// from before the floor (the host path's speculative copy-back); raised values are patched into it directly.
// A CTA looks at kSlotsPerCta (clip, block) slots at once -- one thread each, one barrier -- and then
// rewrites the flagged ones with all of its warps (a CTA per slot costs ~16 us in launch + barrier latency
// for the 3008 slots of a 64 x 30 s batch even when nothing is flagged; tools/probes/floor_probe.cu).
constexpr int kSlotsPerCta = 8, kFloorThreads = 512;
__global__ void __launch_bounds__(kFloorThreads)
db_floor_blocks_kernel(float* __restrict__ x, long long n_slots, int n_bands, long long T, float coef, float amin, float ref,
                       float top_db, const float* __restrict__ gmax, float* __restrict__ block_min, float* reset_next,
                       int* n_raised, PeakExchange xchg, float* __restrict__ host_mirror) {
    if (reset_next != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *reset_next = 0.f;
    __shared__ float s_peak;
    const float refc = fmaxf(ref, amin);
    const float floor_db = to_db_one(resolve_peak(gmax, xchg, &s_peak), coef, amin, refc) - top_db;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = kFloorThreads / 32, U = 5;
    const long long nblk = (T + kMinBlockFrames - 1) / kMinBlockFrames;
    __shared__ unsigned s_flags;
    const long long slot0 = (long long)blockIdx.x * kSlotsPerCta;
    if (threadIdx.x < 32) {
        bool f = false;
        if (lane < kSlotsPerCta && slot0 + lane < n_slots) {
            const float m = block_min[slot0 + lane];
            block_min[slot0 + lane] = INFINITY;
            f = to_db_one(m, coef, amin, refc) < floor_db;
        }
        const unsigned fl = __ballot_sync(0xffffffffu, f);
        if (lane == 0) s_flags = fl;
    }
    __syncthreads();
    unsigned fl = s_flags;
    while (fl) {
        const int i = __ffs(fl) - 1;
        fl &= fl - 1;
        const long long slot = slot0 + i, b = slot / nblk, blk = slot - b * nblk;
        if (threadIdx.x == 0 && n_raised != nullptr) n_raised[1 + atomicAdd(n_raised, 1)] = (int)slot;
        // a warp per band row: 64 frames = two coalesced 128-byte accesses; U rows in flight per warp
        const long long t0 = blk * kMinBlockFrames;
        float* xb = x + b * n_bands * T + t0 + lane;
        const bool ok0 = t0 + lane < T, ok1 = t0 + lane + 32 < T;
        for (int m0 = warp; m0 < n_bands; m0 += U * NW) {
            float v0[U], v1[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int m = m0 + u * NW;
                v0[u] = (m < n_bands && ok0) ? xb[(long long)m * T] : INFINITY;
                v1[u] = (m < n_bands && ok1) ? xb[(long long)m * T + 32] : INFINITY;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int m = m0 + u * NW;
                if (v0[u] < floor_db) {
                    xb[(long long)m * T] = floor_db;
                    if (host_mirror != nullptr) host_mirror[(xb - x) + (long long)m * T] = floor_db;
                }
                if (v1[u] < floor_db) {
                    xb[(long long)m * T + 32] = floor_db;
                    if (host_mirror != nullptr) host_mirror[(xb - x) + (long long)m * T + 32] = floor_db;
                }
            }
        }
    }
}
__global__ void from_db_kernel(const float* __restrict__ x, long long n, float ref, float div, float* __restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = ref * powf(10.0f, x[i] / div);
}

// ---- DCT (rows, n_in) @ D^T and the fused MFCC tail -----------------------------------------
__global__ void dct_kernel(const float* __restrict__ x, long long rows, int n_in, const float* __restrict__ D,
                           int n_out, float* __restrict__ out) {
    extern __shared__ float s_x[];  // [8][n_in]
    const long long r0 = (long long)blockIdx.x * 8;
    const int nr = int(min(8LL, rows - r0));
    for (int i = threadIdx.x; i < nr * n_in; i += blockDim.x) s_x[i] = x[r0 * n_in + i];
    __syncthreads();
    for (int o = threadIdx.x; o < nr * n_out; o += blockDim.x) {
        const int r = o / n_out, k = o - r * n_out;
        const float* d = D + (long long)k * n_in;
        const float* xr = s_x + r * n_in;
        float acc = 0.f;
        for (int m = 0; m < n_in; ++m) acc = fmaf(xr[m], __ldg(d + m), acc);
        out[(r0 + r) * n_out + k] = acc;
    }
}

// mel (B, n_mels, T) -> dB -> DCT over the mel axis -> lifter -> (B, n_mfcc, T); lanes along T
__global__ void mfcc_tail_kernel(const float* __restrict__ mel, int n_mels, long long T, const float* __restrict__ D,
                                 int n_mfcc, const float* __restrict__ lifter, int apply_db, float amin, float ref,
                                 int use_top, float top_db, const float* __restrict__ gmax, float* __restrict__ out) {
    extern __shared__ float s_m[];          // [n_mels][33] dB tile, then D^T [n_mels][n_mfcc]
    float* s_d = s_m + n_mels * 33;
    const long long b = blockIdx.y, t0 = (long long)blockIdx.x * 32;
    const int nt = int(min(32LL, T - t0));
    const float refc = fmaxf(ref, amin);
    float floor_db = -INFINITY;
    if (apply_db && use_top) floor_db = to_db_one(__ldg(gmax), 10.0f, amin, refc) - top_db;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int m = warp; m < n_mels; m += nw) {
        float v = 0.f;
        if (lane < nt) {
            v = __ldg(mel + (b * n_mels + m) * T + t0 + lane);
            if (apply_db) v = fmaxf(to_db_one(v, 10.0f, amin, refc), floor_db);
        }
        s_m[m * 33 + lane] = v;
    }
    for (int i = threadIdx.x; i < n_mels * n_mfcc; i += blockDim.x) {
        const int k = i / n_mels, m = i - k * n_mels;
        s_d[m * n_mfcc + k] = __ldg(D + i);
    }
    __syncthreads();
    for (int k = warp; k < n_mfcc; k += nw) {
        float a0 = 0.f, a1 = 0.f;
        int m = 0;
        for (; m + 1 < n_mels; m += 2) {
            a0 = fmaf(s_d[m * n_mfcc + k], s_m[m * 33 + lane], a0);
            a1 = fmaf(s_d[(m + 1) * n_mfcc + k], s_m[(m + 1) * 33 + lane], a1);
        }
        if (m < n_mels) a0 = fmaf(s_d[m * n_mfcc + k], s_m[m * 33 + lane], a0);
        float v = a0 + a1;
        if (lifter) v *= __ldg(lifter + k);
        if (lane < nt) out[(b * n_mfcc + k) * T + t0 + lane] = v;
    }
}

// The same tail as a register-tiled product (the default): a CTA takes 128 frames of one clip; every thread first
// converts its own mel column to dB into a shared [n_mels][128] tile (coalesced reads along T, no transposition);
// then a warp owns 32 frames x all coefficients, lane = (coefficient group of KT, four consecutive frames): per mel
// band KT/2 64-bit reads of the transposed, zero-padded DCT row, one 128-bit read of four dB values and 2*KT packed
// FMAs -- 0.65 instructions and 0.005 shared-memory wavefronts per multiply-add; the kernel above spends 3 and 0.06,
// and reloads the DCT matrix for every 32 frames (c4: 0.92 ms of 6.4).
constexpr int kTailFrames = 128;
template <int KT>
__global__ void __launch_bounds__(kTailFrames)
mfcc_tail_tiled_kernel(const float* __restrict__ mel, int n_mels, long long T, const float* __restrict__ D, int n_mfcc,
                       const float* __restrict__ lifter, int apply_db, float amin, float ref, int use_top, float top_db,
                       const float* __restrict__ gmax, float* __restrict__ out) {
    static_assert(KT % 2 == 0, "coefficient pairs");
    constexpr int ROW = 4 * KT;                   // padded coefficients per DCT row
    extern __shared__ __align__(16) float s_tail[];
    float* s_x = s_tail;                          // [n_mels][128] dB values
    float* s_d = s_tail + n_mels * kTailFrames;   // [n_mels][ROW]: D^T, zero beyond n_mfcc
    const long long b = blockIdx.y, t0 = (long long)blockIdx.x * kTailFrames, t = t0 + threadIdx.x;
    const float refc = fmaxf(ref, amin);
    float floor_db = -INFINITY;
    if (apply_db && use_top) floor_db = to_db_one(__ldg(gmax), 10.0f, amin, refc) - top_db;
    // the thread's column: n_mels 4-byte async copies in four groups, all in flight at once (a register-staged loop
    // keeps 8 loads per thread in flight and leaves the kernel waiting on HBM latency); the DCT rows are fetched
    // while they travel, and each quarter is converted and consumed as it lands
    const float* col = mel + b * n_mels * T + (t < T ? t : 0);
    const int mq = (n_mels + 3) / 4;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        for (int m = q * mq; m < min(n_mels, (q + 1) * mq); ++m) cp_async4(s_x + m * kTailFrames + threadIdx.x, col + (long long)m * T);
        cp_async_commit();
    }
    for (int i = threadIdx.x; i < n_mels * ROW; i += kTailFrames) {
        const int m = i / ROW, k = i - m * ROW;
        s_d[i] = (k < n_mfcc) ? __ldg(D + (long long)k * n_mels + m) : 0.f;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int kg = lane >> 3, tj = lane & 7;      // coefficients kg*KT .. +KT, frames 32*warp + 4*tj .. +4
    float2 acc[4][KT / 2];
#pragma unroll
    for (int f = 0; f < 4; ++f)
#pragma unroll
        for (int i = 0; i < KT / 2; ++i) acc[f][i] = make_float2(0.f, 0.f);
    const float4* xr = reinterpret_cast<const float4*>(s_x + 32 * warp + 4 * tj);
    const float2* dr = reinterpret_cast<const float2*>(s_d + kg * KT);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (q == 0) cp_async_wait_group<3>();
        else if (q == 1) cp_async_wait_group<2>();
        else if (q == 2) cp_async_wait_group<1>();
        else cp_async_wait_group<0>();
        const int m_lo = q * mq, m_hi = min(n_mels, (q + 1) * mq);
        if (apply_db) {
#pragma unroll 8
            for (int m = m_lo; m < m_hi; ++m) {
                float* px = s_x + m * kTailFrames + threadIdx.x;
                *px = fmaxf(to_db_one(*px, 10.0f, amin, refc), floor_db);
            }
        }
        __syncwarp();  // a lane reads four columns of its own warp
#pragma unroll 2
    for (int m = m_lo; m < m_hi; ++m) {
        const float4 x = xr[m * (kTailFrames / 4)];
        float2 d[KT / 2];
#pragma unroll
        for (int i = 0; i < KT / 2; ++i) d[i] = dr[m * (ROW / 2) + i];
#pragma unroll
        for (int i = 0; i < KT / 2; ++i) {
            acc[0][i] = pfma(d[i].x, d[i].y, x.x, x.x, acc[0][i]);
            acc[1][i] = pfma(d[i].x, d[i].y, x.y, x.y, acc[1][i]);
            acc[2][i] = pfma(d[i].x, d[i].y, x.z, x.z, acc[2][i]);
            acc[3][i] = pfma(d[i].x, d[i].y, x.w, x.w, acc[3][i]);
        }
    }
    }
    const long long tf = t0 + 32 * warp + 4 * tj;
#pragma unroll
    for (int i = 0; i < KT / 2; ++i) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k = kg * KT + 2 * i + h;
            if (k < n_mfcc) {
                const float l = lifter ? __ldg(lifter + k) : 1.f;
                float* o = out + (b * n_mfcc + k) * T + tf;
#pragma unroll
                for (int f = 0; f < 4; ++f)
                    if (tf + f < T) o[f] = (h ? acc[f][i].y : acc[f][i].x) * l;
            }
        }
    }
}
template <int KT>
cudaError_t launch_mfcc_tail_tiled(const float* mel, long long B, int n_mels, long long T, const float* D, int n_mfcc,
                                   const float* lifter, int apply_db, float amin, float ref, int use_top, float top_db,
                                   const float* gmax, float* out, cudaStream_t s) {
    const size_t smem = size_t(n_mels) * (kTailFrames + 4 * KT) * 4;
    cudaError_t e = cudaFuncSetAttribute(mfcc_tail_tiled_kernel<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid((unsigned)((T + kTailFrames - 1) / kTailFrames), (unsigned)B);
    mfcc_tail_tiled_kernel<KT><<<grid, kTailFrames, smem, s>>>(mel, n_mels, T, D, n_mfcc, lifter, apply_db, amin, ref, use_top,
                                                            top_db, gmax, out);
    return cudaGetLastError();
}

// ---- O(n^2) DFT fallback: any n_fft, same epilogues as the planned kernels -----------------
template <int EP>
__global__ void __launch_bounds__(256) fwd_naive_kernel(const FwdParams p, int nwarps) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int TT = p.tile_frames, n_fft = p.n_fft;
    const int b = blockIdx.y, t0 = blockIdx.x * TT, nt = min(TT, p.T - t0);
    const int xs = (n_fft + 3) & ~3, ps = (p.F + 3) & ~3;
    float* s_x = reinterpret_cast<float*>(smem_raw);  // [nwarps][xs] windowed frames
    float* s_p = s_x + nwarps * xs;                    // [nwarps][ps] |X|^p of the warp's frame (EP_MEL)
    float* s_out = s_p + (EP == EP_MEL ? nwarps * ps : 0);
    const int out_floats = (EP == EP_MEL) ? ((p.n_bands * (TT + 1) + 3) & ~3) : 0;
    float* s_mel = s_out + out_floats;
    __shared__ float s_red[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* yb = p.y + (long long)b * p.ldy;
    float* xw = s_x + warp * xs;
    MelSmem ms{};
    if constexpr (EP == EP_MEL) {
        ms = mel_smem_carve<32>(s_mel, p.n_bands, p.n_w4);
        const int words = (int)packed_bank_words(p.n_bands, p.n_w4, 32);
        for (int i = threadIdx.x; i < words; i += 256) s_mel[i] = __ldg(p.bank + i);
        __syncthreads();
    }
    if (warp < nwarps) {
        for (int f = warp; f < nt; f += nwarps) {
            const int t = t0 + f;
            const bool valid = t < p.T_valid;
            for (int n = lane; n < n_fft; n += 32)
                xw[n] = valid ? load_padded(yb, p.L, t * p.hop - p.pad + n, p.pad_mode) * __ldg(p.window + n) : 0.f;
            __syncwarp();
            for (int k = lane; k < p.F; k += 32) {
                float re = 0.f, im = 0.f;
                int idx = 0;
                for (int n = 0; n < n_fft; ++n) {
                    const float2 w = __ldg(p.tw_plan + idx);  // exp(-2*pi*i*idx/n_fft)
                    re = fmaf(xw[n], w.x, re);
                    im = fmaf(xw[n], w.y, im);
                    idx += k;
                    if (idx >= n_fft) idx -= n_fft;
                }
                const float2 X = make_float2(re, im);
                if constexpr (EP == EP_MEL) {
                    float pw;
                    if (p.power_mode == POW_SQUARE) pw = spectral_power<POW_SQUARE>(X, p.power);
                    else if (p.power_mode == POW_ABS) pw = spectral_power<POW_ABS>(X, p.power);
                    else pw = spectral_power<POW_GENERAL>(X, p.power);
                    s_p[warp * ps + k] = pw;
                } else {
                    epilogue_bin_global<EP>(p, ((long long)b * p.T + t) * p.F + k, X);
                }
            }
            __syncwarp();
            if constexpr (EP == EP_MEL) {
                mel_project_group<32, 1>(ms, p.n_bands, lane, s_p + warp * ps, s_out, TT + 1, f);
                __syncwarp();
            }
        }
    }
    if constexpr (EP == EP_MEL) {
        __syncthreads();
        const float vmax = mel_store_tile<256>(p, b, t0, nt, s_out, TT, 3, 0.f);  // TT = 8
        if (p.gmax != nullptr) block_max_to_global<256>(vmax, p.gmax, s_red, p.xchg);
    }
}

cudaError_t launch_fwd_naive(int ep, FwdParams& p, cudaStream_t s) {
    const int TT = 8;
    const size_t xs = size_t((p.n_fft + 3) & ~3) * 4, ps = (ep == EP_MEL) ? size_t((p.F + 3) & ~3) * 4 : 0;
    const size_t fixed = (ep == EP_MEL)
        ? size_t((p.n_bands * (TT + 1) + 3) & ~3) * 4 + size_t(packed_bank_words(p.n_bands, p.n_w4, 32)) * 4 : 0;
    int nwarps = 8;
    while (nwarps > 1 && size_t(nwarps) * (xs + ps) + fixed > 200 * 1024) nwarps >>= 1;
    const size_t smem = size_t(nwarps) * (xs + ps) + fixed;
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    p.tile_frames = TT;
    dim3 grid((p.T + TT - 1) / TT, p.B);
    cudaError_t e;
#define MLXA_LAUNCH(EPV)                                                                               \
    e = cudaFuncSetAttribute(fwd_naive_kernel<EPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return e;                                                                    \
    fwd_naive_kernel<EPV><<<grid, 256, smem, s>>>(p, nwarps);
    if (ep == EP_STFT) { MLXA_LAUNCH(EP_STFT) }
    else if (ep == EP_MEL) { MLXA_LAUNCH(EP_MEL) }
    else { MLXA_LAUNCH(EP_GL) }
#undef MLXA_LAUNCH
    return cudaGetLastError();
}

// frames[row, n] = irfft(spec[row, :F_in], n=n_fft)[n]  (1/n normalised; imag of DC/Nyquist ignored)
__global__ void irdft_naive_kernel(const float2* __restrict__ spec, long long rows, int F_in, int n_fft,
                                   const float2* __restrict__ tw, float* __restrict__ frames) {
    extern __shared__ float2 s_X[];  // [F] bins of this row
    const long long row = blockIdx.x;
    const int F = n_fft / 2 + 1;
    for (int k = threadIdx.x; k < F; k += blockDim.x) {
        float2 v = (k < F_in) ? spec[row * F_in + k] : make_float2(0.f, 0.f);
        if (k == 0 || 2 * k == n_fft) v.y = 0.f;
        s_X[k] = v;
    }
    __syncthreads();
    const float inv = 1.0f / float(n_fft);
    for (int n = threadIdx.x; n < n_fft; n += blockDim.x) {
        float acc = s_X[0].x;
        int idx = 0;
        for (int k = 1; k < F; ++k) {
            idx += n;
            if (idx >= n_fft) idx -= n_fft;
            const float2 w = __ldg(tw + idx);  // exp(-2*pi*i*k*n/n_fft); need Re(X * conj(w))
            const float term = fmaf(s_X[k].x, w.x, s_X[k].y * w.y);
            acc += (2 * k == n_fft) ? term : 2.0f * term;
        }
        frames[row * n_fft + n] = acc * inv;
    }
}

cudaError_t launch_irdft_naive(const float2* spec, long long rows, int F_in, int n_fft, const float2* tw_full,
                               float* frames, cudaStream_t s) {
    const size_t smem = size_t(n_fft / 2 + 1) * 8;
    irdft_naive_kernel<<<(unsigned)rows, 256, smem, s>>>(spec, rows, F_in, n_fft, tw_full, frames);
    return cudaGetLastError();
}

// ---- NumPy-compatible uniform stream (pcg64.cuh): each thread jumps to its chunk ---------------
constexpr int kPcgChunk = 32;
constexpr int kPcgJumpBits = 48;  // stream offsets below 2^48 draws
__global__ void pcg64_uniform_kernel(unsigned long long s_hi, unsigned long long s_lo, unsigned long long i_hi,
                                     unsigned long long i_lo, double low, double range, long long n, float* __restrict__ out) {
    const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long i0 = c * kPcgChunk;
    if (i0 >= n) return;
    const u128 inc = ((u128)i_hi << 64) | i_lo;
    u128 st = pcg_advance(((u128)s_hi << 64) | s_lo, inc, (unsigned long long)i0);
    const int m = (int)min((long long)kPcgChunk, n - i0);
    for (int j = 0; j < m; ++j) out[i0 + j] = pcg_uniform_f32(st, inc, low, range);
}
cudaError_t run_pcg64_uniform(unsigned long long s_hi, unsigned long long s_lo, unsigned long long i_hi,
                              unsigned long long i_lo, double low, double range, long long n, float* out, cudaStream_t s) {
    const long long chunks = (n + kPcgChunk - 1) / kPcgChunk;
    pcg64_uniform_kernel<<<(unsigned)((chunks + 127) / 128), 128, 0, s>>>(s_hi, s_lo, i_hi, i_lo, low, range, n, out);
    return cudaGetLastError();
}

// ---- Griffin-Lim's random start in one pass: S * exp(i * uniform(-pi, pi)) in the physical (B, T, F) layout -----
// The phases are NumPy's stream drawn in LOGICAL (B, F, T) order (reference griffinlim.py:112-123), so value
// n = (b*F + f)*T + t belongs at physical element (b, t, f).  A block owns 32 bins x kInitT frames of one clip: each
// thread jumps to the start of a run of kInitRun frames of one bin (one multiply-add per set bit of the offset, from the
// host-built table of 2^i-step multipliers), draws the run into a shared tile, and the block then walks the tile
// along the bins: coalesced reads of the magnitudes and writes of the complex spectrum.  The angles never exist in HBM.
constexpr int kInitT = 256, kInitRun = 32;
struct PcgJumpTable {
    unsigned long long m_hi[kPcgJumpBits], m_lo[kPcgJumpBits], p_hi[kPcgJumpBits], p_lo[kPcgJumpBits];
};
__global__ void __launch_bounds__(256) pcg64_polar_kernel(const PcgJumpTable tbl, unsigned long long s_hi, unsigned long long s_lo,
                                                          unsigned long long i_hi, unsigned long long i_lo, double low, double range,
                                                          const float* __restrict__ mag, long long F, long long T,
                                                          float2* __restrict__ out) {
    __shared__ float s_ang[kInitT][33];
    const int fi = threadIdx.x & 31, ch = threadIdx.x >> 5;
    const long long b = blockIdx.z, f = (long long)blockIdx.y * 32 + fi, t0 = (long long)blockIdx.x * kInitT;
    const long long ts = t0 + ch * kInitRun;
    if (f < F && ts < T) {
        const u128 inc = ((u128)i_hi << 64) | i_lo;
        u128 st = ((u128)s_hi << 64) | s_lo;
        unsigned long long delta = (unsigned long long)((b * F + f) * T + ts);
        for (int i = 0; delta != 0; ++i, delta >>= 1)
            if (delta & 1) st = st * (((u128)tbl.m_hi[i] << 64) | tbl.m_lo[i]) + (((u128)tbl.p_hi[i] << 64) | tbl.p_lo[i]);
        const int m = (int)min((long long)kInitRun, T - ts);
        for (int j = 0; j < m; ++j) s_ang[ch * kInitRun + j][fi] = pcg_uniform_f32(st, inc, low, range);
    }
    __syncthreads();
    if (f < F) {
        const int m = (int)max(0LL, min((long long)kInitRun, T - ts));
#pragma unroll 4
        for (int j = 0; j < m; ++j) {
            const long long o = (b * T + ts + j) * F + f;
            float sn, cs;
            sincosf(s_ang[ch * kInitRun + j][fi], &sn, &cs);
            const float mg = __ldg(mag + o);
            out[o] = make_float2(mg * cs, mg * sn);
        }
    }
}
cudaError_t run_pcg64_polar(unsigned long long s_hi, unsigned long long s_lo, unsigned long long i_hi, unsigned long long i_lo,
                            double low, double range, const float* mag, long long B, long long F, long long T, float2* out,
                            cudaStream_t s) {
    PcgJumpTable tbl;  // step 2^i of s <- s*mult + inc:  M_0 = mult, P_0 = inc;  M_{i+1} = M_i^2, P_{i+1} = (M_i + 1) P_i
    u128 m = pcg_mult(), pl = ((u128)i_hi << 64) | i_lo;
    for (int i = 0; i < kPcgJumpBits; ++i) {
        tbl.m_hi[i] = (unsigned long long)(m >> 64); tbl.m_lo[i] = (unsigned long long)m;
        tbl.p_hi[i] = (unsigned long long)(pl >> 64); tbl.p_lo[i] = (unsigned long long)pl;
        pl = (m + 1) * pl;
        m *= m;
    }
    dim3 grid((unsigned)((T + kInitT - 1) / kInitT), (unsigned)((F + 31) / 32), (unsigned)B);
    pcg64_polar_kernel<<<grid, 256, 0, s>>>(tbl, s_hi, s_lo, i_hi, i_lo, low, range, mag, F, T, out);
    return cudaGetLastError();
}

// ---- host launch wrappers -----------------------------------------------------------------
cudaError_t run_pad(const float* x, long long B, int L, int pad, int mode, float* out, cudaStream_t s) {
    pad_kernel<<<grid_for(B * ((long long)L + 2 * pad), kThreads), kThreads, 0, s>>>(x, B, L, pad, mode, out);
    return cudaGetLastError();
}
cudaError_t run_frame(const float* x, long long B, long long L, int fl, int hop, long long T, float* out, cudaStream_t s) {
    frame_kernel<<<grid_for(B * T * fl, kThreads), kThreads, 0, s>>>(x, B, L, fl, hop, T, out);
    return cudaGetLastError();
}
cudaError_t run_wss(const float* w, int n_fft, int hop, long long T, long long out_len, float* wss, cudaStream_t s) {
    wss_kernel<<<grid_for(out_len, kThreads), kThreads, 0, s>>>(w, n_fft, hop, T, out_len, wss);
    return cudaGetLastError();
}
cudaError_t run_ola(const float* frames, const float* w, long long B, long long T, int n_fft, int hop,
                    long long ola_len, long long trim, long long out_len, long long ldy, float* y, cudaStream_t s) {
    ola_kernel<<<grid_for(B * out_len, kThreads), kThreads, 0, s>>>(frames, w, B, T, n_fft, hop, ola_len, trim, out_len, ldy, y);
    return cudaGetLastError();
}
cudaError_t run_magnitude(const float2* z, long long n, float* out, cudaStream_t s) {
    magnitude_kernel<<<grid_for(n, kThreads), kThreads, 0, s>>>(z, n, out);
    return cudaGetLastError();
}
cudaError_t run_phase(const float2* z, long long n, float* out, cudaStream_t s) {
    phase_kernel<<<grid_for(n, kThreads), kThreads, 0, s>>>(z, n, out);
    return cudaGetLastError();
}
cudaError_t run_polar(const float* mag, const float* ang, long long n, float2* out, cudaStream_t s) {
    polar_kernel<<<grid_for(n, kThreads), kThreads, 0, s>>>(mag, ang, n, out);
    return cudaGetLastError();
}
cudaError_t run_momentum(const float* u, const float* u_prev, float m, long long B, long long n, long long ld, float* y,
                         cudaStream_t s) {
    if (u == y && u_prev == nullptr) return cudaSuccess;
    momentum_kernel<<<grid_for(B * n, kThreads), kThreads, 0, s>>>(u, u_prev, m, B, n, ld, y);
    return cudaGetLastError();
}
cudaError_t run_fill(float* x, long long n, float v, cudaStream_t s) {
    fill_kernel<<<grid_for(n, kThreads), kThreads, 0, s>>>(x, n, v);
    return cudaGetLastError();
}
cudaError_t run_transpose_f32(const float* in, long long B, long long R, long long C, float* out, cudaStream_t s) {
    dim3 grid((unsigned)((C + 31) / 32), (unsigned)((R + 31) / 32), (unsigned)B), blk(32, 8);
    transpose_kernel<float><<<grid, blk, 0, s>>>(in, R, C, out);
    return cudaGetLastError();
}
cudaError_t run_transpose_c64(const float2* in, long long B, long long R, long long C, float2* out, cudaStream_t s) {
    dim3 grid((unsigned)((C + 31) / 32), (unsigned)((R + 31) / 32), (unsigned)B), blk(32, 8);
    transpose_kernel<float2><<<grid, blk, 0, s>>>(in, R, C, out);
    return cudaGetLastError();
}
cudaError_t run_peak_publish(const PeakExchange& xchg, float* gmax, cudaStream_t s) {
    peak_publish_kernel<<<1, 1, 0, s>>>(xchg, gmax);
    return cudaGetLastError();
}
cudaError_t run_max(const float* x, long long n, float* gmax, cudaStream_t s) {
    max_kernel<<<grid_for(n, kThreads * 8, 148u * 8u), kThreads, 0, s>>>(x, n, gmax);
    return cudaGetLastError();
}
cudaError_t run_to_db(const float* x, long long n, float coef, float amin, float ref_host, const float* ref_dev,
                      int use_top, float top_db, const float* gmax, float* out, float* reset_next, const PeakExchange& xchg,
                      cudaStream_t s) {
    to_db_kernel<<<grid_for(n, kThreads * 8, 148u * 16u), kThreads, 0, s>>>(x, n, coef, amin, ref_host, ref_dev, use_top, top_db, gmax, out, reset_next, xchg);
    return cudaGetLastError();
}
cudaError_t run_db_floor(float* x, long long n, float coef, float amin, float ref, float top_db, const float* gmax,
                         float* reset_next, cudaStream_t s) {
    db_floor_kernel<<<grid_for(n, kThreads * 8, 148u * 16u), kThreads, 0, s>>>(x, n, coef, amin, ref, top_db, gmax, reset_next);
    return cudaGetLastError();
}
cudaError_t run_db_floor_blocks(float* x, long long B, int n_bands, long long T, float coef, float amin, float ref,
                                float top_db, const float* gmax, float* block_min, float* reset_next, int* n_raised,
                                const PeakExchange& xchg, float* host_mirror, cudaStream_t s) {
    const long long n_slots = B * ((T + kMinBlockFrames - 1) / kMinBlockFrames);
    const long long grid = (n_slots + kSlotsPerCta - 1) / kSlotsPerCta;
    if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    db_floor_blocks_kernel<<<(unsigned)grid, kFloorThreads, 0, s>>>(x, n_slots, n_bands, T, coef, amin, ref, top_db, gmax,
                                                                     block_min, reset_next, n_raised, xchg, host_mirror);
    return cudaGetLastError();
}
cudaError_t run_from_db(const float* x, long long n, float ref, float div, float* out, cudaStream_t s) {
    from_db_kernel<<<grid_for(n, kThreads * 4), kThreads, 0, s>>>(x, n, ref, div, out);
    return cudaGetLastError();
}
cudaError_t run_dct(const float* x, long long rows, int n_in, const float* D, int n_out, float* out, cudaStream_t s) {
    const size_t smem = size_t(8) * n_in * 4;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(dct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    dct_kernel<<<(unsigned)((rows + 7) / 8), 128, smem, s>>>(x, rows, n_in, D, n_out, out);
    return cudaGetLastError();
}
cudaError_t run_mfcc_tail(const float* mel, long long B, int n_mels, long long T, const float* D, int n_mfcc,
                          const float* lifter, int apply_db, float amin, float ref, int use_top, float top_db,
                          const float* gmax, float* out, cudaStream_t s) {
    const int kt = 2 * ((n_mfcc + 7) / 8);  // coefficients per lane group (four groups per warp), even
    static const bool rows_only = getenv("MLXA_MFCC_TAIL_ROWS") != nullptr;  // the earlier kernel, for A/B runs
    if (!rows_only && kt <= 16 && size_t(n_mels) * (kTailFrames + 4 * kt) * 4 <= 113 * 1024 && B <= 65535) {
#define MLXA_TAIL(KT) return launch_mfcc_tail_tiled<KT>(mel, B, n_mels, T, D, n_mfcc, lifter, apply_db, amin, ref, use_top, top_db, gmax, out, s)
        switch (kt) {
            case 2: MLXA_TAIL(2);
            case 4: MLXA_TAIL(4);
            case 6: MLXA_TAIL(6);
            case 8: MLXA_TAIL(8);
            case 10: MLXA_TAIL(10);
            case 12: MLXA_TAIL(12);
            case 14: MLXA_TAIL(14);
            default: MLXA_TAIL(16);
        }
#undef MLXA_TAIL
    }
    const size_t smem = size_t(n_mels) * 33 * 4 + size_t(n_mels) * n_mfcc * 4;
    cudaError_t e = cudaFuncSetAttribute(mfcc_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid((unsigned)((T + 31) / 32), (unsigned)B);
    mfcc_tail_kernel<<<grid, 256, smem, s>>>(mel, n_mels, T, D, n_mfcc, lifter, apply_db, amin, ref, use_top, top_db, gmax, out);
    return cudaGetLastError();
}

}  // namespace mlxa
