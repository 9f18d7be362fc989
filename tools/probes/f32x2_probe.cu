// Probe: throughput of packed FP32 (FFMA2 / FADD2) against scalar FFMA / FADD on sm_100a, and of a mix with LDS.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
template <int MODE>
__global__ void k(float* out, int iters) {
    float s = threadIdx.x * 1e-3f;
    if (MODE == 0) {  // scalar FFMA, 16 chains
        float a[16];
        for (int i = 0; i < 16; ++i) a[i] = s + i;
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], 0.999f, 1e-4f);
        float r = 0; for (int i = 0; i < 16; ++i) r += a[i];
        if (r == 123.f) out[0] = r;
    } else if (MODE == 1) {  // FFMA2, 8 chains of pairs (same flops as mode 0)
        u64 a[8]; const u64 m = pk(0.999f, 0.998f), c = pk(1e-4f, 2e-4f);
        for (int i = 0; i < 8; ++i) a[i] = pk(s + i, s - i);
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fma2(a[i], m, c);
        u64 r = 0; for (int i = 0; i < 8; ++i) r ^= a[i];
        if (r == 123) out[0] = 1;
    } else if (MODE == 2) {  // scalar FADD 16 chains
        float a[16];
        for (int i = 0; i < 16; ++i) a[i] = s + i;
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = a[i] + 1e-4f;
        float r = 0; for (int i = 0; i < 16; ++i) r += a[i];
        if (r == 123.f) out[0] = r;
    } else if (MODE == 3) {  // FADD2 8 chains
        u64 a[8]; const u64 c = pk(1e-4f, 2e-4f);
        for (int i = 0; i < 8; ++i) a[i] = pk(s + i, s - i);
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = add2(a[i], c);
        u64 r = 0; for (int i = 0; i < 8; ++i) r ^= a[i];
        if (r == 123) out[0] = 1;
    } else if (MODE == 4) {  // FFMA2 16 chains (twice the flops of mode 1 per iteration)
        u64 a[16]; const u64 m = pk(0.999f, 0.998f), c = pk(1e-4f, 2e-4f);
        for (int i = 0; i < 16; ++i) a[i] = pk(s + i, s - i);
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fma2(a[i], m, c);
        u64 r = 0; for (int i = 0; i < 16; ++i) r ^= a[i];
        if (r == 123) out[0] = 1;
    }
}
int main() {
    float* out; cudaMalloc(&out, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096, blocks = 148 * 4, threads = 512;
    auto run = [&](const char* name, auto launch, double flops_per_thread_iter) {
        launch(); cudaDeviceSynchronize();
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-22s %8.3f ms  %7.2f TFLOP/s  (%s)\n", name, ms, flops_per_thread_iter * iters * blocks * threads / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    };
    run("FFMA x16", [&] { k<0><<<blocks, threads>>>(out, iters); }, 32);
    run("FFMA2 x8", [&] { k<1><<<blocks, threads>>>(out, iters); }, 32);
    run("FFMA2 x16", [&] { k<4><<<blocks, threads>>>(out, iters); }, 64);
    run("FADD x16", [&] { k<2><<<blocks, threads>>>(out, iters); }, 16);
    run("FADD2 x8", [&] { k<3><<<blocks, threads>>>(out, iters); }, 16);
    return 0;
}
