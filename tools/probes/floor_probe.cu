// Probe: why does a 3008-CTA "read one float per CTA and exit" kernel take ~20 us?  Variants timed with events.
#include <cstdio>
#include <cuda_runtime.h>
#include <math_constants.h>
__device__ __forceinline__ float to_db_one(float x, float coef, float amin, float refc) {
    return (coef * 0.30102999566398120f) * __log2f(__fdividef(fmaxf(x, amin), refc));
}
// A: current skeleton (CTA per slot)
__global__ void kA(float* x, const float* gmax, float* bmin, int nblk, float top_db) {
    const float floor_db = to_db_one(__ldg(gmax), 10.f, 1e-10f, 1.f) - top_db;
    __shared__ float s_min;
    const long long b = blockIdx.y;
    for (long long blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        __syncthreads();
        if (threadIdx.x == 0) { s_min = bmin[b * nblk + blk]; bmin[b * nblk + blk] = CUDART_INF_F; }
        __syncthreads();
        if (to_db_one(s_min, 10.f, 1e-10f, 1.f) >= floor_db) continue;
        x[threadIdx.x] = floor_db;
    }
}
// B: thread per slot, flagged slots compacted into a list
__global__ void kB(const float* gmax, float* bmin, int n, float top_db, int* list, int* count) {
    const float floor_db = to_db_one(__ldg(gmax), 10.f, 1e-10f, 1.f) - top_db;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float m = bmin[i];
    bmin[i] = CUDART_INF_F;
    if (to_db_one(m, 10.f, 1e-10f, 1.f) < floor_db) list[atomicAdd(count, 1)] = i;
}
__global__ void kEmpty(float* x) { if (x == nullptr) x[0] = 0; }
int main() {
    const int B = 64, nblk = 47, n = B * nblk;
    float *x, *gmax, *bmin; int *list, *count;
    cudaMalloc(&x, 1 << 20); cudaMalloc(&gmax, 4); cudaMalloc(&bmin, n * 4); cudaMalloc(&list, n * 4); cudaMalloc(&count, 4);
    float one = 100.f; cudaMemcpy(gmax, &one, 4, cudaMemcpyHostToDevice);
    cudaMemset(bmin, 0x3f, n * 4);  // ~0.74
    cudaMemset(count, 0, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](const char* name, auto f) {
        for (int i = 0; i < 20; ++i) f();
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        for (int i = 0; i < 200; ++i) f();
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-28s %7.2f us per launch (%s)\n", name, ms * 1000 / 200, cudaGetErrorString(cudaGetLastError()));
    };
    run("empty 3008x256", [&] { kEmpty<<<dim3(47, 64), 256>>>(x); });
    run("empty 148x256", [&] { kEmpty<<<148, 256>>>(x); });
    run("A cta-per-slot 3008x256", [&] { kA<<<dim3(47, 64), 256>>>(x, gmax, bmin, nblk, 500.f); });
    run("A cta-per-slot 3008x64", [&] { kA<<<dim3(47, 64), 64>>>(x, gmax, bmin, nblk, 500.f); });
    run("B thread-per-slot 12x256", [&] { kB<<<(n + 255) / 256, 256>>>(gmax, bmin, n, 500.f, list, count); });
    run("B thread-per-slot 94x32", [&] { kB<<<(n + 31) / 32, 32>>>(gmax, bmin, n, 500.f, list, count); });
    return 0;
}
