// Probe: warp-shuffle throughput against shared-memory loads (both go through the SM's MIO/LSU path).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters) {
    __shared__ float s[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) s[i] = i;
    __syncthreads();
    float a[8];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x + i;
    const int lane = threadIdx.x & 31, src = (32 - lane) & 31;
    int idx = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = __shfl_sync(0xffffffffu, a[i], src);            // arbitrary-lane shuffle
            if (MODE == 1) a[i] = __shfl_xor_sync(0xffffffffu, a[i], 16);
            if (MODE == 2) a[i] += s[(idx + 32 * i) & 4095];                       // LDS.32, conflict-free
            if (MODE == 3) { float2 v = reinterpret_cast<float2*>(s)[(idx + 32 * i) & 2047]; a[i] += v.x + v.y; }  // LDS.64
        }
        idx += 257;
    }
    float r = 0; for (int i = 0; i < 8; ++i) r += a[i];
    if (r == 123.456f) out[0] = r;
}
int main() {
    float* out; cudaMalloc(&out, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2048, blocks = 148 * 2, threads = 512;
    const char* names[] = {"SHFL idx", "SHFL xor", "LDS.32", "LDS.64"};
    for (int m = 0; m < 4; ++m) {
        auto launch = [&] {
            if (m == 0) k<0><<<blocks, threads>>>(out, iters);
            if (m == 1) k<1><<<blocks, threads>>>(out, iters);
            if (m == 2) k<2><<<blocks, threads>>>(out, iters);
            if (m == 3) k<3><<<blocks, threads>>>(out, iters);
        };
        launch(); cudaDeviceSynchronize();
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double warp_instr = 8.0 * iters * blocks * (threads / 32);
        printf("%-9s %7.3f ms  %.2f warp-instr/clk/SM (at 1.965 GHz)\n", names[m], ms, warp_instr / 148 / (ms * 1e-3 * 1.965e9));
    }
    return 0;
}
