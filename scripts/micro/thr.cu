// Single-warp issue throughput on sm_100a for the instruction mix of the in-CTA LU: independent DFMA, LDS.64, SHFL.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* clk, double seed, int iters) {
    __shared__ double sh[256];
    sh[threadIdx.x] = seed + threadIdx.x;
    __syncthreads();
    double y = seed * 0.5, z = 1e-9;
    double a[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) a[i] = seed + i + threadIdx.x;
    long long t0, t1;
    // 32 independent DFMA chains
    t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 32; ++i) a[i] = fma(a[i], y, z);
    }
    t1 = clock64();
    if (threadIdx.x == 0) clk[0] = t1 - t0;
    // LDS.64 broadcast-ish loads feeding DFMA
    t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 32; ++i) a[i] = fma(a[i], sh[(i * 8 + it) & 255], z);
    }
    t1 = clock64();
    if (threadIdx.x == 0) clk[1] = t1 - t0;
    // shuffles (double) independent
    t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 32; ++i) a[i] = __shfl_sync(0xffffffffu, a[i], (threadIdx.x & ~7) | (it & 7));
    }
    t1 = clock64();
    if (threadIdx.x == 0) clk[2] = t1 - t0;
    double s = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += a[i];
    out[threadIdx.x] = s;
}
int main() {
    double* out; long long* clk;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&clk, 16 * 8);
    const int iters = 64;
    for (int threads : {32, 128, 512}) {
        for (int rep = 0; rep < 2; ++rep) k<<<1, threads>>>(out, clk, 1.5, iters);
        long long h[16];
        cudaMemcpy(h, clk, sizeof h, cudaMemcpyDeviceToHost);
        printf("threads %d: dfma %.2f cyc/instr, lds+dfma %.2f cyc/pair, shfl.f64 %.2f cyc/op\n", threads,
               h[0] / (32.0 * iters), h[1] / (32.0 * iters), h[2] / (32.0 * iters));
    }
    return 0;
}
