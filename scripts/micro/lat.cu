// Dependent-issue latencies on sm_100a that bound the serial pivot chain of the in-CTA LU: DFMA, DMUL, 1/x (double),
// warp shuffle of a double, shared-memory round trip, __syncthreads with 4 warps.  Build: nvcc -arch=sm_100a lat.cu -o lat
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* clk, double seed) {
    __shared__ double sh[128];
    double x = seed + threadIdx.x * 1e-9, y = 1.0000001;
    long long t0, t1;
    int n = 0;
    // DFMA chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 256; ++i) x = fma(x, y, 1e-9);
    t1 = clock64();
    if (threadIdx.x == 0) clk[n] = (t1 - t0);
    n++;
    // DMUL chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 256; ++i) x = x * y;
    t1 = clock64();
    if (threadIdx.x == 0) clk[n] = (t1 - t0);
    n++;
    // reciprocal chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) x = 1.0 / (x + 3.0);
    t1 = clock64();
    if (threadIdx.x == 0) clk[n] = (t1 - t0) * 4;
    n++;
    // shuffle chain (double = 2 x 32-bit)
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 256; ++i) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
    t1 = clock64();
    if (threadIdx.x == 0) clk[n] = (t1 - t0);
    n++;
    // shared store -> sync -> load chain
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; ++i) {
        sh[threadIdx.x] = x;
        __syncthreads();
        x = sh[(threadIdx.x + 1) & 127] + 1.0;
    }
    t1 = clock64();
    if (threadIdx.x == 0) clk[n] = (t1 - t0);
    n++;
    // bare __syncthreads
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 256; ++i) __syncthreads();
    t1 = clock64();
    if (threadIdx.x == 0) clk[n] = (t1 - t0);
    n++;
    // independent DFMA throughput, 1 warp: 8 chains
    double a0 = x, a1 = x + 1, a2 = x + 2, a3 = x + 3, a4 = x + 4, a5 = x + 5, a6 = x + 6, a7 = x + 7;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        a0 = fma(a0, y, 1e-9); a1 = fma(a1, y, 1e-9); a2 = fma(a2, y, 1e-9); a3 = fma(a3, y, 1e-9);
        a4 = fma(a4, y, 1e-9); a5 = fma(a5, y, 1e-9); a6 = fma(a6, y, 1e-9); a7 = fma(a7, y, 1e-9);
    }
    t1 = clock64();
    if (threadIdx.x == 0) clk[n] = (t1 - t0);
    n++;
    out[threadIdx.x] = x + a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
int main() {
    double* out; long long* clk;
    cudaMalloc(&out, 128 * 8); cudaMalloc(&clk, 16 * 8);
    for (int rep = 0; rep < 2; ++rep) k<<<1, 128>>>(out, clk, 1.5);
    long long h[16];
    cudaMemcpy(h, clk, sizeof h, cudaMemcpyDeviceToHost);
    const char* names[] = {"dfma_dep", "dmul_dep", "recip_dep", "shfl_double_dep", "sts_sync_lds_dadd", "syncthreads", "dfma_8chains_per_op"};
    for (int i = 0; i < 7; ++i) printf("%s: %.1f cycles/op\n", names[i], h[i] / 256.0);
    return 0;
}
