// Which SM does block b of a 1-D grid land on (3 CTAs/SM by shared memory)?  Decides whether the LU's panel roles can be packed
// three to an SM by block index alone.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* smid, long long* t0) {
    extern __shared__ double sm[];
    if (threadIdx.x == 0) {
        unsigned s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
        smid[blockIdx.x] = (int)s;
        t0[blockIdx.x] = clock64();
    }
    // stay resident long enough for the whole first wave to be placed
    long long t = clock64();
    while (clock64() - t < 200000) { }
    if (sm[threadIdx.x] == 123.0) smid[0] = -1;
}
int main() {
    const int n = 1200;
    int* d; long long* dt;
    cudaMalloc(&d, n * 4); cudaMalloc(&dt, n * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 70 * 1024);
    for (int rep = 0; rep < 2; ++rep) k<<<n, 128, 70 * 1024>>>(d, dt);
    static int h[n];
    cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    int same1 = 0, same2 = 0;
    for (int b = 0; b < 148; ++b) { same1 += h[b] == h[b + 148]; same2 += h[b] == h[b + 296]; }
    printf("blocks b and b+148 on the same SM: %d of 148; b and b+296: %d of 148\n", same1, same2);
    for (int b = 0; b < 460; ++b) printf("%d%c", h[b], (b % 37 == 36) ? '\n' : ' ');
    printf("\n");
    return 0;
}
