// FP64 peak probes for the roofline denominators (B200, sm_100a).
//
// MEASURED_PEAKS.json carries only the HBM copy bandwidth and the bf16 GEMM rate; the
// lifting-line path is FP64 throughout, so bench.py measures the FP64 ceilings live with
// these kernels:  (1) dependent-chain-free DFMA streams (vector FP64 pipe) and
// (2) DMMA mma.sync m8n8k4 / m16n8k16 streams (FP64 tensor path).  Nothing here touches
// memory in the timed loop, so the numbers are issue/pipe ceilings.
#include <cuda_runtime.h>
#include <cstdint>
#include "lls_internal.h"

namespace lls {

__global__ void __launch_bounds__(256) probe_dfma_kernel(double* out, int iters, double seed) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1.0, a2 = a0 + 2.0, a3 = a0 + 3.0;
    double a4 = a0 + 4.0, a5 = a0 + 5.0, a6 = a0 + 6.0, a7 = a0 + 7.0;
    const double m = 0.999999, c = 1e-9;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456) out[0] = s;  // never true; keeps the chains alive
}

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__device__ __forceinline__ void dmma_m16n8k16(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
        "{%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
        : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
        : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
          "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

__global__ void __launch_bounds__(256) probe_dmma884_kernel(double* out, int iters, double seed) {
    double a = seed * 1e-3 + threadIdx.x * 1e-6, b = 1e-3;
    double c[8][2];
#pragma unroll
    for (int u = 0; u < 8; ++u) { c[u][0] = u; c[u][1] = -u; }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int u = 0; u < 8; ++u) dmma_m8n8k4(c[u][0], c[u][1], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int u = 0; u < 8; ++u) s += c[u][0] + c[u][1];
    if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256) probe_dmma16816_kernel(double* out, int iters, double seed) {
    double a[8], b[4];
#pragma unroll
    for (int u = 0; u < 8; ++u) a[u] = seed * 1e-3 + (threadIdx.x + u) * 1e-6;
#pragma unroll
    for (int u = 0; u < 4; ++u) b[u] = 1e-3 * (u + 1);
    double c[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) c[u][v] = u - v;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
#pragma unroll
            for (int u = 0; u < 4; ++u) dmma_m16n8k16(c[u], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) s += c[u][v];
    if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256) probe_copy_kernel(const double2* __restrict__ src, double2* __restrict__ dst, size_t n2) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n2; i += stride) dst[i] = src[i];
}

template <class Launch>
static float time_launch(cudaStream_t st, int reps, Launch&& launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    launch();  // warm-up
    cudaStreamSynchronize(st);
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0, st);
        launch();
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return best;
}

}  // namespace lls

using namespace lls;

// out[0] = DFMA TFLOP/s, out[1] = DMMA m8n8k4 TFLOP/s, out[2] = DMMA m16n8k16 TFLOP/s,
// out[3] = device copy GB/s (read+write bytes, 1 GiB buffers)
extern "C" int lls_probe_peaks(double* out4, void* stream_v) {
    cudaStream_t st = (cudaStream_t)stream_v;
    int dev = 0, sms = 0;
    LLS_CUDA_TRY(cudaGetDevice(&dev));
    LLS_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double* scratch = nullptr;
    LLS_CUDA_TRY(cudaMalloc(&scratch, 64));
    const int grid = sms * 8, block = 256;

    {   // burst figure: ~1 ms kernels
        const int iters = 2048;
        float ms = time_launch(st, 5, [&] { probe_dfma_kernel<<<grid, block, 0, st>>>(scratch, iters, 1.0); });
        out4[0] = (double)grid * block * iters * 64.0 * 2.0 / (ms * 1e-3) * 1e-12;
    }
    {
        const int iters = 2048;
        float ms = time_launch(st, 5, [&] { probe_dmma884_kernel<<<grid, block, 0, st>>>(scratch, iters, 1.0); });
        out4[1] = (double)grid * (block / 32) * iters * 32.0 * 512.0 / (ms * 1e-3) * 1e-12;
    }
    {
        const int iters = 2048;
        float ms = time_launch(st, 5, [&] { probe_dmma16816_kernel<<<grid, block, 0, st>>>(scratch, iters, 1.0); });
        out4[2] = (double)grid * (block / 32) * iters * 8.0 * (2.0 * 16 * 8 * 16) / (ms * 1e-3) * 1e-12;
    }
    {
        const size_t bytes = (size_t)1 << 30;
        double2 *a = nullptr, *b = nullptr;
        if (cudaMalloc(&a, bytes) == cudaSuccess && cudaMalloc(&b, bytes) == cudaSuccess) {
            cudaMemsetAsync(a, 0, bytes, st);
            float ms = time_launch(st, 5, [&] { probe_copy_kernel<<<sms * 16, 256, 0, st>>>(a, b, bytes / 16); });
            out4[3] = 2.0 * bytes / (ms * 1e-3) * 1e-9;
        } else {
            out4[3] = 0.0;
        }
        cudaFree(a);
        cudaFree(b);
    }
    LLS_CUDA_TRY(cudaGetLastError());
    cudaFree(scratch);
    return LLS_OK;
}
