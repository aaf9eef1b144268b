// FP64 tensor (DMMA) throughput on sm_100a as a function of resident warps per SM, register-fed and shared-memory-fed,
// for the two instruction shapes (m8n8k4, m16n8k16).  Decides the warp count / instruction shape of the LU trailing update.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/dmma_occ scripts/micro/dmma_occ.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>

constexpr int NB = 64, TS = NB + 4;

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma16816(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
        : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
        : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
__device__ __forceinline__ void dmma1688(double (&d)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
        : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

// ---- register-fed ----------------------------------------------------------------------------------------------------
__global__ void reg884(double* out, int iters, double seed) {
    double a = seed * 1e-3 + threadIdx.x * 1e-6, b = 1e-3;
    double c[16][2];
#pragma unroll
    for (int u = 0; u < 16; ++u) { c[u][0] = u; c[u][1] = -u; }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int u = 0; u < 16; ++u) dmma884(c[u][0], c[u][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int u = 0; u < 16; ++u) s += c[u][0] + c[u][1];
    if (s == 123.456) out[0] = s;
}
__global__ void reg16816(double* out, int iters, double seed) {
    double a[8], b[4];
#pragma unroll
    for (int u = 0; u < 8; ++u) a[u] = seed * 1e-3 + (threadIdx.x + u) * 1e-6;
#pragma unroll
    for (int u = 0; u < 4; ++u) b[u] = 1e-3 * (u + 1);
    double c[8][4];
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) c[u][v] = u - v;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) dmma16816(c[u], a, b);
    }
    double s = 0;
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) s += c[u][v];
    if (s == 123.456) out[0] = s;
}

// ---- shared-memory-fed 64 x 64 x 64 tile products (the LU trailing update's inner loop) ---------------------------------
// variant 0: m8n8k4, 4 warps, warp tile 32 x 32 (what lu.cu does)
// variant 1: m16n8k16, 4 warps, warp tile 32 x 32
// variant 2: m8n8k4, 8 warps, warp tile 32 x 16
// variant 3: m16n8k16, 8 warps, warp tile 32 x 16
// variant 4: m16n8k8, 4 warps, warp tile 32 x 32
template <int VARIANT>
__global__ void smem_tile(double* out, const double* in, int iters) {
    extern __shared__ __align__(16) double sm[];
    double* sA = sm;
    double* sB = sm + NB * TS;
    for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) {
        sA[(e >> 6) * TS + (e & 63)] = in[e];
        sB[(e >> 6) * TS + (e & 63)] = in[NB * NB + e];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    double s = 0;
    if (VARIANT == 0) {
        const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
        double acc[4][4][2] = {};
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
#pragma unroll 4
            for (int k0 = 0; k0 < NB; k0 += 4) {
                double a[4], b[4];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) a[mi] = sA[(wm + mi * 8 + g) * TS + k0 + q];
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) b[ni] = sB[(k0 + q) * TS + wn + ni * 8 + g];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
            }
        }
        // fragment (mi, ni): rows wm + mi*8 + g, cols wn + ni*8 + 2q, 2q+1
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                out[(size_t)blockIdx.x * NB * NB + (wm + mi * 8 + g) * NB + wn + ni * 8 + 2 * q] = acc[mi][ni][0];
                out[(size_t)blockIdx.x * NB * NB + (wm + mi * 8 + g) * NB + wn + ni * 8 + 2 * q + 1] = acc[mi][ni][1];
            }
    } else if (VARIANT == 2) {
        const int wm = (warp >> 2) * 32, wn = (warp & 3) * 16;
        double acc[4][2][2] = {};
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
#pragma unroll 4
            for (int k0 = 0; k0 < NB; k0 += 4) {
                double a[4], b[2];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) a[mi] = sA[(wm + mi * 8 + g) * TS + k0 + q];
#pragma unroll
                for (int ni = 0; ni < 2; ++ni) b[ni] = sB[(k0 + q) * TS + wn + ni * 8 + g];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 2; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
            }
        }
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) {
                out[(size_t)blockIdx.x * NB * NB + (wm + mi * 8 + g) * NB + wn + ni * 8 + 2 * q] = acc[mi][ni][0];
                out[(size_t)blockIdx.x * NB * NB + (wm + mi * 8 + g) * NB + wn + ni * 8 + 2 * q + 1] = acc[mi][ni][1];
            }
    } else if (VARIANT == 1 || VARIANT == 3) {
        constexpr int NJ = (VARIANT == 1) ? 4 : 2;
        const int wm = (VARIANT == 1) ? (warp >> 1) * 32 : (warp >> 2) * 32;
        const int wn = (VARIANT == 1) ? (warp & 1) * 32 : (warp & 3) * 16;
        double acc[2][NJ][4] = {};
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k0 = 0; k0 < NB; k0 += 16) {
                double a[2][8], b[NJ][4];
#pragma unroll
                for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                    for (int i = 0; i < 8; ++i) a[mi][i] = sA[(wm + mi * 16 + g + 8 * (i & 1)) * TS + k0 + q + 4 * (i >> 1)];
#pragma unroll
                for (int ni = 0; ni < NJ; ++ni)
#pragma unroll
                    for (int i = 0; i < 4; ++i) b[ni][i] = sB[(k0 + q + 4 * i) * TS + wn + ni * 8 + g];
#pragma unroll
                for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                    for (int ni = 0; ni < NJ; ++ni) dmma16816(acc[mi][ni], a[mi], b[ni]);
            }
        }
        // c[i]: row g + 8 (i >> 1), col 2q + (i & 1)
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < NJ; ++ni)
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    out[(size_t)blockIdx.x * NB * NB + (wm + mi * 16 + g + 8 * (i >> 1)) * NB + wn + ni * 8 + 2 * q + (i & 1)] = acc[mi][ni][i];
    } else if (VARIANT == 4) {
        const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
        double acc[2][4][4] = {};
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
#pragma unroll 2
            for (int k0 = 0; k0 < NB; k0 += 8) {
                double a[2][4], b[4][2];
#pragma unroll
                for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                    for (int i = 0; i < 4; ++i) a[mi][i] = sA[(wm + mi * 16 + g + 8 * (i & 1)) * TS + k0 + q + 4 * (i >> 1)];
#pragma unroll
                for (int ni = 0; ni < 4; ++ni)
#pragma unroll
                    for (int i = 0; i < 2; ++i) b[ni][i] = sB[(k0 + q + 4 * i) * TS + wn + ni * 8 + g];
#pragma unroll
                for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) dmma1688(acc[mi][ni], a[mi], b[ni]);
            }
        }
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni)
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    out[(size_t)blockIdx.x * NB * NB + (wm + mi * 16 + g + 8 * (i >> 1)) * NB + wn + ni * 8 + 2 * q + (i & 1)] = acc[mi][ni][i];
    }
    if (s == 123.456) out[0] = s;
}

template <class F>
static float timeit(F&& f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

template <int VARIANT>
static void run_smem(const char* name, int threads, const double* d_in, double* d_out, const double* h_ref, int sms) {
    cudaFuncSetAttribute(smem_tile<VARIANT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    // correctness: one iteration against the host product
    smem_tile<VARIANT><<<1, threads, 2 * NB * TS * 8>>>(d_out, d_in, 1);
    static double h[NB * NB];
    cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost);
    double err = 0;
    for (int e = 0; e < NB * NB; ++e) err = fmax(err, fabs(h[e] - h_ref[e]));
    printf("%s: max abs err vs host product %.3e (%s)\n", name, err, cudaGetErrorString(cudaGetLastError()));
    const int iters = 200;
    for (int per_sm : {1, 2, 3, 4, 6}) {
        // dynamic shared memory sized so that exactly per_sm CTAs fit per SM (227 KB usable, 1 KB reserved per CTA)
        int smem = 227 * 1024 / per_sm - 2048;
        if (smem < 2 * NB * TS * 8) smem = 2 * NB * TS * 8;
        int occ = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, smem_tile<VARIANT>, threads, smem);
        const int grid = sms * occ;
        float ms = timeit([&] { smem_tile<VARIANT><<<grid, threads, smem>>>(d_out, d_in, iters); });
        double tf = (double)grid * iters * 2.0 * NB * NB * NB / (ms * 1e-3) * 1e-12;
        printf("  %s  CTAs/SM %d (%d warps/SM): %.2f TFLOP/s\n", name, occ, occ * threads / 32, tf);
    }
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *d_out, *d_in;
    cudaMalloc(&d_out, (size_t)sms * 8 * NB * NB * 8);
    cudaMalloc(&d_in, 2 * NB * NB * 8);
    static double h_in[2 * NB * NB], h_ref[NB * NB];
    srand(1);
    for (int e = 0; e < 2 * NB * NB; ++e) h_in[e] = (rand() % 2001 - 1000) * 1e-3;
    for (int i = 0; i < NB; ++i)
        for (int j = 0; j < NB; ++j) {
            double s = 0;
            for (int k = 0; k < NB; ++k) s += h_in[i * NB + k] * h_in[NB * NB + k * NB + j];
            h_ref[i * NB + j] = s;
        }
    cudaMemcpy(d_in, h_in, sizeof h_in, cudaMemcpyHostToDevice);
    printf("SMs %d\n", sms);
    const int iters = 4096;
    for (int warps : {4, 8, 12, 16, 32, 64}) {
        const int threads = warps >= 32 ? 1024 : warps * 32, ctas = warps >= 32 ? warps / 32 : 1;
        float ms = timeit([&] { reg884<<<sms * ctas, threads>>>(d_out, iters, 1.0); });
        double tf = (double)sms * warps * iters * 64.0 * 512.0 / (ms * 1e-3) * 1e-12;
        float ms2 = timeit([&] { reg16816<<<sms * ctas, threads>>>(d_out, iters, 1.0); });
        double tf2 = (double)sms * warps * iters * 8.0 * 4096.0 / (ms2 * 1e-3) * 1e-12;
        printf("register-fed, %2d warps/SM: m8n8k4 %.2f TFLOP/s, m16n8k16 %.2f TFLOP/s\n", warps, tf, tf2);
    }
    run_smem<0>("smem m8n8k4   4 warps 32x32", 128, d_in, d_out, h_ref, sms);
    run_smem<1>("smem m16n8k16 4 warps 32x32", 128, d_in, d_out, h_ref, sms);
    run_smem<4>("smem m16n8k8  4 warps 32x32", 128, d_in, d_out, h_ref, sms);
    run_smem<2>("smem m8n8k4   8 warps 32x16", 256, d_in, d_out, h_ref, sms);
    run_smem<3>("smem m16n8k16 8 warps 32x16", 256, d_in, d_out, h_ref, sms);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
