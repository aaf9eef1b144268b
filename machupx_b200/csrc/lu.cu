// K3: dense FP64 LU factorisation and triangular solves (stands in for numpy.linalg.solve -> LAPACK
// dgesv at scene.py:827 and scene.py:896).
//
// Right-looking blocked LU on a row-major (ld x ld) matrix, ld a multiple of LU_NB, whose padding
// beyond n is the identity.  Per block step k:
//   diag   : factor the LU_NB x LU_NB diagonal block in shared memory (no pivoting; the lifting-line
//            matrices are diagonally dominated by 2|w_i|) and form inv(L_kk), inv(U_kk)
//   panel  : L21 = A21 inv(U_kk), U12 = inv(L_kk) A12 as DMMA tile products; y_k = inv(L_kk) b_k
//   update : A22 -= L21 U12 on DMMA tensor cores (mma.sync m8n8k4 f64), b_2 -= L21 y_k
// followed by a block back-substitution with the stored inverse diagonal blocks.
#include "lls_internal.h"

namespace lls {

constexpr int NB = LU_NB;          // 64
constexpr int TS = NB + 4;         // shared-memory row stride (doubles): conflict-free DMMA fragment loads
constexpr int TILE_BYTES = NB * TS * 8;

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// 64x64 tile global -> shared (row stride TS), optional scale; 128 threads
__device__ __forceinline__ void load_tile(double* __restrict__ s, const double* __restrict__ g, size_t ldg, double scale) {
    const int t = threadIdx.x;
    const int c2 = t & 31, r0 = t >> 5;
#pragma unroll 4
    for (int r = r0; r < NB; r += 4) {
        double2 v = *reinterpret_cast<const double2*>(g + (size_t)r * ldg + 2 * c2);
        v.x *= scale;
        v.y *= scale;
        *reinterpret_cast<double2*>(s + r * TS + 2 * c2) = v;
    }
}

// acc (per warp 32x32 of the 64x64 tile) += sA(64 x 64, [m][k]) * sB(64 x 64, [k][n])
__device__ __forceinline__ void mma_tile(const double* __restrict__ sA, const double* __restrict__ sB, double (&acc)[4][4][2]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
    const int g = lane >> 2, q = lane & 3;
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

template <bool LOAD>
__device__ __forceinline__ void acc_io(double (&acc)[4][4][2], double* __restrict__ g, size_t ldg) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
    const int gq = lane >> 2, q = lane & 3;
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            double2* p = reinterpret_cast<double2*>(g + (size_t)(wm + mi * 8 + gq) * ldg + wn + ni * 8 + 2 * q);
            if (LOAD) {
                double2 v = *p;
                acc[mi][ni][0] = v.x;
                acc[mi][ni][1] = v.y;
            } else {
                *p = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
            }
        }
}

// ---- diagonal block ----------------------------------------------------------------------------
// One CTA (256 threads).  Shared: A (factors), Li (inverse of unit-lower L), Ui (inverse of U).
__global__ void __launch_bounds__(256) lu_diag_kernel(double* __restrict__ M, int ld, int k, double* __restrict__ inv,
                                                      double* __restrict__ scalars, int* __restrict__ status) {
    extern __shared__ double sm[];
    double* A = sm;
    double* Li = sm + NB * (NB + 1);
    double* Ui = sm + 2 * NB * (NB + 1);
    constexpr int S = NB + 1;
    double* blk = M + (size_t)k * NB * ld + (size_t)k * NB;
    const int t = threadIdx.x;
    for (int e = t; e < NB * NB; e += 256) A[(e >> 6) * S + (e & 63)] = blk[(size_t)(e >> 6) * ld + (e & 63)];
    __syncthreads();
    double minpiv = 1e300;
    bool bad = false;
    for (int p = 0; p < NB; ++p) {
        const double piv = A[p * S + p];
        minpiv = fmin(minpiv, fabs(piv));
        bad |= !(fabs(piv) > 0.0);  // zero or NaN
        const double rp = 1.0 / piv;
        if (t > p && t < NB) A[t * S + p] *= rp;
        __syncthreads();
        // trailing rank-1 update: rows p+1.., cols p+1..
        const int m = NB - 1 - p;
        for (int e = t; e < m * m; e += 256) {
            const int r = p + 1 + e / m, c = p + 1 + e % m;
            A[r * S + c] -= A[r * S + p] * A[p * S + c];
        }
        __syncthreads();
    }
    if (t == 0) {
        if (bad) atomicOr(status, LLS_STATUS_NAN);
        scalars[1] = (k == 0) ? minpiv : fmin(minpiv, scalars[1]);
    }
    // inverses by columns: threads 0..63 -> column of inv(L); threads 64..127 -> column of inv(U)
    if (t < NB) {
        const int c = t;
        for (int r = 0; r < NB; ++r) Li[r * S + c] = (r == c) ? 1.0 : 0.0;
        for (int r = c + 1; r < NB; ++r) {
            double s = 0.0;
            for (int q = c; q < r; ++q) s += A[r * S + q] * Li[q * S + c];
            Li[r * S + c] = -s;
        }
    } else if (t < 2 * NB) {
        const int c = t - NB;
        for (int r = 0; r < NB; ++r) Ui[r * S + c] = 0.0;
        Ui[c * S + c] = 1.0 / A[c * S + c];
        for (int r = c - 1; r >= 0; --r) {
            double s = 0.0;
            for (int q = r + 1; q <= c; ++q) s += A[r * S + q] * Ui[q * S + c];
            Ui[r * S + c] = -s / A[r * S + r];
        }
    }
    __syncthreads();
    double* gL = inv + (size_t)k * 2 * NB * NB;
    double* gU = gL + NB * NB;
    for (int e = t; e < NB * NB; e += 256) {
        const int r = e >> 6, c = e & 63;
        blk[(size_t)r * ld + c] = A[r * S + c];
        gL[e] = Li[r * S + c];
        gU[e] = Ui[r * S + c];
    }
}

// ---- panel: TRSMs as tile products with the inverse diagonal blocks ---------------------------------
// grid.x = 2 * rem (+1 if rhs): b < rem -> L21 block row (k+1+b); rem <= b < 2 rem -> U12 block column; last -> y_k
__global__ void __launch_bounds__(128) lu_panel_kernel(double* __restrict__ M, int ld, int k, int rem,
                                                       const double* __restrict__ inv, double* __restrict__ rhs) {
    extern __shared__ double sm[];
    double* sA = sm;
    double* sB = sm + NB * TS;
    const int b = blockIdx.x;
    const double* gL = inv + (size_t)k * 2 * NB * NB;
    const double* gU = gL + NB * NB;
    if (b == 2 * rem) {  // y_k = inv(L_kk) b_k
        __shared__ double y[NB];
        const int t = threadIdx.x;
        if (t < NB) {
            double s = 0.0;
            for (int q = 0; q <= t; ++q) s += gL[t * NB + q] * rhs[(size_t)k * NB + q];
            y[t] = s;
        }
        __syncthreads();
        if (t < NB) rhs[(size_t)k * NB + t] = y[t];
        return;
    }
    double* target;
    if (b < rem) {  // L21 = A21 * inv(U)
        target = M + (size_t)(k + 1 + b) * NB * ld + (size_t)k * NB;
        load_tile(sA, target, ld, 1.0);
        load_tile(sB, gU, NB, 1.0);
    } else {        // U12 = inv(L) * A12
        target = M + (size_t)k * NB * ld + (size_t)(k + 1 + (b - rem)) * NB;
        load_tile(sA, gL, NB, 1.0);
        load_tile(sB, target, ld, 1.0);
    }
    __syncthreads();
    double acc[4][4][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    mma_tile(sA, sB, acc);
    acc_io<false>(acc, target, ld);
}

// ---- trailing update A22 -= L21 U12 (and b_2 -= L21 y_k) ----------------------------------------------
__global__ void __launch_bounds__(128) lu_update_kernel(double* __restrict__ M, int ld, int k, double* __restrict__ rhs) {
    extern __shared__ double sm[];
    double* sA = sm;
    double* sB = sm + NB * TS;
    const int bi = k + 1 + blockIdx.y, bj = k + 1 + blockIdx.x;
    const double* gA = M + (size_t)bi * NB * ld + (size_t)k * NB;   // L21 block
    const double* gB = M + (size_t)k * NB * ld + (size_t)bj * NB;   // U12 block
    double* gC = M + (size_t)bi * NB * ld + (size_t)bj * NB;
    load_tile(sA, gA, ld, -1.0);
    load_tile(sB, gB, ld, 1.0);
    double acc[4][4][2];
    acc_io<true>(acc, gC, ld);
    __syncthreads();
    mma_tile(sA, sB, acc);
    acc_io<false>(acc, gC, ld);
    if (rhs != nullptr && blockIdx.x == 0 && threadIdx.x < NB) {
        const int r = threadIdx.x;
        double s = 0.0;
        for (int q = 0; q < NB; ++q) s += sA[r * TS + q] * rhs[(size_t)k * NB + q];  // sA holds -L21
        rhs[(size_t)bi * NB + r] += s;
    }
}

// ---- block back-substitution, one launch per block column k (from the last to the first) -------------
// every CTA i <= k forms x_k = inv(U_kk) y_k; CTA k stores it, CTAs i < k fold it into their y_i
__global__ void __launch_bounds__(NB) lu_backsolve_kernel(const double* __restrict__ M, int ld, int k,
                                                          const double* __restrict__ inv, double* __restrict__ rhs,
                                                          double* __restrict__ x) {
    __shared__ double xk[NB];
    const int t = threadIdx.x, i = blockIdx.x;
    const double* gU = inv + (size_t)k * 2 * NB * NB + NB * NB;
    double s = 0.0;
    for (int q = t; q < NB; ++q) s += gU[t * NB + q] * rhs[(size_t)k * NB + q];
    xk[t] = s;
    __syncthreads();
    if (i == k) {
        x[(size_t)k * NB + t] = s;
    } else {
        const double* U = M + (size_t)i * NB * ld + (size_t)k * NB;
        double a = 0.0;
        for (int q = 0; q < NB; ++q) a += U[(size_t)t * ld + q] * xk[q];
        rhs[(size_t)i * NB + t] -= a;
    }
}

// forward substitution with the stored inv(L_kk): CTA k stores y_k, CTAs i > k fold it into their b_i
__global__ void __launch_bounds__(NB) lu_forwardsolve_kernel(const double* __restrict__ M, int ld, int k,
                                                             const double* __restrict__ inv, double* __restrict__ rhs,
                                                             double* __restrict__ y) {
    __shared__ double yk[NB];
    const int t = threadIdx.x, i = k + blockIdx.x;
    const double* gL = inv + (size_t)k * 2 * NB * NB;
    double s = 0.0;
    for (int q = 0; q <= t; ++q) s += gL[t * NB + q] * rhs[(size_t)k * NB + q];
    yk[t] = s;
    __syncthreads();
    if (i == k) {
        y[(size_t)k * NB + t] = s;
    } else {
        const double* Lb = M + (size_t)i * NB * ld + (size_t)k * NB;
        double a = 0.0;
        for (int q = 0; q < NB; ++q) a += Lb[(size_t)t * ld + q] * yk[q];
        rhs[(size_t)i * NB + t] -= a;
    }
}

static bool g_attr_done = false;

int lu_solve_only(Scene* s, const double* M, double* rhs, cudaStream_t st) {
    const int ld = s->ld, nblk = ld / NB;
    double* y = s->partial;
    for (int k = 0; k < nblk; ++k) {
        lu_forwardsolve_kernel<<<nblk - k, NB, 0, st>>>(M, ld, k, s->inv, rhs, y);
        count_launch();
    }
    for (int k = nblk - 1; k >= 0; --k) {
        lu_backsolve_kernel<<<k + 1, NB, 0, st>>>(M, ld, k, s->inv, y, rhs);
        count_launch();
    }
    LLS_CUDA_TRY(cudaGetLastError());
    return LLS_OK;
}


int lu_factor_solve(Scene* s, double* M, double* rhs, double* x, cudaStream_t st) {
    const int ld = s->ld, nblk = ld / NB;
    if (!g_attr_done) {
        LLS_CUDA_TRY(cudaFuncSetAttribute(lu_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * NB * (NB + 1) * 8));
        LLS_CUDA_TRY(cudaFuncSetAttribute(lu_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * TILE_BYTES));
        LLS_CUDA_TRY(cudaFuncSetAttribute(lu_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * TILE_BYTES));
        g_attr_done = true;
    }
    for (int k = 0; k < nblk; ++k) {
        lu_diag_kernel<<<1, 256, 3 * NB * (NB + 1) * 8, st>>>(M, ld, k, s->inv, s->scalars, s->status);
        count_launch();
        const int rem = nblk - 1 - k;
        const int pgrid = 2 * rem + (rhs ? 1 : 0);
        if (pgrid > 0) {
            lu_panel_kernel<<<pgrid, 128, 2 * TILE_BYTES, st>>>(M, ld, k, rem, s->inv, rhs);
            count_launch();
        }
        if (rem > 0) {
            lu_update_kernel<<<dim3(rem, rem), 128, 2 * TILE_BYTES, st>>>(M, ld, k, rhs);
            count_launch();
        }
    }
    if (rhs) {
        for (int k = nblk - 1; k >= 0; --k) {
            lu_backsolve_kernel<<<k + 1, NB, 0, st>>>(M, ld, k, s->inv, rhs, x);
            count_launch();
        }
    }
    LLS_CUDA_TRY(cudaGetLastError());
    return LLS_OK;
}

}  // namespace lls
