// K3: dense FP64 LU factorisation and triangular solves (stands in for numpy.linalg.solve -> LAPACK
// dgesv at scene.py:827 and scene.py:896).
//
// Right-looking blocked LU (block NB = 64) on a row-major (ld x ld) matrix, ld a multiple of NB, whose
// padding beyond n is the identity.  ONE launch per block step, with the next step's panel folded in
// (look-ahead inside the kernel):
//
//   lu_step_kernel(k):  every CTA owns one 64 x 64 tile (bi, bj) of the trailing matrix and does
//       C = A[bi,bj] - L[bi,k] U[k,bj]                         (DMMA mma.sync m8n8k4, FP64 tensor path)
//     and then, by role,
//       diag  (bi = bj = k+1) : factors C in registers (2-D cyclic 4 x 8 register tiles, one barrier per
//                               pivot), publishes the factors and releases flag[k+1]
//       L/U   (bj = k+1 / bi = k+1): waits for flag[k+1], then L21 = C U11^-1 or U12 = L11^-1 C as a
//                               register-resident forward substitution (one thread per row / column)
//       other : stores C.
//   The diag CTA is block 0 and the panel CTAs follow it in the grid order, so the producer of the flag
//   is resident before any consumer can spin on it.
//   The right-hand side rides along as an augmented column: b[bi] -= L[bi,k] y_k in the CTAs of the first
//   tile column, y_{k+1} = L11^-1 b[k+1] in the diag CTA.  Back substitution is a separate sweep.
//
// No pivoting: the lifting-line matrices are dominated by the 2|w_i| diagonal (DESIGN.md "LU"); the
// smallest pivot magnitude is tracked and zero/NaN pivots raise LLS_STATUS_NAN.
#include "lls_internal.h"

namespace lls {

constexpr int NB = LU_NB;          // 64
constexpr int TS = NB + 4;         // shared-memory row stride (doubles): conflict-free DMMA fragment loads
constexpr int TILE_BYTES = NB * TS * 8;
constexpr int LU_THREADS = 128;

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
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

// shared (row stride TS) -> 64x64 tile in global memory, coalesced
__device__ __forceinline__ void store_tile(double* __restrict__ g, size_t ldg, const double* __restrict__ s) {
    const int t = threadIdx.x;
    const int c2 = t & 31, r0 = t >> 5;
#pragma unroll 4
    for (int r = r0; r < NB; r += 4)
        *reinterpret_cast<double2*>(g + (size_t)r * ldg + 2 * c2) = *reinterpret_cast<const double2*>(s + r * TS + 2 * c2);
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

// accumulator fragments <-> a row-major tile (global with ldg, or shared with TS)
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

// ---- in-CTA factorisation of a 64 x 64 block held in shared memory (row stride TS) -------------------
// 128 threads as a 16 x 8 grid; thread (ty, tx) keeps rows ty + 16 i (i < 4) and columns tx + 8 j (j < 8) in
// registers (cyclic, so the shrinking active window stays balanced).  Per pivot p the owners publish row p and
// column p through a double-buffered shared line, everyone rescales and applies the rank-1 update: one barrier
// per pivot.  Returns min |pivot| and flags a zero / NaN pivot.
struct PivotLine {
    double row[2][NB];
    double col[2][NB];
};

__device__ __forceinline__ void factor_block(double* __restrict__ sC, PivotLine& pl, double& minpiv, bool& bad) {
    const int t = threadIdx.x, ty = t >> 3, tx = t & 7;
    double a[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) a[i][j] = sC[(ty + 16 * i) * TS + tx + 8 * j];
    minpiv = 1e300;
    bad = false;
#pragma unroll 1
    for (int p = 0; p < NB; ++p) {
        const int buf = p & 1;
        const int pi = p >> 4, pty = p & 15, pj = p >> 3, ptx = p & 7;
        if (ty == pty) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (i == pi) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) pl.row[buf][tx + 8 * j] = a[i][j];
                }
        }
        if (tx == ptx) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j == pj) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) pl.col[buf][ty + 16 * i] = a[i][j];
                }
        }
        __syncthreads();
        const double piv = pl.row[buf][p];
        minpiv = fmin(minpiv, fabs(piv));
        bad |= !(fabs(piv) > 0.0);
        const double rp = 1.0 / piv;
        double r8[8], c4[4];
#pragma unroll
        for (int j = 0; j < 8; ++j) r8[j] = pl.row[buf][tx + 8 * j];
#pragma unroll
        for (int i = 0; i < 4; ++i) c4[i] = pl.col[buf][ty + 16 * i] * rp;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = ty + 16 * i;
            if (row > p) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int col = tx + 8 * j;
                    if (col > p) a[i][j] = fma(-c4[i], r8[j], a[i][j]);
                    else if (col == p) a[i][j] = c4[i];
                }
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) sC[(ty + 16 * i) * TS + tx + 8 * j] = a[i][j];
    __syncthreads();
}

// forward substitution  T x = c  with a lower-triangular T held column-major in shared memory
// (T(r, q) at sT[q * TS + r]) and its reciprocal diagonal dinv; x lives in registers (one system per thread).
__device__ __forceinline__ void trsv_lower64(const double* __restrict__ sT, const double* __restrict__ dinv, double (&x)[NB]) {
#pragma unroll
    for (int q = 0; q < NB; ++q) {
        x[q] *= dinv[q];
        const double xq = x[q];
#pragma unroll
        for (int r = q + 1; r < NB; ++r) x[r] = fma(-sT[q * TS + r], xq, x[r]);
    }
}

// y = L^-1 b for the unit-lower factor held row-major in sC; warp 0 only, lane l keeps b[l] and b[l + 32]
__device__ __forceinline__ void forward_rhs_warp(const double* __restrict__ sC, double* __restrict__ b) {
    const int lane = threadIdx.x;
    double x0 = b[lane], x1 = b[lane + 32];
#pragma unroll 1
    for (int q = 0; q < NB; ++q) {
        const double xq = __shfl_sync(0xffffffffu, (q < 32) ? x0 : x1, q & 31);
        if (lane > q) x0 = fma(-sC[lane * TS + q], xq, x0);
        if (lane + 32 > q) x1 = fma(-sC[(lane + 32) * TS + q], xq, x1);
    }
    b[lane] = x0;
    b[lane + 32] = x1;
}

// ---- one block step ------------------------------------------------------------------------------------
// FIRST: no trailing update (factor block 0 and its panel); otherwise update with panel k and factor panel k+1.
// Grid (1-D): block 0 = diag tile, 1..rem-1 = L tiles, rem..2rem-2 = U tiles, then the (rem-1)^2 others,
// where rem = number of block rows/cols of the trailing matrix handled here.
template <bool FIRST>
__global__ void __launch_bounds__(LU_THREADS, 3) lu_step_kernel(double* __restrict__ M, int ld, int k, int rem,
                                                                double* __restrict__ rhs, double* __restrict__ dbuf,
                                                                int* __restrict__ flags, int epoch,
                                                                double* __restrict__ scalars, int* __restrict__ status) {
    extern __shared__ __align__(16) double sm[];
    double* sA = sm;
    double* sB = sm + NB * TS;
    __shared__ PivotLine pl;
    __shared__ double dinv[NB];
    const int t = threadIdx.x;
    const int b = blockIdx.x;
    const int base = FIRST ? 0 : k + 1;  // first block row/col of the trailing matrix
    int bi, bj;
    enum { DIAG, LCOL, UROW, OTHER } role;
    if (b == 0) {
        role = DIAG; bi = base; bj = base;
    } else if (b < rem) {
        role = LCOL; bi = base + b; bj = base;
    } else if (b < 2 * rem - 1) {
        role = UROW; bi = base; bj = base + (b - rem + 1);
    } else {
        role = OTHER;
        const int o = b - (2 * rem - 1);
        bi = base + 1 + o / (rem - 1);
        bj = base + 1 + o % (rem - 1);
    }
    double* gC = M + (size_t)bi * NB * ld + (size_t)bj * NB;
    double acc[4][4][2];
    if (!FIRST) {
        load_tile(sA, M + (size_t)bi * NB * ld + (size_t)k * NB, ld, -1.0);   // -L[bi,k]
        load_tile(sB, M + (size_t)k * NB * ld + (size_t)bj * NB, ld, 1.0);    //  U[k,bj]
        acc_io<true>(acc, gC, ld);
        __syncthreads();
        mma_tile(sA, sB, acc);
        if (role == OTHER) {
            acc_io<false>(acc, gC, ld);
            return;
        }
        if (rhs != nullptr && bj == base && t < NB) {   // b[bi] -= L[bi,k] y_k   (sA holds -L)
            double s = 0.0;
            const double* yk = rhs + (size_t)k * NB;
#pragma unroll 8
            for (int q = 0; q < NB; ++q) s = fma(sA[t * TS + q], yk[q], s);
            rhs[(size_t)bi * NB + t] += s;
        }
        __syncthreads();
        acc_io<false>(acc, sB, TS);   // C, row-major in shared memory
        __syncthreads();
    } else {
        load_tile(sB, gC, ld, 1.0);
        __syncthreads();
    }
    double* sC = sB;
    const int blk = base;                       // index of the diagonal block being factored
    double* dl = dbuf + (size_t)(blk & 1) * NB * NB;

    if (role == DIAG) {
        double minpiv;
        bool bad;
        factor_block(sC, pl, minpiv, bad);
        store_tile(dl, NB, sC);
        __threadfence();
        __syncthreads();
        if (t == 0) st_release(flags + blk, epoch);
        store_tile(gC, ld, sC);
        if (t == 0) {
            if (bad) atomicOr(status, LLS_STATUS_NAN);
            scalars[1] = (blk == 0) ? minpiv : fmin(minpiv, scalars[1]);
        }
        if (rhs != nullptr && t < 32) forward_rhs_warp(sC, rhs + (size_t)blk * NB);
        return;
    }

    // panel roles: wait for the factored diagonal block
    if (t == 0) {
        while (ld_acquire(flags + blk) != epoch) __nanosleep(64);
    }
    __syncthreads();
    {
        // sT(r, q) at sA[q * TS + r]:  UROW needs T = L11 (unit lower), LCOL needs T = U11^T
        const int c = t & 63, r0 = t >> 6;
        for (int r = r0; r < NB; r += 2) {
            const double v = __ldcg(dl + r * NB + c);
            if (role == UROW) sA[c * TS + r] = v;   // T(r, c) = L(r, c), stored at [q = c][r]
            else sA[r * TS + c] = v;                // T(c, r) = U(r, c), stored at [q = r][c]
        }
        if (t < NB) dinv[t] = (role == UROW) ? 1.0 : 1.0 / __ldcg(dl + t * NB + t);
    }
    __syncthreads();
    if (t < NB) {
        const int se = (role == UROW) ? TS : 1, st = (role == UROW) ? 1 : TS;   // x[e] = C(e, t) or C(t, e)
        double x[NB];
#pragma unroll
        for (int e = 0; e < NB; ++e) x[e] = sC[e * se + t * st];
        trsv_lower64(sA, dinv, x);
#pragma unroll
        for (int e = 0; e < NB; ++e) sC[e * se + t * st] = x[e];
    }
    __syncthreads();
    store_tile(gC, ld, sC);
}

// ---- block back-substitution, one launch per block column k (from the last to the first) -------------
// every CTA i <= k solves U_kk x_k = y_k (warp-serial, 64 steps); CTA k stores x_k, CTAs i < k fold it into y_i
__global__ void __launch_bounds__(NB) lu_backsolve_kernel(const double* __restrict__ M, int ld, int k,
                                                          double* __restrict__ rhs, double* __restrict__ x) {
    __shared__ double xk[NB];
    __shared__ double sU[NB][NB + 1];
    const int t = threadIdx.x, i = blockIdx.x;
    const double* Ukk = M + (size_t)k * NB * ld + (size_t)k * NB;
    for (int r = 0; r < NB; ++r) sU[r][t] = Ukk[(size_t)r * ld + t];
    xk[t] = rhs[(size_t)k * NB + t];
    __syncthreads();
    if (t < 32) {
        double x0 = xk[t], x1 = xk[t + 32];
#pragma unroll 1
        for (int q = NB - 1; q >= 0; --q) {
            const double d = 1.0 / sU[q][q];
            double xq = __shfl_sync(0xffffffffu, (q < 32) ? x0 : x1, q & 31) * d;
            if (t == (q & 31)) {
                if (q < 32) x0 = xq; else x1 = xq;
            }
            if (t < q) x0 = fma(-sU[t][q], xq, x0);
            if (t + 32 < q) x1 = fma(-sU[t + 32][q], xq, x1);
        }
        xk[t] = x0;
        xk[t + 32] = x1;
    }
    __syncthreads();
    if (i == k) {
        x[(size_t)k * NB + t] = xk[t];
    } else {
        const double* U = M + (size_t)i * NB * ld + (size_t)k * NB;
        double a = 0.0;
        for (int q = 0; q < NB; ++q) a += U[(size_t)t * ld + q] * xk[q];
        rhs[(size_t)i * NB + t] -= a;
    }
}

// forward substitution for lls_lu_solve: every CTA i >= k solves L_kk y_k = b_k; CTA k stores it, the others fold it in
__global__ void __launch_bounds__(NB) lu_forwardsolve_kernel(const double* __restrict__ M, int ld, int k,
                                                             double* __restrict__ rhs, double* __restrict__ y) {
    __shared__ double yk[NB];
    __shared__ double sL[NB][NB + 1];
    const int t = threadIdx.x, i = k + blockIdx.x;
    const double* Lkk = M + (size_t)k * NB * ld + (size_t)k * NB;
    for (int r = 0; r < NB; ++r) sL[r][t] = Lkk[(size_t)r * ld + t];
    yk[t] = rhs[(size_t)k * NB + t];
    __syncthreads();
    if (t < 32) {
        double x0 = yk[t], x1 = yk[t + 32];
#pragma unroll 1
        for (int q = 0; q < NB; ++q) {
            const double xq = __shfl_sync(0xffffffffu, (q < 32) ? x0 : x1, q & 31);
            if (t > q) x0 = fma(-sL[t][q], xq, x0);
            if (t + 32 > q) x1 = fma(-sL[t + 32][q], xq, x1);
        }
        yk[t] = x0;
        yk[t + 32] = x1;
    }
    __syncthreads();
    if (i == k) {
        y[(size_t)k * NB + t] = yk[t];
    } else {
        const double* Lb = M + (size_t)i * NB * ld + (size_t)k * NB;
        double a = 0.0;
        for (int q = 0; q < NB; ++q) a += Lb[(size_t)t * ld + q] * yk[q];
        rhs[(size_t)i * NB + t] -= a;
    }
}

static bool g_attr_done = false;

static int lu_prepare(Scene* s) {
    if (!g_attr_done) {
        LLS_CUDA_TRY(cudaFuncSetAttribute(lu_step_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * TILE_BYTES));
        LLS_CUDA_TRY(cudaFuncSetAttribute(lu_step_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * TILE_BYTES));
        g_attr_done = true;
    }
    (void)s;
    return LLS_OK;
}

int lu_solve_only(Scene* s, const double* M, double* rhs, cudaStream_t st) {
    const int ld = s->ld, nblk = ld / NB;
    double* y = s->partial;
    for (int k = 0; k < nblk; ++k) {
        lu_forwardsolve_kernel<<<nblk - k, NB, 0, st>>>(M, ld, k, rhs, y);
        count_launch();
    }
    for (int k = nblk - 1; k >= 0; --k) {
        lu_backsolve_kernel<<<k + 1, NB, 0, st>>>(M, ld, k, y, rhs);
        count_launch();
    }
    LLS_CUDA_TRY(cudaGetLastError());
    return LLS_OK;
}

int lu_factor_solve(Scene* s, double* M, double* rhs, double* x, cudaStream_t st) {
    const int ld = s->ld, nblk = ld / NB;
    LLS_TRY(lu_prepare(s));
    const int epoch = ++s->lu_epoch;
    double* dbuf = s->inv;          // 2 x (NB x NB) published diagonal factors
    int* flags = s->lu_flags;
    {
        const int rem = nblk;       // block 0 and its panel
        lu_step_kernel<true><<<2 * rem - 1, LU_THREADS, 2 * TILE_BYTES, st>>>(M, ld, -1, rem, rhs, dbuf, flags, epoch, s->scalars, s->status);
        count_launch();
    }
    for (int k = 0; k + 1 < nblk; ++k) {
        const int rem = nblk - 1 - k;
        lu_step_kernel<false><<<rem * rem, LU_THREADS, 2 * TILE_BYTES, st>>>(M, ld, k, rem, rhs, dbuf, flags, epoch, s->scalars, s->status);
        count_launch();
    }
    if (rhs) {
        for (int k = nblk - 1; k >= 0; --k) {
            lu_backsolve_kernel<<<k + 1, NB, 0, st>>>(M, ld, k, rhs, x);
            count_launch();
        }
    }
    LLS_CUDA_TRY(cudaGetLastError());
    return LLS_OK;
}

}  // namespace lls
