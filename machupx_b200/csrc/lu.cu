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
#include <cstdlib>

namespace lls {

#ifdef LLS_LU_CLOCKS
// debug build only (scripts/lu_clocks.py): clock64 stamps of the diag CTA (slots 0..8) and of the first L-panel CTA (slots 12..20)
// of every block step, to see how long the serial chain takes when it shares its SMs with trailing-update CTAs
__device__ long long g_lu_clocks[256][24];
#define LU_CLK(slot)                                                                                  \
    do {                                                                                              \
        if (threadIdx.x == 0 && base < 256) {                                                         \
            if (role == DIAG) g_lu_clocks[base][slot] = clock64();                                    \
            else if (role == LCOL && bi == base + 1) g_lu_clocks[base][12 + slot] = clock64();        \
        }                                                                                             \
    } while (0)
#else
#define LU_CLK(slot) do { } while (0)
#endif

constexpr int NB = LU_NB;          // 64
constexpr int TS = NB + 4;         // shared-memory row stride (doubles): conflict-free DMMA fragment loads
constexpr int TILE_BYTES = NB * TS * 8;
constexpr int LU_THREADS = 128;
constexpr int PUB_DOUBLES = NB * TS + NB;       // published diagonal-block image: LU (padded rows, upper part scaled), 1/diag(U)
static_assert((NB * TS / 2) % LU_THREADS == 0, "image copy assumes a whole number of double2 per thread");

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

// Everything below is written for CTAs of NT = 128 threads (4 warps, warp tile 32 x 32 of the 64 x 64 tile) or NT = 256 threads
// (8 warps, warp tile 32 x 16).  In a 256-thread CTA the panel roles (factorisation, panel solves) still run on threads 0..127:
// they meet at named barrier 1 (panel_sync) while the upper four warps wait at barrier 0.
template <int NT>
__device__ __forceinline__ void panel_sync() {
    if (NT == 128) __syncthreads();
    else asm volatile("bar.sync 1, 128;\n" ::: "memory");
}

// 64x64 tile global -> shared (row stride TS) with cp.async (16 bytes per request, all requests of a thread in flight)
template <int NT = LU_THREADS>
__device__ __forceinline__ void load_tile_async(double* __restrict__ s, const double* __restrict__ g, size_t ldg) {
    const int t = threadIdx.x;
    const int c2 = t & 31, r0 = t >> 5;
#pragma unroll
    for (int r = r0; r < NB; r += NT / 32) {
        const unsigned dst = (unsigned)__cvta_generic_to_shared(s + r * TS + 2 * c2);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(g + (size_t)r * ldg + 2 * c2) : "memory");
    }
}
__device__ __forceinline__ void async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory"); }

// shared (row stride TS) -> 64x64 tile in global memory, coalesced (threads 0..127)
__device__ __forceinline__ void store_tile(double* __restrict__ g, size_t ldg, const double* __restrict__ s) {
    const int t = threadIdx.x;
    const int c2 = t & 31, r0 = t >> 5;
#pragma unroll 4
    for (int r = r0; r < NB; r += 4)
        *reinterpret_cast<double2*>(g + (size_t)r * ldg + 2 * c2) = *reinterpret_cast<const double2*>(s + r * TS + 2 * c2);
}

template <int NT>
struct WarpTile {
    static constexpr int NI = (NT == 128) ? 4 : 2;       // 8-column fragments per warp
    __device__ static __forceinline__ int wm() { const int warp = threadIdx.x >> 5; return (NT == 128) ? (warp >> 1) * 32 : (warp >> 2) * 32; }
    __device__ static __forceinline__ int wn() { const int warp = threadIdx.x >> 5; return (NT == 128) ? (warp & 1) * 32 : (warp & 3) * 16; }
};

// acc (per warp 32 x 32 or 32 x 16 of the 64x64 tile) += sA(64 x 64, [m][k]) * sB(64 x 64, [k][n])
template <int NT = LU_THREADS>
__device__ __forceinline__ void mma_tile(const double* __restrict__ sA, const double* __restrict__ sB, double (&acc)[4][WarpTile<NT>::NI][2]) {
    constexpr int NI = WarpTile<NT>::NI;
    const int lane = threadIdx.x & 31;
    const int wm = WarpTile<NT>::wm(), wn = WarpTile<NT>::wn();
    const int g = lane >> 2, q = lane & 3;
#pragma unroll 4
    for (int k0 = 0; k0 < NB; k0 += 4) {
        double a[4], b[NI];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) a[mi] = sA[(wm + mi * 8 + g) * TS + k0 + q];
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) b[ni] = sB[(k0 + q) * TS + wn + ni * 8 + g];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < NI; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
    }
}

// accumulator fragments <-> a row-major tile (global with ldg, or shared with TS)
// LOAD negates on the way in and the store negates on the way out: the tile product is accumulated as
// -(C) + L U, so neither operand tile needs a sign flip in shared memory
template <bool LOAD, int NT = LU_THREADS>
__device__ __forceinline__ void acc_io(double (&acc)[4][WarpTile<NT>::NI][2], double* __restrict__ g, size_t ldg) {
    constexpr int NI = WarpTile<NT>::NI;
    const int lane = threadIdx.x & 31;
    const int wm = WarpTile<NT>::wm(), wn = WarpTile<NT>::wn();
    const int gq = lane >> 2, q = lane & 3;
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) {
            double2* p = reinterpret_cast<double2*>(g + (size_t)(wm + mi * 8 + gq) * ldg + wn + ni * 8 + 2 * q);
            if (LOAD) {
                double2 v = *p;
                acc[mi][ni][0] = -v.x;
                acc[mi][ni][1] = -v.y;
            } else {
                *p = make_double2(-acc[mi][ni][0], -acc[mi][ni][1]);
            }
        }
}

// The C tile of a trailing update as raw fragments, fetched with volatile loads so they are ISSUED where they are written (ahead
// of the operand copies): ptxas otherwise sinks plain loads below the cp.async wait and the two memory latencies of a tile
// (operands, then C) add up instead of overlapping.
template <int NT = LU_THREADS>
__device__ __forceinline__ void c_fetch(double2 (&cn)[4][WarpTile<NT>::NI], const double* __restrict__ g, size_t ldg) {
    const int lane = threadIdx.x & 31;
    const int wm = WarpTile<NT>::wm(), wn = WarpTile<NT>::wn();
    const int gq = lane >> 2, q = lane & 3;
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < WarpTile<NT>::NI; ++ni) {
            const double* p = g + (size_t)(wm + mi * 8 + gq) * ldg + wn + ni * 8 + 2 * q;
            asm volatile("ld.global.v2.f64 {%0, %1}, [%2];\n" : "=d"(cn[mi][ni].x), "=d"(cn[mi][ni].y) : "l"(p) : "memory");
        }
}
template <int NT = LU_THREADS>
__device__ __forceinline__ void c_negate(double (&acc)[4][WarpTile<NT>::NI][2], const double2 (&cn)[4][WarpTile<NT>::NI]) {
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < WarpTile<NT>::NI; ++ni) {
            acc[mi][ni][0] = -cn[mi][ni].x;
            acc[mi][ni][1] = -cn[mi][ni].y;
        }
}

// ---- in-CTA factorisation of a 64 x 64 block held in shared memory (row stride TS) -------------------
// 128 threads as a 16 x 8 grid; thread (ty, tx) keeps rows ty + 16 i (i < 4) and columns tx + 8 j (j < 8) in
// registers (cyclic, so the shrinking active window stays balanced).  Per pivot p the owners of row p and column p
// publish them through a double-buffered shared line ALREADY MASKED (row entries left of / on the pivot and column
// entries above / on it are zero), so the rank-1 update needs no compares: one barrier, a handful of 16-byte shared
// loads, one reciprocal and (4 - I)(8 - J) DFMA per pivot.  The pivots are walked in 8 groups of 8 with the register
// row I = p / 16 and column J = p / 8 that hold the pivot known at compile time, so no register array is ever indexed
// dynamically and dead register rows / columns drop out of the update.  Multipliers (L) go straight to shared memory;
// the registers end up holding U.
struct PivotLine {
    double row[2][NB];    // [buf][tx * 8 + j]   masked pivot row
    double col[2][NB];    // [buf][ty * 4 + i]   masked pivot column (unscaled)
    double rcp[NB];       // reciprocal pivots
    double piv[2];
};

// 1 / d to ~1e-18 relative: hardware seed (reciprocal of the high word, ~2^-20) + one cubic Newton step.  Not
// correctly rounded (neither needed nor expected of LU multipliers); 3 dependent DFMA instead of the 5 of `1.0 / d`.
__device__ __forceinline__ double fast_rcp(double d) {
    double x0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x0) : "d"(d));
    const double e = fma(-d, x0, 1.0);
    const double q = fma(e, e, e);
    return fma(x0, q, x0);
}

template <int J, int NT>
__device__ __forceinline__ void factor_group(double (&a)[4][8], PivotLine& pl, int ty, int tx) {
    constexpr int I = J >> 1;
    // compact rolled loop on purpose: these few hundred bytes of code run 8 times per group out of the instruction
    // cache; the fully unrolled variant was fetch-bound and slower (measured on B200)
#pragma unroll 1
    for (int pp = 0; pp < 8; ++pp) {
        const int p = 8 * J + pp, buf = pp & 1, pty = p & 15;
        if (ty == pty) {            // part of row p lives in register row I
            if (tx == pp) pl.piv[buf] = a[I][J];
            double2* dst = reinterpret_cast<double2*>(&pl.row[buf][tx * 8]);
            const double vJ = (tx > pp) ? a[I][J] : 0.0;
            if (J & 1) {
                dst[J / 2] = make_double2(0.0, vJ);
            } else {
                dst[J / 2] = make_double2(vJ, a[I][(J + 1) & 7]);
            }
#pragma unroll
            for (int m = J / 2 + 1; m < 4; ++m) dst[m] = make_double2(a[I][2 * m], a[I][2 * m + 1]);
        }
        if (tx == pp) {             // part of column p lives in register column J
            double2* dst = reinterpret_cast<double2*>(&pl.col[buf][ty * 4]);
            const double vI = (ty > pty) ? a[I][J] : 0.0;
            if (I & 1) {
                dst[I / 2] = make_double2(0.0, vI);
            } else {
                dst[I / 2] = make_double2(vI, a[(I + 1) & 3][J]);
            }
            if (I / 2 == 0) dst[1] = make_double2(a[2][J], a[3][J]);
        }
        panel_sync<NT>();
        double r8[8], c4[4];
        {
            const double2* rr = reinterpret_cast<const double2*>(&pl.row[buf][tx * 8]);
            const double2* cc = reinterpret_cast<const double2*>(&pl.col[buf][ty * 4]);
            const double piv = pl.piv[buf];
#pragma unroll
            for (int m = J / 2; m < 4; ++m) {
                const double2 v = rr[m];
                r8[2 * m] = v.x;
                r8[2 * m + 1] = v.y;
            }
            double2 cv[2];
#pragma unroll
            for (int m = I / 2; m < 2; ++m) cv[m] = cc[m];
            const double rp = fast_rcp(piv);
            if (ty == pty && tx == pp) pl.rcp[p] = rp;
#pragma unroll
            for (int m = I / 2; m < 2; ++m) {
                c4[2 * m] = cv[m].x * rp;
                c4[2 * m + 1] = cv[m].y * rp;
            }
        }
        // column J of lanes tx == pp keeps the UNSCALED multipliers (their pivot-row entry is masked to zero), rows <= p
        // keep U: both are finished off after the last pivot
#pragma unroll
        for (int i = I; i < 4; ++i)
#pragma unroll
            for (int j = J; j < 8; ++j) a[i][j] = fma(-c4[i], r8[j], a[i][j]);
    }
}

template <int NT = LU_THREADS>
__device__ __forceinline__ void factor_block(double* __restrict__ sC, PivotLine& pl) {
    const int t = threadIdx.x, ty = t >> 3, tx = t & 7;
    double a[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) a[i][j] = sC[(ty + 16 * i) * TS + tx + 8 * j];
    factor_group<0, NT>(a, pl, ty, tx);
    factor_group<1, NT>(a, pl, ty, tx);
    factor_group<2, NT>(a, pl, ty, tx);
    factor_group<3, NT>(a, pl, ty, tx);
    factor_group<4, NT>(a, pl, ty, tx);
    factor_group<5, NT>(a, pl, ty, tx);
    factor_group<6, NT>(a, pl, ty, tx);
    factor_group<7, NT>(a, pl, ty, tx);
    panel_sync<NT>();      // pl.rcp complete
    // registers hold U on and above the diagonal and the unscaled multipliers below it: L(r, c) = a(r, c) / U(c, c)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const double rc = pl.rcp[tx + 8 * j];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = ty + 16 * i, col = tx + 8 * j;
            sC[row * TS + col] = (row > col) ? a[i][j] * rc : a[i][j];
        }
    }
    panel_sync<NT>();
}

// ---- X W = C for an upper-triangular W, in place on a 64 x 64 tile held in the same 2-D cyclic register layout -----
// W is given PRE-SCALED by its reciprocal diagonal, W'(p, c) = W(p, c) / W(p, p) for c > p, at sW[p * wr + c * wc],
// and dinv = 1 / diag W.  With a_p the current (unscaled) pivot column of the tile, x_p = a_p dinv_p and
// a_c -= a_p W'(p, c): the only dependent work between two pivots is one shuffle and one DFMA.  Column p of the tile
// lives in the 8 lanes that share ty, so the multipliers travel by warp shuffle: no barrier in the whole solve.  Serves
//   L21 = C U11^-1              (sW = published image, upper part = U11 rows scaled; wr = TS, wc = 1)
//   U12^T = C^T (L11^T)^-1      (W = L11^T, unit diagonal: W(p, c) = L11(c, p) = image[c][p]; wr = 1, wc = TS; dinv = 1;
//                                the tile is loaded / stored transposed)
//   inv(U11) = I U11^-1         (tile = identity), used by the back-substitution chain.
// RI = register rows per thread: 4 for a whole tile, 2 when two CTAs share a panel tile (each solves 32 of the 64 independent rows)
template <int J, bool TR, int RI>
__device__ __forceinline__ void panel_group(double (&a)[RI][8], const double* __restrict__ sW, const double* __restrict__ dinv,
                                            int tx, int lane) {
    constexpr int wr = TR ? 1 : TS, wc = TR ? TS : 1;
#pragma unroll 1
    for (int pp = 0; pp < 8; ++pp) {
        const int p = 8 * J + pp;
        const int src = (lane & ~7) | pp;
        const double* wrow = sW + p * wr + tx * wc;
        double w[8];
        w[J] = (tx > pp) ? wrow[8 * J * wc] : 0.0;
#pragma unroll
        for (int j = J + 1; j < 8; ++j) w[j] = wrow[8 * j * wc];
        const double rp = TR ? 1.0 : dinv[p];
        double ap[RI];
#pragma unroll
        for (int i = 0; i < RI; ++i) ap[i] = __shfl_sync(0xffffffffu, a[i][J], src);
#pragma unroll
        for (int i = 0; i < RI; ++i) a[i][J] = (tx == pp) ? ap[i] * rp : fma(-ap[i], w[J], a[i][J]);
#pragma unroll
        for (int j = J + 1; j < 8; ++j)
#pragma unroll
            for (int i = 0; i < RI; ++i) a[i][j] = fma(-ap[i], w[j], a[i][j]);
    }
}

template <bool TR, int RI = 4>
__device__ __forceinline__ void panel_solve(double (&a)[RI][8], const double* __restrict__ sW, const double* __restrict__ dinv) {
    const int tx = threadIdx.x & 7, lane = threadIdx.x & 31;
    panel_group<0, TR, RI>(a, sW, dinv, tx, lane);
    panel_group<1, TR, RI>(a, sW, dinv, tx, lane);
    panel_group<2, TR, RI>(a, sW, dinv, tx, lane);
    panel_group<3, TR, RI>(a, sW, dinv, tx, lane);
    panel_group<4, TR, RI>(a, sW, dinv, tx, lane);
    panel_group<5, TR, RI>(a, sW, dinv, tx, lane);
    panel_group<6, TR, RI>(a, sW, dinv, tx, lane);
    panel_group<7, TR, RI>(a, sW, dinv, tx, lane);
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

// Growth guard of the unpivoted factorisation, run by the diag CTA off the critical path (after the block is published):
// scalars[1] / scalars[2] carry the smallest / largest pivot magnitude of the factorisation so far; a zero or NaN pivot raises
// LLS_STATUS_NAN (LAPACK: "singular matrix"), and after the last diagonal block a smallest pivot below PIVOT_RATIO times the
// largest raises LLS_STATUS_SMALL_PIVOT: the matrix is not the diagonally dominant one the solver relies on (the lifting-line
// systems have pivot ratios of O(10)) and partial pivoting would have started to matter.
constexpr double PIVOT_RATIO = 1e-6;
template <int NT>
__device__ __forceinline__ void pivot_guard(const double* __restrict__ sC, PivotLine& pl, int blk, int nblk, double* __restrict__ scalars,
                                            int* __restrict__ status) {
    const int t = threadIdx.x;
    if (t < NB) {
        const double d = fabs(sC[t * TS + t]);
        // the identity padding beyond n contributes pivots of exactly 1: they say nothing about the matrix and are left out
        double m = (d == 1.0) ? 1.0e300 : d, M = (d == 1.0) ? 0.0 : d;
        bool bad = !(d > 0.0) || !(d < 1.0e300);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            m = fmin(m, __shfl_xor_sync(0xffffffffu, m, o));
            M = fmax(M, __shfl_xor_sync(0xffffffffu, M, o));
        }
        bad = __any_sync(0xffffffffu, bad);
        if ((t & 31) == 0) {
            if (bad) atomicOr(status, LLS_STATUS_NAN);
            pl.piv[t >> 5] = m;
            pl.rcp[t >> 5] = M;        // (the reciprocal pivots were consumed before the block was published)
        }
    }
    panel_sync<NT>();
    if (t == 0) {
        double m = fmin(pl.piv[0], pl.piv[1]), M = fmax(pl.rcp[0], pl.rcp[1]);
        if (blk > 0) {
            m = fmin(m, scalars[1]);
            M = fmax(M, scalars[2]);
        }
        scalars[1] = m;
        scalars[2] = M;
        if (blk == nblk - 1 && m < PIVOT_RATIO * M) atomicOr(status, LLS_STATUS_SMALL_PIVOT);
    }
}

// ---- keeping the serial chain's SM to itself ---------------------------------------------------------------------------
// The in-register factorisation of the diagonal block is latency- and FP64-issue-bound on ONE CTA; trailing-update CTAs that
// share its SM keep the FP64 pipe busy with DMMA and stretch it 2-3.5x (measured per block step with scripts/lu_clocks.py:
// 22.8 k cycles alone, 40-80 k beside update CTAs), which then is the length of every launch whose trailing update is shorter.
// So the diag CTA announces the SM it runs on, and an update CTA that finds itself on that SM parks (after its operand loads
// are in flight, before it touches the tensor pipe) until the diagonal block is published.  At most two CTA slots idle for
// the length of the factorisation; the rest of the machine is unaffected.
struct ChainGuard {
    int* diag_sm;      // word the diag CTA writes its tag to
    const int* flag;   // publish flag of the diagonal block being factored in this launch
    int epoch;
    int tag;           // (epoch, block, SM id) of this CTA: equals *diag_sm iff the diag CTA of this launch runs on this SM
};
__device__ __forceinline__ int chain_tag(int epoch, int blk) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    return ((epoch & 0x1fff) << 18) | ((blk & 0x3ff) << 8) | (int)(smid & 0xff);
}
__device__ __forceinline__ void chain_announce(const ChainGuard& g) {     // diag CTA, one thread
    asm volatile("st.relaxed.gpu.global.s32 [%0], %1;\n" ::"l"(g.diag_sm), "r"(g.tag) : "memory");
}
__device__ __forceinline__ void chain_yield(const ChainGuard& g) {        // update CTA, one thread
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(g.diag_sm) : "memory");
    if (v == g.tag) {
        while (ld_acquire(g.flag) != g.epoch) __nanosleep(128);
    }
}

// ---- persistent trailing-update workers ----------------------------------------------------------------------------
// The tiles of the trailing matrix that carry no panel role are walked in serpentine order (left to right on even tile
// rows, right to left on odd ones), so two consecutive tiles always share the L tile or the U tile and exactly ONE 32 KB
// operand has to arrive per tile.  A worker CTA owns three shared-memory slots (L, U, one being filled by cp.async for
// the next tile) and keeps the next C tile in registers, so the loads of tile t+1 fly while tile t is on the tensor
// pipe: the DMMA phases of a CTA run back to back instead of load -> mma -> store once per CTA.  Tiles are handed out
// in chunks of consecutive serpentine indices through one atomic counter per block step; a worker grabs one chunk
// ahead, so the counter's latency is hidden as well.  Panel-role CTAs join the same loop once their panel is done.
struct UpdateGeom {
    const double* L;   size_t l_stride;    // L tile of tile row r: L + r * l_stride, leading dimension ld
    const double* U;   size_t u_stride;    // U tile of tile column c: U + c * u_stride, leading dimension ldu
    size_t ldu;
    double* C;                             // tile (r, c): C + (r * ld + c) * NB, leading dimension ld
    size_t ld;
    int rows, cols;
};

__device__ __forceinline__ void serpentine(int o, int cols, int& r, int& c) {
    r = o / cols;
    const int cc = o - r * cols;
    c = (r & 1) ? (cols - 1 - cc) : cc;
}

template <int NT>
__device__ __forceinline__ void c_prefetch(double2 (&cn)[4][WarpTile<NT>::NI], const double* __restrict__ g, size_t ldg) {
    const int lane = threadIdx.x & 31;
    const int wm = WarpTile<NT>::wm(), wn = WarpTile<NT>::wn();
    const int gq = lane >> 2, q = lane & 3;
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < WarpTile<NT>::NI; ++ni)
            cn[mi][ni] = *reinterpret_cast<const double2*>(g + (size_t)(wm + mi * 8 + gq) * ldg + wn + ni * 8 + 2 * q);
}

struct WorkerShared {
    int first;
    int ahead[2];
};

template <int NT>
__device__ __forceinline__ void lu_update_worker(const UpdateGeom& G, int* __restrict__ counter, int chunk, double* __restrict__ sm,
                                                 WorkerShared& ws, const ChainGuard& guard) {
    const int t = threadIdx.x;
    const int T = G.rows * G.cols;
    const int dbg = chunk >> 8;      // timing ablations (LLS_LU_WORKER_DBG): 1 no C stores, 2 no C loads, 4 no operand loads
    chunk &= 255;
    if (t == 0) {
        ws.first = atomicAdd(counter, chunk);
    }
    __syncthreads();
    int o = ws.first;
    if (o >= T) return;
    int oe = min(o + chunk, T);
    int par = 0;                      // parity of the chunk being worked on: start of the next chunk is ws.ahead[par]
    int sl = 0, su = 1, sf = 2;       // slots holding L, U, and the free one
    int r, c;
    serpentine(o, G.cols, r, c);
    load_tile_async<NT>(sm + sl * (NB * TS), G.L + (size_t)r * G.l_stride, G.ld);
    load_tile_async<NT>(sm + su * (NB * TS), G.U + (size_t)c * G.u_stride, G.ldu);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    constexpr int NI = WarpTile<NT>::NI;
    double2 cn[4][NI];
    c_prefetch<NT>(cn, G.C + ((size_t)r * G.ld + c) * NB, G.ld);
    bool chunk_start = true;
    bool first_tile = true;
    while (true) {
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        if (first_tile && t == 0 && guard.diag_sm != nullptr) chain_yield(guard);
        first_tile = false;
        __syncthreads();              // operands of this tile visible; every warp is done with the previous tile's slots
        // successor of this tile: next in the chunk, or the first tile of the chunk grabbed ahead
        int no = o + 1;
        if (no >= oe) no = chunk_start ? T : ws.ahead[par];       // (a one-tile chunk has nothing grabbed ahead yet: see below)
        int pend = 0;
        if (chunk_start && t == 0) pend = atomicAdd(counter, chunk);
        if (chunk_start && o + 1 >= oe) {
            // chunk of a single tile: the grab cannot hide behind the tile; take it synchronously
            if (t == 0) ws.ahead[par] = pend;
            __syncthreads();
            no = ws.ahead[par];
        }
        const bool has_next = no < T;
        int nr = r, nc = c;
        bool both = false, newL = false;
        if (has_next) {
            serpentine(no, G.cols, nr, nc);
            newL = (nr != r);
            both = newL && (nc != c);
            if (!(dbg & 4)) {
                if (newL) load_tile_async<NT>(sm + sf * (NB * TS), G.L + (size_t)nr * G.l_stride, G.ld);
                else load_tile_async<NT>(sm + sf * (NB * TS), G.U + (size_t)nc * G.u_stride, G.ldu);
            }
            asm volatile("cp.async.commit_group;\n" ::: "memory");
        }
        double acc[4][NI][2];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < NI; ++ni) {
                acc[mi][ni][0] = -cn[mi][ni].x;
                acc[mi][ni][1] = -cn[mi][ni].y;
            }
        if (has_next && !(dbg & 2)) c_prefetch<NT>(cn, G.C + ((size_t)nr * G.ld + nc) * NB, G.ld);
        mma_tile<NT>(sm + sl * (NB * TS), sm + su * (NB * TS), acc);
        if (!(dbg & 1)) acc_io<false, NT>(acc, G.C + ((size_t)r * G.ld + c) * NB, G.ld);
        if (chunk_start && o + 1 < oe && t == 0) ws.ahead[par] = pend;   // read after the next tile's barrier at the earliest
        if (!has_next) break;
        if (newL) { const int x = sl; sl = sf; sf = x; } else { const int x = su; su = sf; sf = x; }
        if (both) {
            // chunk boundary that changes both operands: the old U slot is free only now
            __syncthreads();
            if (!(dbg & 4)) load_tile_async<NT>(sm + su * (NB * TS), G.U + (size_t)nc * G.u_stride, G.ldu);
            asm volatile("cp.async.commit_group;\n" ::: "memory");
        }
        chunk_start = (no >= oe);
        if (chunk_start) {
            oe = min(no + chunk, T);
            par ^= 1;
        }
        o = no;
        r = nr;
        c = nc;
    }
}

// ---- one block step ------------------------------------------------------------------------------------
// FIRST: no trailing update (factor block 0 and its panel); otherwise update with panel k and factor panel k+1.
// Grid (1-D): block 0 = diag tile, 1..rem-1 = L tiles, rem..2rem-2 = U tiles, then the (rem-1)^2 others,
// where rem = number of block rows/cols of the trailing matrix handled here.
// PASSES = 2: the trailing tiles take the updates of panels k AND k+1 in one visit (rank-128 update, half the traffic on C
// and twice the tensor work per tile visit); the tiles of block row / column k+1 were brought up to date, and panel k+1
// factored, by a thin PASSES = 1 launch just before.  The look-ahead roles then factor panel k + PASSES.
// WORKERS: the tiles without a panel role are not mapped to CTAs; blocks >= 2 rem - 1 are persistent workers (and the panel
// CTAs turn into workers when their panel is done), see lu_update_worker.  `counters` holds one work counter per block step,
// zeroed by the FIRST launch of every factorisation.
template <bool FIRST, int PASSES, bool WORKERS, int SPLIT = 1>
__global__ void __launch_bounds__(WORKERS ? 256 : LU_THREADS, WORKERS ? 2 : 3) lu_step_kernel(double* __restrict__ M, int ld, int k, int rem,
                                                                double* __restrict__ rhs, double* __restrict__ dbuf,
                                                                double* __restrict__ uinv, int* __restrict__ flags, int epoch,
                                                                double* __restrict__ scalars, int* __restrict__ status, CaseStrides cs,
                                                                int* __restrict__ counters, int chunk, int nworkers, int* __restrict__ diag_sm,
                                                                int* __restrict__ claims, int panel_sms, int inv_block) {
    constexpr int NT = WORKERS ? 256 : LU_THREADS;   // threads per CTA; the panel roles always run on threads 0..127
    static_assert(!(WORKERS && (FIRST || PASSES != 1)), "workers take one panel per visit");
    static_assert(SPLIT == 1 || SPLIT == 2 || SPLIT == 4, "CTAs per panel tile");
    static_assert(!(SPLIT > 1 && (WORKERS || PASSES != 1)), "split panel tiles belong to the chain-bound launches");
    // programmatic dependent launch: the CTAs of block step k+1 are placed on the SMs while step k drains and start the moment
    // its memory is visible, instead of paying the launch latency between two launches of the chain (no-ops in a plain launch)
    asm volatile("griddepcontrol.wait;\n" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
    if (LLS_CASE_INACTIVE(cs)) return;
    M = case_ptr(M, cs.M);
    if (rhs != nullptr) rhs = case_work_ptr(rhs, cs.work);
    dbuf = case_work_ptr(dbuf, cs.work);
    uinv = case_work_ptr(uinv, cs.work);
    flags = case_work_ptr(flags, cs.work);
    scalars = case_work_ptr(scalars, cs.work);
    status = case_work_ptr(status, cs.work);
    extern __shared__ __align__(16) double sm[];
    double* sA = sm;
    double* sB = sm + NB * TS;
    __shared__ __align__(16) PivotLine pl;
    __shared__ __align__(16) double dinv[NB];
    __shared__ WorkerShared wsh;
    const int t = threadIdx.x;
    const int b = blockIdx.x;
    const int base = FIRST ? 0 : k + PASSES;  // first block row/col of the trailing matrix handled here
    if (FIRST && b == 0) {      // per-factorisation counters: worker chunks, role claims (two per block step)
        if (counters != nullptr) {
            counters = case_work_ptr(counters, cs.work);
            for (int e = t; e <= ld / NB; e += LU_THREADS) counters[e] = 0;
        }
        if (claims != nullptr) {
            int* c = case_work_ptr(claims, cs.work);
            for (int e = t; e < 2 * (ld / NB + 1); e += LU_THREADS) c[e] = 0;
        }
    }
    ChainGuard guard;
    guard.diag_sm = (diag_sm != nullptr) ? case_work_ptr(diag_sm, cs.work) : nullptr;
    guard.flag = flags + base;
    guard.epoch = epoch;
    guard.tag = chain_tag(epoch, base);
    if (b == 0 && t == 0 && guard.diag_sm != nullptr) chain_announce(guard);
    UpdateGeom geom;
    if (WORKERS) {
        const size_t o = (size_t)(base + 1) * NB;
        geom.L = M + o * ld + (size_t)k * NB;
        geom.l_stride = (size_t)NB * ld;
        geom.U = M + (size_t)k * NB * ld + o;
        geom.u_stride = NB;
        geom.ldu = ld;
        geom.C = M + o * ld + o;
        geom.ld = ld;
        geom.rows = geom.cols = rem - 1;
        // grid order: diag CTA, `nworkers` workers, then the panel tiles.  The panel CTAs that do not fit beside the workers start
        // when slots free up, find the diagonal block published long ago and never spin; the few that start at once do the waiting
        if (b >= 1 && b <= nworkers) {
            lu_update_worker<NT>(geom, counters + base, chunk, sm, wsh, guard);
            return;
        }
    }
    // index among the roles: 0 diag, 1 .. 2 rem - 2 the L and U panel tiles, then the other tiles
    int pb = (WORKERS && b > 0) ? b - nworkers : b;     // (WORKERS: the panel roles follow the workers in the grid)
    if (!WORKERS && !FIRST && claims != nullptr && b > 0) {
        // Roles by SM instead of by block index: CTAs that land on the first `panel_sms` SMs take the panel tiles (three to an SM),
        // everybody else the update tiles; each side falls back to the other queue when its own is empty, and as the grid has
        // exactly one CTA per tile every tile is taken.  The latency-bound panel solves then share their SMs with each other (they
        // spin or run dependent FP64 chains) instead of with DMMA streams that stretch them 2-3x, and the update tiles get SMs
        // whose tensor pipe they do not have to share with the chain.
        __shared__ int s_role;
        if (t == 0) {
            claims = case_work_ptr(claims, cs.work) + 2 * base;
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            const int P = 2 * rem - 2, T = (rem - 1) * (rem - 1);
            int idx;
            if ((int)smid < panel_sms) {
                const int p = atomicAdd(claims, 1);
                idx = (p < P) ? 1 + p : 1 + P + atomicAdd(claims + 1, 1);
            } else {
                const int o = atomicAdd(claims + 1, 1);
                idx = (o < T) ? 1 + P + o : 1 + atomicAdd(claims, 1);
            }
            s_role = idx;
        }
        __syncthreads();
        pb = s_role;
    }
    int bi, bj;
    // INV (block `inv_block` = 2 rem - 1, right behind the panel tiles, if there is one): inverts U11 for the back-substitution chain.  It waits for
    // the diagonal block like a panel CTA and runs beside them; with the inversion inside the diag CTA that CTA was the last
    // of a chain-bound launch to finish (by 5-16 k cycles for trailing matrices of 24-46 block rows).
    enum { DIAG, LCOL, UROW, OTHER, INV } role;
    // SPLIT > 1: that many CTAs per panel tile (and for the inverse).  All update the whole tile (the tensor work is not on anybody's
    // critical path here), then each solves 64 / SPLIT of its 64 independent rows (L tiles, inverse) or columns (U tiles): that
    // fraction of the FMAs and shuffles per pivot of the issue-bound substitution.  Used while launches are chain-bound and SMs
    // are idle anyway.
    int half = 0;     // which part of the tile this CTA solves
    const int npanel = 2 * SPLIT * (rem - 1);
    const int ninv = (inv_block >= 0) ? SPLIT : 0;
    if (inv_block >= 0 && pb >= inv_block && pb < inv_block + ninv) {
        role = INV; bi = base; bj = base; half = pb - inv_block;
    } else if (pb == 0) {
        role = DIAG; bi = base; bj = base;
    } else if (pb <= npanel / 2) {
        half = (pb - 1) % SPLIT;
        role = LCOL; bi = base + 1 + (pb - 1) / SPLIT; bj = base;
    } else if (pb <= npanel) {
        half = (pb - 1 - npanel / 2) % SPLIT;
        role = UROW; bi = base; bj = base + 1 + (pb - 1 - npanel / 2) / SPLIT;
    } else {
        role = OTHER;
        const int o = pb - (npanel + 1) - ninv;
        bi = base + 1 + o / (rem - 1);
        bj = base + 1 + o % (rem - 1);
    }
    double* gC = M + (size_t)bi * NB * ld + (size_t)bj * NB;
    LU_CLK(0);
    if (role == INV) {
        // no tile of its own
    } else if (!FIRST) {
        double acc[4][WarpTile<NT>::NI][2];
        {
            double2 cn[4][WarpTile<NT>::NI];
            c_fetch<NT>(cn, gC, ld);                                                  // C, in flight together with ...
            load_tile_async<NT>(sA, M + (size_t)bi * NB * ld + (size_t)k * NB, ld);   // L[bi,k]
            load_tile_async<NT>(sB, M + (size_t)k * NB * ld + (size_t)bj * NB, ld);   // U[k,bj]
            async_wait_all();
            if (role == OTHER && t == 0 && guard.diag_sm != nullptr) chain_yield(guard);
            c_negate<NT>(acc, cn);                                                    // -C
        }
        __syncthreads();
        LU_CLK(1);
        mma_tile<NT>(sA, sB, acc);                                                // -C + L U
        if (PASSES == 2) {
            __syncthreads();
            load_tile_async<NT>(sA, M + (size_t)bi * NB * ld + (size_t)(k + 1) * NB, ld);   // L[bi,k+1]
            load_tile_async<NT>(sB, M + (size_t)(k + 1) * NB * ld + (size_t)bj * NB, ld);   // U[k+1,bj]
            async_wait_all();
            __syncthreads();
            mma_tile<NT>(sA, sB, acc);
        }
        LU_CLK(2);
        if (role == OTHER) {
            acc_io<false, NT>(acc, gC, ld);                                       // C - L U
            return;
        }
        if (rhs != nullptr && bj == base && half == 0 && t < NB) {   // b[bi] -= L[bi,kk] y_kk for the last panel applied (the thin launch did the other)
            double s = 0.0;
            const double* yk = rhs + (size_t)(k + PASSES - 1) * NB;
#pragma unroll 8
            for (int q = 0; q < NB; ++q) s = fma(sA[t * TS + q], yk[q], s);
            rhs[(size_t)bi * NB + t] -= s;
        }
        __syncthreads();
        acc_io<false, NT>(acc, sB, TS);   // C - L U, row-major in shared memory
        __syncthreads();
    } else {
        load_tile_async<NT>(sB, gC, ld);
        async_wait_all();
        __syncthreads();
    }
    if (NT == LU_THREADS || t < LU_THREADS) {      // ---- panel roles: threads 0..127, barrier panel_sync<NT>() ----
    double* sC = sB;
    const int blk = base;                       // index of the diagonal block being factored
    // published image of the factored diagonal block (per parity), already in the padded shared-memory layout so that
    // consumers copy it flat: row-major LU with the strictly upper part scaled by 1 / diag(U), then 1 / diag(U)
    double* img = dbuf + (size_t)(blk & 1) * PUB_DOUBLES;
    constexpr int IMG2 = NB * TS / 2;           // double2 elements of the image

    if (role == DIAG) {
        LU_CLK(3);
        factor_block<NT>(sC, pl);
        LU_CLK(4);
        if (t < NB) dinv[t] = 1.0 / sC[t * TS + t];
        panel_sync<NT>();
        for (int e = t; e < NB * NB; e += LU_THREADS) {
            const int r = e >> 6, c = e & 63;
            const double v = sC[r * TS + c];
            sA[r * TS + c] = (c > r) ? v * dinv[r] : v;
        }
        panel_sync<NT>();
        {
            double2* g0 = reinterpret_cast<double2*>(img);
            const double2* s0 = reinterpret_cast<const double2*>(sA);
#pragma unroll
            for (int e = t; e < IMG2; e += LU_THREADS) g0[e] = s0[e];
            if (t < NB) img[NB * TS + t] = dinv[t];
        }
        __threadfence();
        panel_sync<NT>();
        if (t == 0) st_release(flags + blk, epoch);
        LU_CLK(5);
        // everything below is off the critical path of the panel CTAs
        store_tile(gC, ld, sC);
        pivot_guard<NT>(sC, pl, blk, ld / NB, scalars, status);
        if (rhs != nullptr && t < 32) forward_rhs_warp(sC, rhs + (size_t)blk * NB);
        // W = U11, pre-scaled rows in sA: the solve below turns the identity into inv(U11) (unless an INV CTA does that)
    } else {
        // panel roles: wait for the factored diagonal block
        if (t == 0) {
            while (ld_acquire(flags + blk) != epoch) __nanosleep(32);
        }
        panel_sync<NT>();
        const double2* g = reinterpret_cast<const double2*>(img);
        double2* sT2 = reinterpret_cast<double2*>(sA);
        double2 v[IMG2 / LU_THREADS];
#pragma unroll
        for (int u = 0; u < IMG2 / LU_THREADS; ++u) v[u] = __ldcg(g + t + u * LU_THREADS);
#pragma unroll
        for (int u = 0; u < IMG2 / LU_THREADS; ++u) sT2[t + u * LU_THREADS] = v[u];
        if (t < NB) dinv[t] = (role == UROW) ? 1.0 : __ldcg(img + NB * TS + t);
        // LCOL: W = U11 (scaled upper part, row-major); UROW (transposed problem): W = L11^T, W(p, c) = image[c][p]
        panel_sync<NT>();
    }
    const bool solves = !(role == DIAG && inv_block >= 0);
    const bool whole = SPLIT == 1 || role == DIAG;
    constexpr int RI = 4 / SPLIT;            // register rows per thread of a part
    constexpr int PR = 16 * RI;              // rows (columns) of a part
    const int ty = t >> 3, tx = t & 7;
    // 2-D cyclic register tile: a[i][j] <-> (row ty + 16 i, col tx + 8 j) of X; UROW works on the transposed tile
    const int sr = (role == UROW) ? 1 : TS, sc = (role == UROW) ? TS : 1;
    if (solves && whole) {
        double a[4][8];
        if (role == DIAG || role == INV) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) a[i][j] = (ty + 16 * i == tx + 8 * j) ? 1.0 : 0.0;
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) a[i][j] = sC[(ty + 16 * i) * sr + (tx + 8 * j) * sc];
        }
        LU_CLK(6);
        if (role == UROW) panel_solve<true>(a, sA, dinv);
        else panel_solve<false>(a, sA, dinv);
        LU_CLK(7);
        panel_sync<NT>();   // everyone is done reading the tile (and, in the diag CTA, the factors for the right-hand side)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) sC[(ty + 16 * i) * sr + (tx + 8 * j) * sc] = a[i][j];
    } else if (solves) {
        // this CTA's part: rows (L tile, inverse) / columns (U tile) PR half .. PR half + PR - 1, i.e. register rows RI half ..
        double a[RI][8];
        const int r0 = ty + PR * half;
#pragma unroll
        for (int i = 0; i < RI; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j)
                a[i][j] = (role == INV) ? ((r0 + 16 * i == tx + 8 * j) ? 1.0 : 0.0) : sC[(r0 + 16 * i) * sr + (tx + 8 * j) * sc];
        LU_CLK(6);
        if (role == UROW) panel_solve<true, RI>(a, sA, dinv);
        else panel_solve<false, RI>(a, sA, dinv);
        LU_CLK(7);
        panel_sync<NT>();
#pragma unroll
        for (int i = 0; i < RI; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) sC[(r0 + 16 * i) * sr + (tx + 8 * j) * sc] = a[i][j];
    }
    panel_sync<NT>();
    if (solves) {
        double* dst = (role == DIAG || role == INV) ? uinv + (size_t)blk * NB * NB : gC;
        const size_t ldd = (role == DIAG || role == INV) ? (size_t)NB : (size_t)ld;
        if (whole) {
            store_tile(dst, ldd, sC);
        } else if (role != UROW) {      // rows PR half .. +PR-1, all 64 columns
            const int c2 = t & 31, rr = t >> 5;
#pragma unroll 4
            for (int r = PR * half + rr; r < PR * half + PR; r += 4)
                *reinterpret_cast<double2*>(dst + (size_t)r * ldd + 2 * c2) = *reinterpret_cast<const double2*>(sC + r * TS + 2 * c2);
        } else {                        // all 64 rows, columns PR half .. +PR-1 (PR / 2 double2 per row)
            constexpr int W2 = PR / 2;
            const int c2 = t % W2, rr = t / W2;
#pragma unroll 4
            for (int r = rr; r < NB; r += LU_THREADS / W2)
                *reinterpret_cast<double2*>(dst + (size_t)r * ldd + PR * half + 2 * c2) = *reinterpret_cast<const double2*>(sC + r * TS + PR * half + 2 * c2);
        }
    }
    LU_CLK(8);
    }   // panel roles
    if (WORKERS && rem > 1) {
        __syncthreads();      // the tile has left shared memory: the slots are the worker's now
        ChainGuard none = guard;
        none.diag_sm = nullptr;       // the panel of this launch is done
        lu_update_worker<NT>(geom, counters + base, chunk, sm, wsh, none);
    }
}

// ---- block back-substitution as ONE launch: a chain of CTAs, one per block row --------------------------------
// CTA of block row i folds x_k (k = last .. i+1) into y_i as they are published, then x_i = inv(U_ii) y_i with the
// inverse diagonal block the factorisation left in uinv, and publishes x_i.  Blocks are numbered from the LAST block
// row, so every producer is scheduled before its consumers.  Thread t owns row t of the tile; the tile row and the
// inverse row are fetched into registers before the wait, so the chain step is: flag -> 64-double x_k -> two
// 64-term dot products -> release.
constexpr int BS_THREADS = NB;

__global__ void __launch_bounds__(BS_THREADS) lu_backsolve_chain_kernel(const double* __restrict__ M, int ld, int nblk,
                                                                        const double* __restrict__ uinv,
                                                                        const double* __restrict__ y, double* __restrict__ x,
                                                                        int* __restrict__ xflags, int epoch, CaseStrides cs) {
    if (LLS_CASE_INACTIVE(cs)) return;
    M = case_ptr(M, cs.M);
    uinv = case_work_ptr(uinv, cs.work);
    y = case_work_ptr(y, cs.work);
    x = case_work_ptr(x, cs.work);
    xflags = case_work_ptr(xflags, cs.work);
    __shared__ __align__(16) double sx[NB];
    const int t = threadIdx.x;
    const int bi = nblk - 1 - blockIdx.x;
    double yt = y[(size_t)bi * NB + t];
    double2 ui[NB / 2];
    {
        const double2* g = reinterpret_cast<const double2*>(uinv + (size_t)bi * NB * NB + (size_t)t * NB);
#pragma unroll
        for (int c = 0; c < NB / 2; ++c) ui[c] = g[c];
    }
    for (int k = nblk - 1; k > bi; --k) {
        double2 u[NB / 2];
        {
            const double2* g = reinterpret_cast<const double2*>(M + ((size_t)bi * NB + t) * ld + (size_t)k * NB);
#pragma unroll
            for (int c = 0; c < NB / 2; ++c) u[c] = g[c];
        }
        if (t == 0) {
            while (ld_acquire(xflags + k) != epoch) __nanosleep(20);
        }
        __syncthreads();
        sx[t] = __ldcg(x + (size_t)k * NB + t);
        __syncthreads();
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        const double2* sx2 = reinterpret_cast<const double2*>(sx);
#pragma unroll
        for (int c = 0; c < NB / 2; c += 2) {
            const double2 a = sx2[c], b = sx2[c + 1];
            s0 = fma(u[c].x, a.x, s0);
            s1 = fma(u[c].y, a.y, s1);
            s2 = fma(u[c + 1].x, b.x, s2);
            s3 = fma(u[c + 1].y, b.y, s3);
        }
        yt -= (s0 + s1) + (s2 + s3);
        __syncthreads();
    }
    sx[t] = yt;
    __syncthreads();
    {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        const double2* sx2 = reinterpret_cast<const double2*>(sx);
#pragma unroll
        for (int c = 0; c < NB / 2; c += 2) {
            const double2 a = sx2[c], b = sx2[c + 1];
            s0 = fma(ui[c].x, a.x, s0);
            s1 = fma(ui[c].y, a.y, s1);
            s2 = fma(ui[c + 1].x, b.x, s2);
            s3 = fma(ui[c + 1].y, b.y, s3);
        }
        x[(size_t)bi * NB + t] = (s0 + s1) + (s2 + s3);
    }
    __threadfence();
    __syncthreads();
    if (t == 0) st_release(xflags + bi, epoch);
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

// =====================================================================================================================
// Row-partitioned LU for one large scene over the GPUs of a box.  Block rows are dealt
// round-robin (owner of block row k = k mod P), so the trailing update stays balanced.  Per block step k:
//   owner(k):  lu_dist_owner_kernel  - factors the (already updated) diagonal tile, solves the U block row
//                                      U[k, k+1:] = L11^-1 A[k, k+1:], y_k = L11^-1 b_k, and packs
//                                      [diag image | y_k | U tiles] into one contiguous send buffer
//   host:      NCCL broadcast of that buffer (torch.distributed), root = owner(k)
//   all ranks: lu_dist_panel_kernel  - L[i,k] = A[i,k] U11^-1 and b_i -= L[i,k] y_k for the local block rows i > k
//              lu_dist_update_kernel - A[i,j] -= L[i,k] U[k,j] for the local block rows i > k, all j > k (DMMA)
// Back substitution walks the block rows backwards; the owner of block row k forms x_k from its row of U and the replicated
// x_{>k} (lu_dist_back_kernel) and the 64 new values are broadcast.
// =====================================================================================================================
constexpr int DIST_HDR = PUB_DOUBLES + NB;      // doubles before the packed U tiles in the step buffer

__global__ void __launch_bounds__(LU_THREADS, 3) lu_dist_owner_kernel(double* __restrict__ Mloc, int ld, int k, int lbk, int ncols,
                                                                      double* __restrict__ rhs, double* __restrict__ send,
                                                                      double* __restrict__ uinv, int* __restrict__ flags, int epoch,
                                                                      double* __restrict__ scalars, int* __restrict__ status) {
    extern __shared__ __align__(16) double sm[];
    double* sA = sm;
    double* sC = sm + NB * TS;
    __shared__ __align__(16) PivotLine pl;
    __shared__ __align__(16) double dinv[NB];
    const int t = threadIdx.x, b = blockIdx.x;
    const bool diag = (b == 0);
    double* gC = Mloc + (size_t)lbk * NB * ld + (size_t)(k + b) * NB;     // tile (k, k + b)
    constexpr int IMG2 = NB * TS / 2;
    load_tile_async(sC, gC, ld);
    async_wait_all();
    __syncthreads();
    if (diag) {
        factor_block(sC, pl);
        if (t < NB) dinv[t] = 1.0 / sC[t * TS + t];
        __syncthreads();
        for (int e = t; e < NB * NB; e += LU_THREADS) {
            const int r = e >> 6, c = e & 63;
            const double v = sC[r * TS + c];
            sA[r * TS + c] = (c > r) ? v * dinv[r] : v;
        }
        __syncthreads();
        {
            double2* g0 = reinterpret_cast<double2*>(send);
            const double2* s0 = reinterpret_cast<const double2*>(sA);
#pragma unroll
            for (int e = t; e < IMG2; e += LU_THREADS) g0[e] = s0[e];
            if (t < NB) send[NB * TS + t] = dinv[t];
        }
        __threadfence();
        __syncthreads();
        if (t == 0) st_release(flags + k, epoch);
        store_tile(gC, ld, sC);
        pivot_guard<LU_THREADS>(sC, pl, k, ld / NB, scalars, status);
        if (rhs != nullptr && t < 32) {
            forward_rhs_warp(sC, rhs + (size_t)k * NB);
            __syncwarp();
            send[PUB_DOUBLES + t] = rhs[(size_t)k * NB + t];
            send[PUB_DOUBLES + 32 + t] = rhs[(size_t)k * NB + 32 + t];
        }
    } else {
        if (t == 0) {
            while (ld_acquire(flags + k) != epoch) __nanosleep(32);
        }
        __syncthreads();
        const double2* g = reinterpret_cast<const double2*>(send);
        double2* sT2 = reinterpret_cast<double2*>(sA);
        double2 v[IMG2 / LU_THREADS];
#pragma unroll
        for (int u = 0; u < IMG2 / LU_THREADS; ++u) v[u] = __ldcg(g + t + u * LU_THREADS);
#pragma unroll
        for (int u = 0; u < IMG2 / LU_THREADS; ++u) sT2[t + u * LU_THREADS] = v[u];
        if (t < NB) dinv[t] = 1.0;
        __syncthreads();
    }
    {
        const int ty = t >> 3, tx = t & 7;
        const int sr = diag ? TS : 1, sc = diag ? 1 : TS;
        double a[4][8];
        if (diag) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) a[i][j] = (ty + 16 * i == tx + 8 * j) ? 1.0 : 0.0;
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) a[i][j] = sC[(ty + 16 * i) * sr + (tx + 8 * j) * sc];
        }
        if (diag) panel_solve<false>(a, sA, dinv);
        else panel_solve<true>(a, sA, dinv);
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) sC[(ty + 16 * i) * sr + (tx + 8 * j) * sc] = a[i][j];
    }
    __syncthreads();
    if (diag) {
        store_tile(uinv + (size_t)k * NB * NB, NB, sC);
    } else {
        store_tile(gC, ld, sC);
        store_tile(send + DIST_HDR + (size_t)(b - 1) * NB * NB, NB, sC);      // packed copy for the broadcast
    }
    (void)ncols;
}

// L[i,k] = A[i,k] U11^-1 and b_i -= L[i,k] y_k for the local block rows lb >= lb0 (their global index is > k)
__global__ void __launch_bounds__(LU_THREADS, 3) lu_dist_panel_kernel(double* __restrict__ Mloc, int ld, int k, int lb0, int P, int rank,
                                                                      const double* __restrict__ recv, double* __restrict__ rhs) {
    extern __shared__ __align__(16) double sm[];
    double* sA = sm;
    double* sC = sm + NB * TS;
    __shared__ __align__(16) double dinv[NB];
    __shared__ double yk[NB];
    const int t = threadIdx.x;
    const int lb = lb0 + blockIdx.x;
    double* gC = Mloc + (size_t)lb * NB * ld + (size_t)k * NB;
    constexpr int IMG2 = NB * TS / 2;
    load_tile_async(sC, gC, ld);
    {
        const double2* g = reinterpret_cast<const double2*>(recv);
        double2* sT2 = reinterpret_cast<double2*>(sA);
#pragma unroll
        for (int u = 0; u < IMG2 / LU_THREADS; ++u) sT2[t + u * LU_THREADS] = g[t + u * LU_THREADS];
        if (t < NB) {
            dinv[t] = recv[NB * TS + t];
            yk[t] = recv[PUB_DOUBLES + t];
        }
    }
    async_wait_all();
    __syncthreads();
    {
        const int ty = t >> 3, tx = t & 7;
        double a[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) a[i][j] = sC[(ty + 16 * i) * TS + tx + 8 * j];
        panel_solve<false>(a, sA, dinv);
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) sC[(ty + 16 * i) * TS + tx + 8 * j] = a[i][j];
    }
    __syncthreads();
    store_tile(gC, ld, sC);
    if (rhs != nullptr && t < NB) {
        double s = 0.0;
#pragma unroll 8
        for (int q = 0; q < NB; ++q) s = fma(sC[t * TS + q], yk[q], s);
        rhs[(size_t)(lb * P + rank) * NB + t] -= s;
    }
}

// A[i,j] -= L[i,k] U[k,j] for the local block rows lb >= lb0 and all block columns j > k (U tiles from the step buffer)
__global__ void __launch_bounds__(LU_THREADS, 3) lu_dist_update_kernel(double* __restrict__ Mloc, int ld, int k, int lb0, int ncols,
                                                                       const double* __restrict__ utiles) {
    extern __shared__ __align__(16) double sm[];
    double* sA = sm;
    double* sB = sm + NB * TS;
    const int lb = lb0 + blockIdx.x / ncols;
    const int jj = blockIdx.x % ncols;                 // block column k + 1 + jj
    double* gC = Mloc + (size_t)lb * NB * ld + (size_t)(k + 1 + jj) * NB;
    double acc[4][4][2];
    {
        double2 cn[4][4];
        c_fetch(cn, gC, ld);
        load_tile_async(sA, Mloc + (size_t)lb * NB * ld + (size_t)k * NB, ld);
        load_tile_async(sB, utiles + (size_t)jj * NB * NB, NB);
        async_wait_all();
        c_negate(acc, cn);
    }
    __syncthreads();
    mma_tile(sA, sB, acc);
    acc_io<false>(acc, gC, ld);
}

// x_k = inv(U_kk) (y_k - sum_{j>k} U[k,j] x_j) on the owner of block row k.  CTAs split the columns; partial sums meet in
// `acc` (64 doubles + a ticket, zero on entry and left zero), the last CTA to arrive finishes the block.
constexpr int BACK_CHUNK = 1024;
__global__ void __launch_bounds__(256) lu_dist_back_kernel(const double* __restrict__ Mloc, int ld, int k, int lbk,
                                                           const double* __restrict__ uinv, const double* __restrict__ y,
                                                           double* __restrict__ x, double* __restrict__ acc) {
    __shared__ double sv[NB];
    __shared__ int last;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int c0 = (k + 1) * NB + blockIdx.x * BACK_CHUNK;
    const int c1 = min(c0 + BACK_CHUNK, ld);
    const double* rowbase = Mloc + (size_t)lbk * NB * ld;
    if (c0 < c1) {
#pragma unroll 1
        for (int r = warp * 8; r < warp * 8 + 8; ++r) {
            const double* row = rowbase + (size_t)r * ld;
            double s = 0.0;
            for (int c = c0 + lane; c < c1; c += 32) s = fma(row[c], __ldcg(x + c), s);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) atomicAdd(acc + r, s);
        }
    }
    __threadfence();
    __syncthreads();
    if (t == 0) {
        int* ticket = reinterpret_cast<int*>(acc + NB);
        last = (atomicAdd(ticket, 1) == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    if (t < NB) sv[t] = y[(size_t)k * NB + t] - __ldcg(acc + t);
    __syncthreads();
    if (t < NB) {
        const double* ui = uinv + (size_t)k * NB * NB + (size_t)t * NB;
        double s = 0.0;
#pragma unroll 8
        for (int q = 0; q < NB; ++q) s = fma(ui[q], sv[q], s);
        x[(size_t)k * NB + t] = s;
        acc[t] = 0.0;
    }
    if (t == 0) *reinterpret_cast<int*>(acc + NB) = 0;
}

#ifdef LLS_LU_CLOCKS
extern "C" int lls_debug_lu_clocks(long long* out, int steps) {
    return cudaMemcpyFromSymbol(out, g_lu_clocks, sizeof(long long) * 24 * (size_t)(steps < 256 ? steps : 256)) == cudaSuccess ? 0 : 1;
}
#endif

static bool g_attr_done = false;
static int g_pair_min_rem = 64;       // LLS_LU_PAIR_MIN_REM: rank-128 paired visits for trailing matrices of at least this many block rows
static bool g_lu_pairs = true;   // LLS_LU_PAIRS=0 falls back to one rank-64 update per launch
static int g_worker_min_rem = 0;     // LLS_LU_WORKER_MIN_REM: trailing matrices of at least this many block rows use persistent workers (0 = never, the default: measured no faster than the tile-per-CTA kernel)
static int g_worker_chunk = 2;       // LLS_LU_WORKER_CHUNK: consecutive serpentine tiles per grab
static int g_sms = 148;
static int g_worker_dbg = 0;
static bool g_chain_guard = true;     // LLS_LU_CHAIN_GUARD=0: update CTAs do not yield the diag CTA's SM
static int g_worker_reserve = 48;     // LLS_LU_WORKER_RESERVE: CTA slots left to the panel tiles at the start of a launch

static int g_pack_min_rem = 0;        // LLS_LU_PACK_MIN_REM: trailing matrices of at least this many block rows claim roles by SM (0 = never, the default: the claim at CTA start costs more than the isolation returns, 2.83 -> 2.95 ms at N = 4000)
static int g_split4_max_rem = 8;      // LLS_LU_SPLIT4_MAX_REM: ... and four CTAs per panel tile up to this (N = 200 0.107 -> 0.103 ms, N = 1000 0.432 -> 0.422 ms; slower again from 16 up)
static int g_split_max_rem = 36;      // LLS_LU_SPLIT_MAX_REM: trailing matrices of at most this many block rows use two CTAs per panel tile (0 = never)
static bool g_lu_inv_cta = true;      // LLS_LU_INV_CTA=0: the diag CTA inverts U11 itself after publishing
static bool g_lu_pdl = true;          // LLS_LU_PDL=0: plain stream-ordered launches of the block steps

template <class... KArgs, class... Args>
static cudaError_t launch_step(void (*kern)(KArgs...), dim3 grid, int threads, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(threads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = g_lu_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

static int lu_prepare(Scene* s) {
    if (!g_attr_done) {
        const char* env = getenv("LLS_LU_PAIRS");
        if (env != nullptr && env[0] == '0') g_lu_pairs = false;
        env = getenv("LLS_LU_WORKER_MIN_REM");
        if (env != nullptr) g_worker_min_rem = atoi(env);
        env = getenv("LLS_LU_WORKER_CHUNK");
        if (env != nullptr && atoi(env) >= 1) g_worker_chunk = atoi(env);
        env = getenv("LLS_LU_PACK_MIN_REM");
        if (env != nullptr) g_pack_min_rem = atoi(env);
        env = getenv("LLS_LU_PAIR_MIN_REM");
        if (env != nullptr && atoi(env) >= 2) g_pair_min_rem = atoi(env);
        env = getenv("LLS_LU_SPLIT4_MAX_REM");
        if (env != nullptr) g_split4_max_rem = atoi(env);
        env = getenv("LLS_LU_SPLIT_MAX_REM");
        if (env != nullptr) g_split_max_rem = atoi(env);
        env = getenv("LLS_LU_INV_CTA");
        if (env != nullptr && env[0] == '0') g_lu_inv_cta = false;
        env = getenv("LLS_LU_PDL");
        if (env != nullptr && env[0] == '0') g_lu_pdl = false;
        env = getenv("LLS_LU_CHAIN_GUARD");
        if (env != nullptr && env[0] == '0') g_chain_guard = false;
        env = getenv("LLS_LU_WORKER_RESERVE");
        if (env != nullptr) g_worker_reserve = atoi(env);
        env = getenv("LLS_LU_WORKER_DBG");
        if (env != nullptr) g_worker_dbg = atoi(env);
        int dev = 0;
        LLS_CUDA_TRY(cudaGetDevice(&dev));
        LLS_CUDA_TRY(cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev));
        LLS_CUDA_TRY(cudaFuncSetAttribute(lu_step_kernel<true, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * TILE_BYTES));
        LLS_CUDA_TRY(cudaFuncSetAttribute(lu_step_kernel<false, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * TILE_BYTES));
        LLS_CUDA_TRY(cudaFuncSetAttribute(lu_step_kernel<false, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * TILE_BYTES));
        LLS_CUDA_TRY(cudaFuncSetAttribute(lu_step_kernel<false, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * TILE_BYTES));
        LLS_CUDA_TRY(cudaFuncSetAttribute(lu_step_kernel<true, 1, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * TILE_BYTES));
        LLS_CUDA_TRY(cudaFuncSetAttribute(lu_step_kernel<false, 1, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * TILE_BYTES));
        LLS_CUDA_TRY(cudaFuncSetAttribute(lu_step_kernel<true, 1, false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * TILE_BYTES));
        LLS_CUDA_TRY(cudaFuncSetAttribute(lu_step_kernel<false, 1, false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * TILE_BYTES));
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
    lu_backsolve_chain_kernel<<<nblk, BS_THREADS, 0, st>>>(M, ld, nblk, s->uinv, y, rhs, s->lu_flags + nblk + 1, ++s->lu_epoch, CaseStrides());
    count_launch();
    LLS_CUDA_TRY(cudaGetLastError());
    return LLS_OK;
}

int lu_factor_solve(Scene* s, double* M, double* rhs, double* x, cudaStream_t st, bool masked) {
    const int ld = s->ld, nblk = ld / NB;
    // an external matrix (lls_lu_factor) is always a single case; the scene's own matrix follows the bound batch
    const int ncases = (M == s->M) ? s->ncases : 1;
    const CaseStrides cs = (M == s->M) ? strides(s, masked) : CaseStrides();
    LLS_TRY(lu_prepare(s));
    const int epoch = ++s->lu_epoch;
    double* dbuf = s->inv;          // 2 x PUB_DOUBLES published diagonal factors
    int* flags = s->lu_flags;
    int* counters = s->lu_flags + 2 * (nblk + 1);   // one work counter per block step (zeroed by the first launch)
    int* diag_sm = g_chain_guard ? s->lu_flags + 3 * (nblk + 1) : nullptr;   // where the diag CTA announces its SM
    int* claims = s->lu_flags + 4 * (nblk + 1);     // two role-claim counters per block step (zeroed by the first launch)
    const bool workers_ok = (ncases == 1) && g_worker_min_rem > 0;
    // one more CTA per launch inverts U11 (see the INV role); not in a batch, where other cases hide one case's chain and a CTA more
    // per case and launch is only overhead (C3 sweep: 39.6 k -> 36.9 k cases/s with it)
    const int inv_cta = (g_lu_inv_cta && ncases == 1) ? 1 : 0;
    const bool split_ok = (ncases == 1);           // a batch hides the chain of one case behind the other cases: no redundant tile updates there
    {
        const int rem = nblk;       // block 0 and its panel
        if (split_ok && rem <= g_split4_max_rem) {
            const int roles = 1 + 8 * (rem - 1);
            LLS_CUDA_TRY(launch_step(lu_step_kernel<true, 1, false, 4>, dim3(roles + 4 * inv_cta, 1, ncases), LU_THREADS, (size_t)2 * TILE_BYTES, st,
                    M, ld, -1, rem, rhs, dbuf, s->uinv, flags, epoch, s->scalars, s->status, cs, workers_ok ? counters : nullptr, 0, 0, diag_sm, claims, 0,
                    inv_cta ? roles : -1));
        } else if (split_ok && rem <= g_split_max_rem) {
            const int roles = 1 + 4 * (rem - 1);
            LLS_CUDA_TRY(launch_step(lu_step_kernel<true, 1, false, 2>, dim3(roles + 2 * inv_cta, 1, ncases), LU_THREADS, (size_t)2 * TILE_BYTES, st,
                    M, ld, -1, rem, rhs, dbuf, s->uinv, flags, epoch, s->scalars, s->status, cs, workers_ok ? counters : nullptr, 0, 0, diag_sm, claims, 0,
                    inv_cta ? roles : -1));
        } else {
            const int grid = 2 * rem - 1;
            LLS_CUDA_TRY(launch_step(lu_step_kernel<true, 1, false>, dim3(grid + inv_cta, 1, ncases), LU_THREADS, (size_t)2 * TILE_BYTES, st,
                    M, ld, -1, rem, rhs, dbuf, s->uinv, flags, epoch, s->scalars, s->status, cs, workers_ok ? counters : nullptr, 0, 0, diag_sm, claims, 0,
                    inv_cta ? grid : -1));
        }
        count_launch();
    }
    for (int k = 0; k + 1 < nblk;) {
        const int rem = nblk - 1 - k;
        if (workers_ok && rem >= g_worker_min_rem) {
            // panel roles as CTAs 0 .. 2 rem - 2, everything else through persistent workers (two per SM)
            const long long tiles = (long long)(rem - 1) * (rem - 1);
            const long long chunks = (tiles + g_worker_chunk - 1) / g_worker_chunk;
            long long nworkers = 2LL * g_sms - 1 - g_worker_reserve;      // resident CTAs: diag + workers + `reserve` panel CTAs
            if (nworkers > chunks) nworkers = chunks;
            if (nworkers < 1) nworkers = 1;
            LLS_CUDA_TRY(launch_step(lu_step_kernel<false, 1, true>, dim3(2 * rem - 1 + (int)nworkers, 1, 1), 256, (size_t)3 * TILE_BYTES, st,
                M, ld, k, rem, rhs, dbuf, s->uinv, flags, epoch, s->scalars, s->status, cs, counters, g_worker_chunk | (g_worker_dbg << 8), (int)nworkers, diag_sm, nullptr, 0, -1));
            count_launch();
            k += 1;
            continue;
        }
        // pairs pay off once the bulk of a launch outweighs the exposed chain of the thin launch (measured on B200: N = 16000
        // 123 -> 113 ms, N = 8000 17.1 -> 16.5 ms, N = 4000 3.25 -> 3.33 ms when pairing everything)
        if (rem >= g_pair_min_rem && g_lu_pairs) {
            // thin launch: block row / column k+1 only (update with panel k, factor panel k+1) ...
            LLS_CUDA_TRY(launch_step(lu_step_kernel<false, 1, false>, dim3(2 * rem - 1 + inv_cta, 1, ncases), LU_THREADS, (size_t)2 * TILE_BYTES, st,
                M, ld, k, rem, rhs, dbuf, s->uinv, flags, epoch, s->scalars, s->status, cs, nullptr, 0, 0, diag_sm, nullptr, 0,
                inv_cta ? 2 * rem - 1 : -1));
            count_launch();
            // ... then one visit of the rest with panels k and k+1 together, factoring panel k+2 on the way
            const int rem2 = rem - 1;
            const int psm2 = (g_pack_min_rem > 0 && rem2 >= g_pack_min_rem) ? (2 * rem2 - 2 + 2) / 3 : 0;
            const int inv2 = psm2 ? 0 : inv_cta;
            LLS_CUDA_TRY(launch_step(lu_step_kernel<false, 2, false>, dim3(rem2 * rem2 + inv2, 1, ncases), LU_THREADS, (size_t)2 * TILE_BYTES, st,
                M, ld, k, rem2, rhs, dbuf, s->uinv, flags, epoch, s->scalars, s->status, cs, nullptr, 0, 0, diag_sm, psm2 ? claims : nullptr, psm2,
                inv2 ? 2 * rem2 - 1 : -1));
            count_launch();
            k += 2;
        } else {
            const int psm = (g_pack_min_rem > 0 && rem >= g_pack_min_rem) ? (2 * rem - 2 + 2) / 3 : 0;
            if (split_ok && !psm && rem <= g_split_max_rem) {     // chain-bound regime: 2 or 4 CTAs per panel tile and for the inverse
                const int parts = (rem <= g_split4_max_rem) ? 4 : 2;
                const int roles = 1 + 2 * parts * (rem - 1);
                const dim3 grid(roles + parts * inv_cta + (rem - 1) * (rem - 1), 1, ncases);
                if (parts == 4) {
                    LLS_CUDA_TRY(launch_step(lu_step_kernel<false, 1, false, 4>, grid, LU_THREADS, (size_t)2 * TILE_BYTES, st, M, ld, k, rem, rhs, dbuf,
                        s->uinv, flags, epoch, s->scalars, s->status, cs, nullptr, 0, 0, diag_sm, nullptr, 0, inv_cta ? roles : -1));
                } else {
                    LLS_CUDA_TRY(launch_step(lu_step_kernel<false, 1, false, 2>, grid, LU_THREADS, (size_t)2 * TILE_BYTES, st, M, ld, k, rem, rhs, dbuf,
                        s->uinv, flags, epoch, s->scalars, s->status, cs, nullptr, 0, 0, diag_sm, nullptr, 0, inv_cta ? roles : -1));
                }
                count_launch();
                k += 1;
                continue;
            }
            const int inv1 = psm ? 0 : inv_cta;
            LLS_CUDA_TRY(launch_step(lu_step_kernel<false, 1, false>, dim3(rem * rem + inv1, 1, ncases), LU_THREADS, (size_t)2 * TILE_BYTES, st,
                M, ld, k, rem, rhs, dbuf, s->uinv, flags, epoch, s->scalars, s->status, cs, nullptr, 0, 0, diag_sm, psm ? claims : nullptr, psm,
                inv1 ? 2 * rem - 1 : -1));
            count_launch();
            k += 1;
        }
    }
    if (rhs) {
        lu_backsolve_chain_kernel<<<dim3(nblk, 1, ncases), BS_THREADS, 0, st>>>(M, ld, nblk, s->uinv, rhs, x, s->lu_flags + nblk + 1, epoch, cs);
        count_launch();
    }
    LLS_CUDA_TRY(cudaGetLastError());
    return LLS_OK;
}

// ---- host side of the row-partitioned LU (driven step by step from Python, which owns the NCCL broadcasts) -------------
size_t lu_dist_step_doubles(const Scene* s, int k) {
    const int nblk = s->ld / NB;
    return (size_t)DIST_HDR + (size_t)(nblk - 1 - k) * NB * NB;
}

int lu_dist_begin(Scene* s, cudaStream_t st) {
    LLS_TRY(lu_prepare(s));
    static bool attr = false;
    if (!attr) {
        LLS_CUDA_TRY(cudaFuncSetAttribute(lu_dist_owner_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * TILE_BYTES));
        LLS_CUDA_TRY(cudaFuncSetAttribute(lu_dist_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * TILE_BYTES));
        LLS_CUDA_TRY(cudaFuncSetAttribute(lu_dist_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * TILE_BYTES));
        attr = true;
    }
    // the publish flags are cleared in stream order instead of relying on a fresh epoch, so that a whole factorisation
    // (kernels + NCCL broadcasts) can be captured once into a CUDA graph and replayed with the same arguments
    LLS_CUDA_TRY(cudaMemsetAsync(s->lu_flags, 0, sizeof(int) * (size_t)(s->ld / NB + 1), st));
    if (s->lu_epoch == 0) s->lu_epoch = 1;
    return LLS_OK;
}

int lu_dist_owner(Scene* s, int k, double* rhs, double* send, cudaStream_t st) {
    const int nblk = s->ld / NB, P = s->cs.P;
    LLS_REQUIRE(k % P == s->cs.rank, LLS_ERR_ARG, "lu_dist_owner: this rank does not own block row k");
    const int ncols = nblk - 1 - k;
    lu_dist_owner_kernel<<<1 + ncols, LU_THREADS, 2 * TILE_BYTES, st>>>(s->M, s->ld, k, k / P, ncols, rhs, send, s->uinv, s->lu_flags, s->lu_epoch,
                                                                        s->scalars, s->status);
    count_launch();
    LLS_CUDA_TRY(cudaGetLastError());
    return LLS_OK;
}

int lu_dist_update(Scene* s, int k, double* rhs, const double* recv, cudaStream_t st) {
    const int nblk = s->ld / NB, P = s->cs.P, rank = s->cs.rank;
    const int ncols = nblk - 1 - k;
    // first local block whose global index lb * P + rank exceeds k
    int lb0 = (k + 1 - rank + P - 1) / P;
    if (lb0 < 0) lb0 = 0;
    const int nloc = s->n_loc_blocks - lb0;
    if (nloc <= 0) return LLS_OK;
    lu_dist_panel_kernel<<<nloc, LU_THREADS, 2 * TILE_BYTES, st>>>(s->M, s->ld, k, lb0, P, rank, recv, rhs);
    count_launch();
    if (ncols > 0) {
        lu_dist_update_kernel<<<nloc * ncols, LU_THREADS, 2 * TILE_BYTES, st>>>(s->M, s->ld, k, lb0, ncols, recv + DIST_HDR);
        count_launch();
    }
    LLS_CUDA_TRY(cudaGetLastError());
    return LLS_OK;
}

int lu_dist_back(Scene* s, int k, const double* y, double* x, cudaStream_t st) {
    const int P = s->cs.P;
    LLS_REQUIRE(k % P == s->cs.rank, LLS_ERR_ARG, "lu_dist_back: this rank does not own block row k");
    const int cols = s->ld - (k + 1) * NB;
    const int grid = cols > 0 ? (cols + BACK_CHUNK - 1) / BACK_CHUNK : 1;
    lu_dist_back_kernel<<<grid, 256, 0, st>>>(s->M, s->ld, k, k / P, s->uinv, y, x, s->inv + 2 * PUB_DOUBLES);
    count_launch();
    LLS_CUDA_TRY(cudaGetLastError());
    return LLS_OK;
}

}  // namespace lls
