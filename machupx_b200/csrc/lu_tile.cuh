// Tile-level building blocks of the FP64 LU, shared by the single-GPU factorisation (lu.cu) and the row-partitioned one
// (lu_dist.cu): DMMA tile product, cp.async tile movement, the in-CTA factorisation of a 64 x 64 block, the shuffle-based
// panel solves, the pivot guard.  Everything here is __device__ __forceinline__ / constexpr: no symbols are emitted.
#pragma once
#include "lls_internal.h"

namespace lls {

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

// ---- pipelined publication of the factored diagonal block ---------------------------------------------------------------------
// The panel solves consume the factors in the same order the factorisation produces them: group J of a solve (pivots 8 J .. 8 J + 7)
// needs rows 8 J .. of U11 (L-panel tiles) or columns 8 J .. of L11 (U-panel tiles), and those are final as soon as group J of the
// factorisation is done.  So the diag CTA publishes group by group, straight from the registers that hold the finished rows / columns
// (no staging through shared memory), and the panel CTAs trail the factorisation by about two groups instead of starting after all
// 64 pivots: the serial chain of a block step loses most of the ~13 k cycles of a panel solve plus the publish / copy round trip.
// The flag word of a block carries (epoch << 4) + groups published; 8 = the whole block.  The stores of group J are fenced and
// flagged one group LATER (before the stores of group J + 1 are issued), when they have long left the SM: the fence costs nothing.
constexpr int PUB_GROUPS = 8;
__device__ __forceinline__ int pub_token(int epoch, int groups) { return (epoch << 4) + groups; }

template <int J>
__device__ __forceinline__ void publish_group(const double (&a)[4][8], const PivotLine& pl, double* __restrict__ img, int ty, int tx) {
    constexpr int I = J >> 1;
    // rows 8 J .. 8 J + 7 of U11 (scaled by the reciprocal pivot right of the diagonal): row p lives in register row I of the threads
    // with ty == p & 15
    const int pp = ty - ((8 * J) & 15);
    if (pp >= 0 && pp < 8) {
        const int p = 8 * J + pp;
        const double rp = pl.rcp[p];
#pragma unroll
        for (int j = J; j < 8; ++j) {
            const int c = tx + 8 * j;
            if (c > p) img[p * TS + c] = a[I][j] * rp;
            else if (c == p) img[p * TS + c] = a[I][j];
        }
        if (tx == pp) img[NB * TS + p] = rp;
    }
    // columns 8 J .. 8 J + 7 of L11: column 8 J + tx lives in register column J (unscaled multipliers) of the threads with that tx
    {
        const int p = 8 * J + tx;
        const double rp = pl.rcp[p];
#pragma unroll
        for (int i = I; i < 4; ++i) {
            const int r = ty + 16 * i;
            if (r > p) img[r * TS + p] = a[i][J] * rp;
        }
    }
}

template <int J, int NT>
__device__ __forceinline__ void factor_publish_step(double (&a)[4][8], PivotLine& pl, double* __restrict__ img, int* __restrict__ flag,
                                                    int epoch, int ty, int tx) {
    factor_group<J, NT>(a, pl, ty, tx);
    if (J > 0) __threadfence();          // the stores of group J - 1, issued a whole group ago
    panel_sync<NT>();                    // pl.rcp of this group complete; everybody's fence done
    if (J > 0 && threadIdx.x == 0) st_release(flag, pub_token(epoch, J));
    publish_group<J>(a, pl, img, ty, tx);
}

template <int NT = LU_THREADS, bool PIPE = false>
__device__ __forceinline__ void factor_block(double* __restrict__ sC, PivotLine& pl, double* __restrict__ img = nullptr,
                                             int* __restrict__ flag = nullptr, int epoch = 0) {
    const int t = threadIdx.x, ty = t >> 3, tx = t & 7;
    double a[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) a[i][j] = sC[(ty + 16 * i) * TS + tx + 8 * j];
    if (PIPE) {
        factor_publish_step<0, NT>(a, pl, img, flag, epoch, ty, tx);
        factor_publish_step<1, NT>(a, pl, img, flag, epoch, ty, tx);
        factor_publish_step<2, NT>(a, pl, img, flag, epoch, ty, tx);
        factor_publish_step<3, NT>(a, pl, img, flag, epoch, ty, tx);
        factor_publish_step<4, NT>(a, pl, img, flag, epoch, ty, tx);
        factor_publish_step<5, NT>(a, pl, img, flag, epoch, ty, tx);
        factor_publish_step<6, NT>(a, pl, img, flag, epoch, ty, tx);
        factor_publish_step<7, NT>(a, pl, img, flag, epoch, ty, tx);
        __threadfence();
        panel_sync<NT>();
        if (t == 0) st_release(flag, pub_token(epoch, PUB_GROUPS));
    } else {
    factor_group<0, NT>(a, pl, ty, tx);
    factor_group<1, NT>(a, pl, ty, tx);
    factor_group<2, NT>(a, pl, ty, tx);
    factor_group<3, NT>(a, pl, ty, tx);
    factor_group<4, NT>(a, pl, ty, tx);
    factor_group<5, NT>(a, pl, ty, tx);
    factor_group<6, NT>(a, pl, ty, tx);
    factor_group<7, NT>(a, pl, ty, tx);
    panel_sync<NT>();      // pl.rcp complete
    }
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

// The consumer side of the pipelined publication: group J of the solve waits until group J of the factors is published, fetches just
// that part of the image (8 rows of U11 and their reciprocal pivots for an L-panel tile or the inverse, 8 columns of L11 for a U-panel
// tile) into the shared-memory image at sW, and runs.  Threads 0 .. 127 of the CTA, barrier panel_sync<NT>().
template <int J, bool TR, int RI, int NT>
__device__ __forceinline__ void panel_group_piped(double (&a)[RI][8], double* __restrict__ sW, double* __restrict__ dinv,
                                                  const double* __restrict__ img, const int* __restrict__ flag, int epoch, int tx, int lane) {
    const int t = threadIdx.x;
    if (t == 0) {
        const int want = pub_token(epoch, J + 1);
        while (ld_acquire(flag) < want) __nanosleep(20);
    }
    panel_sync<NT>();
    if (TR) {       // columns 8 J .. 8 J + 7 of the image, all 64 rows: thread t fetches 4 doubles of row t / 2
        const int r = t >> 1, c = 8 * J + (t & 1) * 4;
        const double2* g = reinterpret_cast<const double2*>(img + r * TS + c);
        const double2 v0 = __ldcg(g), v1 = __ldcg(g + 1);
        double2* d = reinterpret_cast<double2*>(sW + r * TS + c);
        d[0] = v0;
        d[1] = v1;
    } else {        // rows 8 J .. 8 J + 7, all 64 columns: thread t fetches 4 doubles of row 8 J + t / 16
        const int r = 8 * J + (t >> 4), c = (t & 15) * 4;
        const double2* g = reinterpret_cast<const double2*>(img + r * TS + c);
        const double2 v0 = __ldcg(g), v1 = __ldcg(g + 1);
        double2* d = reinterpret_cast<double2*>(sW + r * TS + c);
        d[0] = v0;
        d[1] = v1;
        if (t < 8) dinv[8 * J + t] = __ldcg(img + NB * TS + 8 * J + t);
    }
    panel_sync<NT>();
    panel_group<J, TR, RI>(a, sW, dinv, tx, lane);
}

template <bool TR, int RI, int NT>
__device__ __forceinline__ void panel_solve_piped(double (&a)[RI][8], double* __restrict__ sW, double* __restrict__ dinv,
                                                  const double* __restrict__ img, const int* __restrict__ flag, int epoch) {
    const int tx = threadIdx.x & 7, lane = threadIdx.x & 31;
    panel_group_piped<0, TR, RI, NT>(a, sW, dinv, img, flag, epoch, tx, lane);
    panel_group_piped<1, TR, RI, NT>(a, sW, dinv, img, flag, epoch, tx, lane);
    panel_group_piped<2, TR, RI, NT>(a, sW, dinv, img, flag, epoch, tx, lane);
    panel_group_piped<3, TR, RI, NT>(a, sW, dinv, img, flag, epoch, tx, lane);
    panel_group_piped<4, TR, RI, NT>(a, sW, dinv, img, flag, epoch, tx, lane);
    panel_group_piped<5, TR, RI, NT>(a, sW, dinv, img, flag, epoch, tx, lane);
    panel_group_piped<6, TR, RI, NT>(a, sW, dinv, img, flag, epoch, tx, lane);
    panel_group_piped<7, TR, RI, NT>(a, sW, dinv, img, flag, epoch, tx, lane);
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
// `first` / `last`: this is the first / last diagonal block THIS rank factors in the factorisation (row-partitioned solve: every rank
// guards the pivots of its own block rows, the host ORs the status words of the ranks).
__device__ __forceinline__ void pivot_guard(const double* __restrict__ sC, PivotLine& pl, bool first, bool last, double* __restrict__ scalars,
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
        if (!first) {
            m = fmin(m, scalars[1]);
            M = fmax(M, scalars[2]);
        }
        scalars[1] = m;
        scalars[2] = M;
        if (last && m < PIVOT_RATIO * M) atomicOr(status, LLS_STATUS_SMALL_PIVOT);
    }
}


}  // namespace lls
