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
#include "lu_tile.cuh"
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
        while (ld_acquire(g.flag) < pub_token(g.epoch, PUB_GROUPS)) __nanosleep(128);
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
                                                                int* __restrict__ claims, int panel_sms, int inv_block, int pipe) {
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
        // pipe: the factors are published group by group from inside the factorisation (lu_tile.cuh), the panel CTAs trail it
        if (pipe) factor_block<NT, true>(sC, pl, img, flags + blk, epoch);
        else factor_block<NT>(sC, pl);
        LU_CLK(4);
        if (t < NB) dinv[t] = 1.0 / sC[t * TS + t];
        panel_sync<NT>();
        for (int e = t; e < NB * NB; e += LU_THREADS) {
            const int r = e >> 6, c = e & 63;
            const double v = sC[r * TS + c];
            sA[r * TS + c] = (c > r) ? v * dinv[r] : v;
        }
        panel_sync<NT>();
        if (!pipe) {
            double2* g0 = reinterpret_cast<double2*>(img);
            const double2* s0 = reinterpret_cast<const double2*>(sA);
#pragma unroll
            for (int e = t; e < IMG2; e += LU_THREADS) g0[e] = s0[e];
            if (t < NB) img[NB * TS + t] = dinv[t];
            __threadfence();
            panel_sync<NT>();
            if (t == 0) st_release(flags + blk, pub_token(epoch, PUB_GROUPS));
        }
        LU_CLK(5);
        // everything below is off the critical path of the panel CTAs
        store_tile(gC, ld, sC);
        pivot_guard<NT>(sC, pl, blk == 0, blk == ld / NB - 1, scalars, status);
        if (rhs != nullptr && t < 32) forward_rhs_warp(sC, rhs + (size_t)blk * NB);
        // W = U11, pre-scaled rows in sA: the solve below turns the identity into inv(U11) (unless an INV CTA does that)
    } else if (pipe) {
        if (t < NB) dinv[t] = 1.0;      // U-panel tiles: unit diagonal; L-panel tiles / inverse: overwritten group by group
    } else {
        // panel roles: wait for the factored diagonal block
        if (t == 0) {
            while (ld_acquire(flags + blk) < pub_token(epoch, PUB_GROUPS)) __nanosleep(32);
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
        if (pipe && role != DIAG) {
            if (role == UROW) panel_solve_piped<true, 4, NT>(a, sA, dinv, img, flags + blk, epoch);
            else panel_solve_piped<false, 4, NT>(a, sA, dinv, img, flags + blk, epoch);
        } else if (role == UROW) panel_solve<true>(a, sA, dinv);
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
        if (pipe) {
            if (role == UROW) panel_solve_piped<true, RI, NT>(a, sA, dinv, img, flags + blk, epoch);
            else panel_solve_piped<false, RI, NT>(a, sA, dinv, img, flags + blk, epoch);
        } else if (role == UROW) panel_solve<true, RI>(a, sA, dinv);
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


// ---- small systems in large batches: one CTA factors and solves one case ---------------------------------------------------------
// BASELINE.json configs[2] (FlyingWing sweep: 4,096 flight states of a 180-point scene, ld = 192 = 3 block rows).  The general
// kernel gives every tile of every case a CTA and lets the panel CTAs spin on their case's diagonal block: four out of five resident
// CTAs are waiting at any time.  Here ONE 128-thread CTA walks through the whole right-looking factorisation of its case (diagonal
// factorisation, U block row, L block column with the right-hand side, DMMA trailing update, tiles through the L2) and then through
// the back substitution with the U blocks themselves (no inverse diagonal blocks needed): nothing ever waits for another CTA, and the
// three CTAs an SM holds keep its FP64 pipe busy with three different cases.
constexpr int SMALL_MAX_NBLK = 4;

// x = U^-1 b for the upper-triangular factor held row-major in sU (row stride TS); warp 0 only, lane l keeps entries l and l + 32
__device__ __forceinline__ void backward_warp(const double* __restrict__ sU, double* __restrict__ b) {
    const int lane = threadIdx.x;
    double x0 = b[lane], x1 = b[lane + 32];
    const double d0 = 1.0 / sU[lane * TS + lane], d1 = 1.0 / sU[(lane + 32) * TS + lane + 32];
#pragma unroll 1
    for (int q = NB - 1; q >= 0; --q) {
        // x_q is final once every column right of q has been folded in: scale it where it lives, then broadcast
        if (q >= 32) { if (lane == q - 32) x1 *= d1; } else { if (lane == q) x0 *= d0; }
        const double xq = __shfl_sync(0xffffffffu, (q < 32) ? x0 : x1, q & 31);
        if (lane < q) x0 = fma(-sU[lane * TS + q], xq, x0);
        if (lane + 32 < q) x1 = fma(-sU[(lane + 32) * TS + q], xq, x1);
    }
    b[lane] = x0;
    b[lane + 32] = x1;
}

__global__ void __launch_bounds__(LU_THREADS, 3) lu_small_case_kernel(double* __restrict__ M, int ld, int nblk, double* __restrict__ rhs,
                                                                      double* __restrict__ x, double* __restrict__ scalars,
                                                                      int* __restrict__ status, CaseStrides cs) {
    if (LLS_CASE_INACTIVE(cs)) return;
    M = case_ptr(M, cs.M);
    rhs = case_work_ptr(rhs, cs.work);
    x = case_work_ptr(x, cs.work);
    scalars = case_work_ptr(scalars, cs.work);
    status = case_work_ptr(status, cs.work);
    extern __shared__ __align__(16) double sm[];
    double* sA = sm;                 // image of the factored diagonal block / L operand
    double* sB = sm + NB * TS;       // working tile / U operand
    __shared__ __align__(16) PivotLine pl;
    __shared__ __align__(16) double dinv[NB];
    __shared__ __align__(16) double vec[NB];
    const int t = threadIdx.x, ty = t >> 3, tx = t & 7;
    auto tile = [&](int bi, int bj) { return M + (size_t)bi * NB * ld + (size_t)bj * NB; };
    for (int k = 0; k < nblk; ++k) {
        // diagonal block: factor, keep the pre-scaled image in sA
        load_tile_async(sB, tile(k, k), ld);
        async_wait_all();
        __syncthreads();
        factor_block(sB, pl);
        if (t < NB) dinv[t] = 1.0 / sB[t * TS + t];
        __syncthreads();
        for (int e = t; e < NB * NB; e += LU_THREADS) {
            const int r = e >> 6, c = e & 63;
            const double v = sB[r * TS + c];
            sA[r * TS + c] = (c > r) ? v * dinv[r] : v;
        }
        store_tile(tile(k, k), ld, sB);
        pivot_guard<LU_THREADS>(sB, pl, k == 0, k == nblk - 1, scalars, status);
        if (t < 32) forward_rhs_warp(sB, rhs + (size_t)k * NB);          // y_k
        __syncthreads();
        if (t < NB) vec[t] = rhs[(size_t)k * NB + t];
        // U block row: U[k,j] = L11^-1 A[k,j] (transposed problem: W = L11^T, unit diagonal)
        for (int j = k + 1; j < nblk; ++j) {
            __syncthreads();
            load_tile_async(sB, tile(k, j), ld);
            async_wait_all();
            __syncthreads();
            double a[4][8];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) a[i][jj] = sB[(ty + 16 * i) + (tx + 8 * jj) * TS];
            panel_solve<true>(a, sA, dinv);          // (unit diagonal: dinv is not read)
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) sB[(ty + 16 * i) + (tx + 8 * jj) * TS] = a[i][jj];
            __syncthreads();
            store_tile(tile(k, j), ld, sB);
        }
        // L block column: L[i,k] = A[i,k] U11^-1, b_i -= L[i,k] y_k
        for (int i2 = k + 1; i2 < nblk; ++i2) {
            __syncthreads();
            load_tile_async(sB, tile(i2, k), ld);
            async_wait_all();
            __syncthreads();
            double a[4][8];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) a[i][jj] = sB[(ty + 16 * i) * TS + tx + 8 * jj];
            panel_solve<false>(a, sA, dinv);
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) sB[(ty + 16 * i) * TS + tx + 8 * jj] = a[i][jj];
            __syncthreads();
            store_tile(tile(i2, k), ld, sB);
            if (t < NB) {
                double sdot = 0.0;
#pragma unroll 8
                for (int q = 0; q < NB; ++q) sdot = fma(sB[t * TS + q], vec[q], sdot);
                rhs[(size_t)i2 * NB + t] -= sdot;
            }
        }
        // trailing update on the FP64 tensor path: A[i,j] -= L[i,k] U[k,j]
        for (int i2 = k + 1; i2 < nblk; ++i2) {
            for (int j = k + 1; j < nblk; ++j) {
                __syncthreads();          // the tiles just stored by this CTA are visible to it; the operand slots are free
                double acc[4][4][2];
                {
                    double2 cn[4][4];
                    c_fetch(cn, tile(i2, j), ld);
                    if (j == k + 1) load_tile_async(sA, tile(i2, k), ld);      // the L operand stays for the whole tile row
                    load_tile_async(sB, tile(k, j), ld);
                    async_wait_all();
                    c_negate(acc, cn);
                }
                __syncthreads();
                mma_tile(sA, sB, acc);
                acc_io<false>(acc, tile(i2, j), ld);
            }
        }
        __syncthreads();
    }
    // back substitution: x_k = U_kk^-1 (y_k - sum_{j > k} U[k,j] x_j), block rows from the last to the first
    for (int k = nblk - 1; k >= 0; --k) {
        double s0 = (t < NB) ? rhs[(size_t)k * NB + t] : 0.0;
        for (int j = k + 1; j < nblk; ++j) {
            __syncthreads();
            load_tile_async(sB, tile(k, j), ld);
            if (t < NB) vec[t] = x[(size_t)j * NB + t];
            async_wait_all();
            __syncthreads();
            if (t < NB) {
                double sdot = 0.0;
#pragma unroll 8
                for (int q = 0; q < NB; ++q) sdot = fma(sB[t * TS + q], vec[q], sdot);
                s0 -= sdot;
            }
        }
        __syncthreads();
        load_tile_async(sB, tile(k, k), ld);
        if (t < NB) vec[t] = s0;
        async_wait_all();
        __syncthreads();
        if (t < 32) backward_warp(sB, vec);
        __syncthreads();
        if (t < NB) x[(size_t)k * NB + t] = vec[t];
        __threadfence_block();
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
static int g_small_batch_min = 16;    // LLS_LU_SMALL_BATCH_MIN: batches of at least this many cases of at most 4 block rows take the one-CTA-per-case kernel (0 = never)
static int g_lu_pipe = 0;             // LLS_LU_PIPE=1: the diag CTA publishes its factors group by group from inside the factorisation and the panel solves trail it (lu_tile.cuh).  Correct, measured SLOWER on B200 (N = 4000: 2.78 -> 2.84 ms, N = 1000: 0.429 -> 0.457 ms): eight extra fences / barriers / scattered stores in the factorisation and eight flag round trips in every panel CTA cost more than the ~13 k cycles of overlap return; off by default
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
        env = getenv("LLS_LU_SMALL_BATCH_MIN");
        if (env != nullptr) g_small_batch_min = atoi(env);
        env = getenv("LLS_LU_PIPE");
        if (env != nullptr) g_lu_pipe = atoi(env) ? 1 : 0;
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
        LLS_CUDA_TRY(cudaFuncSetAttribute(lu_small_case_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * TILE_BYTES));
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
    if (ncases > 1 && rhs != nullptr && nblk <= SMALL_MAX_NBLK && g_small_batch_min > 0 && ncases >= g_small_batch_min) {
        lu_small_case_kernel<<<dim3(1, 1, ncases), LU_THREADS, 2 * TILE_BYTES, st>>>(M, ld, nblk, rhs, x, s->scalars, s->status, cs);
        count_launch();
        LLS_CUDA_TRY(cudaGetLastError());
        return LLS_OK;
    }
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
    const int pipe = (g_lu_pipe && ncases == 1) ? 1 : 0;      // a batch hides one case's chain behind the other cases
    const bool split_ok = (ncases == 1);           // a batch hides the chain of one case behind the other cases: no redundant tile updates there
    {
        const int rem = nblk;       // block 0 and its panel
        if (split_ok && rem <= g_split4_max_rem) {
            const int roles = 1 + 8 * (rem - 1);
            LLS_CUDA_TRY(launch_step(lu_step_kernel<true, 1, false, 4>, dim3(roles + 4 * inv_cta, 1, ncases), LU_THREADS, (size_t)2 * TILE_BYTES, st,
                    M, ld, -1, rem, rhs, dbuf, s->uinv, flags, epoch, s->scalars, s->status, cs, workers_ok ? counters : nullptr, 0, 0, diag_sm, claims, 0,
                    inv_cta ? roles : -1, pipe));
        } else if (split_ok && rem <= g_split_max_rem) {
            const int roles = 1 + 4 * (rem - 1);
            LLS_CUDA_TRY(launch_step(lu_step_kernel<true, 1, false, 2>, dim3(roles + 2 * inv_cta, 1, ncases), LU_THREADS, (size_t)2 * TILE_BYTES, st,
                    M, ld, -1, rem, rhs, dbuf, s->uinv, flags, epoch, s->scalars, s->status, cs, workers_ok ? counters : nullptr, 0, 0, diag_sm, claims, 0,
                    inv_cta ? roles : -1, pipe));
        } else {
            const int grid = 2 * rem - 1;
            LLS_CUDA_TRY(launch_step(lu_step_kernel<true, 1, false>, dim3(grid + inv_cta, 1, ncases), LU_THREADS, (size_t)2 * TILE_BYTES, st,
                    M, ld, -1, rem, rhs, dbuf, s->uinv, flags, epoch, s->scalars, s->status, cs, workers_ok ? counters : nullptr, 0, 0, diag_sm, claims, 0,
                    inv_cta ? grid : -1, pipe));
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
                M, ld, k, rem, rhs, dbuf, s->uinv, flags, epoch, s->scalars, s->status, cs, counters, g_worker_chunk | (g_worker_dbg << 8), (int)nworkers, diag_sm, nullptr, 0, -1, pipe));
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
                inv_cta ? 2 * rem - 1 : -1, pipe));
            count_launch();
            // ... then one visit of the rest with panels k and k+1 together, factoring panel k+2 on the way
            const int rem2 = rem - 1;
            const int psm2 = (g_pack_min_rem > 0 && rem2 >= g_pack_min_rem) ? (2 * rem2 - 2 + 2) / 3 : 0;
            const int inv2 = psm2 ? 0 : inv_cta;
            LLS_CUDA_TRY(launch_step(lu_step_kernel<false, 2, false>, dim3(rem2 * rem2 + inv2, 1, ncases), LU_THREADS, (size_t)2 * TILE_BYTES, st,
                M, ld, k, rem2, rhs, dbuf, s->uinv, flags, epoch, s->scalars, s->status, cs, nullptr, 0, 0, diag_sm, psm2 ? claims : nullptr, psm2,
                inv2 ? 2 * rem2 - 1 : -1, pipe));
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
                        s->uinv, flags, epoch, s->scalars, s->status, cs, nullptr, 0, 0, diag_sm, nullptr, 0, inv_cta ? roles : -1, pipe));
                } else {
                    LLS_CUDA_TRY(launch_step(lu_step_kernel<false, 1, false, 2>, grid, LU_THREADS, (size_t)2 * TILE_BYTES, st, M, ld, k, rem, rhs, dbuf,
                        s->uinv, flags, epoch, s->scalars, s->status, cs, nullptr, 0, 0, diag_sm, nullptr, 0, inv_cta ? roles : -1, pipe));
                }
                count_launch();
                k += 1;
                continue;
            }
            const int inv1 = psm ? 0 : inv_cta;
            LLS_CUDA_TRY(launch_step(lu_step_kernel<false, 1, false>, dim3(rem * rem + inv1, 1, ncases), LU_THREADS, (size_t)2 * TILE_BYTES, st,
                M, ld, k, rem, rhs, dbuf, s->uinv, flags, epoch, s->scalars, s->status, cs, nullptr, 0, 0, diag_sm, psm ? claims : nullptr, psm,
                inv1 ? 2 * rem - 1 : -1, pipe));
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

}  // namespace lls
