// Row-partitioned FP64 LU for one large scene over the GPUs of a box (BASELINE.json configs[3]-[4]).
#include "lls_internal.h"
#include "lu_tile.cuh"
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

namespace lls {

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
                                                                      double* __restrict__ scalars, int* __restrict__ status, int last_owned) {
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
        pivot_guard<LU_THREADS>(sC, pl, lbk == 0, last_owned != 0, scalars, status);
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


// ---- host side of the row-partitioned LU (driven step by step from Python, which owns the NCCL broadcasts) -------------
size_t lu_dist_step_doubles(const Scene* s, int k) {
    const int nblk = s->ld / NB;
    return (size_t)DIST_HDR + (size_t)(nblk - 1 - k) * NB * NB;
}

int lu_dist_begin(Scene* s, cudaStream_t st) {
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
                                                                        s->scalars, s->status, k + P >= nblk ? 1 : 0);
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


// =====================================================================================================================
// The same factorisation with the exchange INSIDE the kernels, over NVLink peer memory (one process per GPU, the mailboxes
// of all ranks mapped into every process through CUDA IPC): no NCCL call and no host code between two block steps.
//
//   mailbox of a rank (device memory it owns, written by everybody):
//       slots[S]        panel step buffers [diag image | y_k | U tiles], slot = global step index mod S
//       flags[RING]     flags[G mod RING] == G + 1 once the panel of global step G has landed in this rank's slot
//       progress[P]     progress[q] = number of global steps rank q has finished consuming (gates the reuse of a slot)
//       xbuf[ld], xflags[nblk]   the replicated solution of the back substitution and its arrival flags (token = solve number)
//
//   owner(k):   lu_dist_owner_p2p_kernel factors the diagonal tile, solves its U block row and STORES the step buffer into the
//               slot of every rank (P2P stores through NVSwitch); the last CTA to finish releases the arrival flags
//               (st.release.sys) - one fused compute + broadcast kernel instead of kernel -> ncclBroadcast -> kernel.
//   all ranks:  lu_dist_panel_p2p_kernel acquires the flag (ld.acquire.sys on its own memory), solves its part of the L panel;
//               lu_dist_update_kernel reads the U tiles from the local slot.
//   look-ahead: the rank that owns block row k+1 updates that row first, factors it and pushes panel k+1 BEFORE it updates the
//               rest of its rows with panel k, so panel k+1 is in everybody's mailbox while they are still busy with step k
//               (launch order in one stream; nothing waits on the host).
//   back substitution: ONE launch per rank, a chain of CTAs (one per local block row) exactly like the single-GPU chain kernel;
//               x_k is stored into every rank's xbuf by its owner.
// Global step index G = (factorisations so far) * nblk + k runs identically on every rank, so tokens never repeat and nothing
// is ever reset.  Every spin carries a deadline: a rank that never arrives raises LLS_STATUS_COMM instead of hanging the box.
// =====================================================================================================================
constexpr int DIST_MAX_P = 8;
constexpr int DIST_RING = 64;
constexpr long long DIST_TIMEOUT_NS = 30LL * 1000 * 1000 * 1000;

struct DistPeers {
    double* slots[DIST_MAX_P];
    int* flags[DIST_MAX_P];       // "panel G complete" (diag image, y and every U tile have landed)
    int* iflags[DIST_MAX_P];      // "diag image of panel G (+ y) has landed": all the L-panel tiles need
    int* progress[DIST_MAX_P];
    double* xbuf[DIST_MAX_P];
    int* xflags[DIST_MAX_P];
    double* mc_slots;         // the slots / the solution of EVERY rank at once: multicast mapping of the mailboxes through the NVSwitch
    double* mc_xbuf;          // (NVLS; null without it): one store lands in all mailboxes, the owner's link carries a panel once, not P - 1 times
    int* own;                 // local: ring of intra-kernel flags of the owner kernel (diag tile published to the U-row CTAs)
    int* counter;             // local: CTAs of the owner kernel that have finished their stores
    int* abort;               // local: set once a deadline has passed: every later wait returns at once
    unsigned long long* stats;   // local: [0] ns the panel kernels waited for a panel, [1] ns the owner kernels waited for a free slot, [2] waits
    int P, rank, S;
    size_t slot_doubles;
};

struct DistComm {
    char* base[DIST_MAX_P] = {nullptr};     // mailbox of every rank as mapped into this process ([rank] = the one we own)
    char* mc = nullptr;                     // multicast mapping of all mailboxes (external symmetric memory only)
    bool external = false;                  // the mailboxes belong to the caller (lls_dist_comm_attach): nothing to free or close
    bool opened[DIST_MAX_P] = {false};
    size_t bytes = 0;
    DistPeers peers;
    long long gstep = 0;                    // global step index of the next factorisation's block step 0
    int solves = 0;                         // back substitutions so far (token of xflags)
    bool connected = false;
};

__device__ __forceinline__ int ld_acquire_sys(const int* p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(int* p, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ long long global_ns() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// one thread: waits until *p (acquire, system scope) satisfies the predicate; false when the deadline passed or a previous wait gave up
template <class Pred>
__device__ __forceinline__ bool dist_wait(const int* p, Pred ok, const DistPeers& pr, int* status, unsigned long long* waited) {
    if (ok(ld_acquire_sys(p))) return true;
    const long long t0 = global_ns();
    int polls = 0;
    while (true) {
        if (ok(ld_acquire_sys(p))) break;
        __nanosleep(64);
        if ((++polls & 255) == 0) {
            if (*reinterpret_cast<volatile int*>(pr.abort) != 0) return false;
            if (global_ns() - t0 > DIST_TIMEOUT_NS) {
                *reinterpret_cast<volatile int*>(pr.abort) = 1;
                atomicOr(status, LLS_STATUS_COMM);
                return false;
            }
        }
    }
    if (waited != nullptr) atomicAdd(waited, (unsigned long long)(global_ns() - t0));
    return true;
}

// 16 bytes to a multicast address: the switch replicates the store into every rank's memory
__device__ __forceinline__ void mc_store2(double* p, double2 v) {
    const float4 b = *reinterpret_cast<const float4*>(&v);
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
}
__device__ __forceinline__ void mc_store1(double* p, double v) {
    asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
// 64 x 64 tile in shared memory (row stride TS) -> the same place in every rank's slot (row stride NB)
__device__ __forceinline__ void store_tile_all(const DistPeers& pr, size_t offset, const double* __restrict__ s) {
    if (pr.mc_slots != nullptr) {
        const int t = threadIdx.x, c2 = t & 31, r0 = t >> 5;
        double* g = pr.mc_slots + offset;
#pragma unroll 4
        for (int r = r0; r < NB; r += 4) mc_store2(g + (size_t)r * NB + 2 * c2, *reinterpret_cast<const double2*>(s + r * TS + 2 * c2));
        return;
    }
    for (int q = 0; q < pr.P; ++q) store_tile(pr.slots[q] + offset, NB, s);
}
// the image of a factored diagonal block (shared memory, IMG2 double2) and its reciprocal pivots into every OTHER rank's slot
__device__ __forceinline__ void store_image_all(const DistPeers& pr, size_t slot_off, const double* __restrict__ sA, const double* __restrict__ dinv) {
    constexpr int IMG2 = NB * TS / 2;
    const int t = threadIdx.x;
    const double2* s0 = reinterpret_cast<const double2*>(sA);
    if (pr.mc_slots != nullptr) {
        double* g = pr.mc_slots + slot_off;
#pragma unroll
        for (int e = t; e < IMG2; e += LU_THREADS) mc_store2(g + 2 * e, s0[e]);
        if (t < NB) mc_store1(g + NB * TS + t, dinv[t]);
        return;
    }
    for (int q = 0; q < pr.P; ++q) {
        if (q == pr.rank) continue;
        double2* g0 = reinterpret_cast<double2*>(pr.slots[q] + slot_off);
#pragma unroll
        for (int e = t; e < IMG2; e += LU_THREADS) g0[e] = s0[e];
        if (t < NB) pr.slots[q][slot_off + NB * TS + t] = dinv[t];
    }
}
// one double of the step buffer (y_k) / of the solution into every rank's mailbox
__device__ __forceinline__ void store_slot_all(const DistPeers& pr, size_t off, double v) {
    if (pr.mc_slots != nullptr) { mc_store1(pr.mc_slots + off, v); return; }
    for (int q = 0; q < pr.P; ++q) pr.slots[q][off] = v;
}
__device__ __forceinline__ void store_x_all(const DistPeers& pr, size_t off, double v) {
    if (pr.mc_xbuf != nullptr) { mc_store1(pr.mc_xbuf + off, v); return; }
    for (int q = 0; q < pr.P; ++q) pr.xbuf[q][off] = v;
}

__global__ void __launch_bounds__(LU_THREADS, 3) lu_dist_owner_p2p_kernel(double* __restrict__ Mloc, int ld, int k, int lbk, int ncols,
                                                                          double* __restrict__ rhs, DistPeers pr, int slot, int G,
                                                                          double* __restrict__ uinv, double* __restrict__ scalars,
                                                                          int* __restrict__ status) {
    extern __shared__ __align__(16) double sm[];
    double* sA = sm;
    double* sC = sm + NB * TS;
    __shared__ __align__(16) PivotLine pl;
    __shared__ __align__(16) double dinv[NB];
    const int t = threadIdx.x, b = blockIdx.x;
    const bool diag = (b == 0);
    double* gC = Mloc + (size_t)lbk * NB * ld + (size_t)(k + b) * NB;     // tile (k, k + b)
    constexpr int IMG2 = NB * TS / 2;
    const size_t slot_off = (size_t)slot * pr.slot_doubles;
    double* mine = pr.slots[pr.rank] + slot_off;                           // this rank's own copy of the step buffer
    load_tile_async(sC, gC, ld);
    if (t == 0 && G >= pr.S) {
        // the slot may be rewritten once every rank has consumed global step G - S (its previous tenant)
        for (int q = 0; q < pr.P; ++q)
            if (q != pr.rank) {
                const int need = G - pr.S + 1;
                dist_wait(pr.progress[pr.rank] + q, [need](int v) { return v >= need; }, pr, status, diag ? pr.stats + 1 : nullptr);
            }
    }
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
        {   // local copy first: the U-row CTAs of this launch are waiting for it
            double2* g0 = reinterpret_cast<double2*>(mine);
            const double2* s0 = reinterpret_cast<const double2*>(sA);
#pragma unroll
            for (int e = t; e < IMG2; e += LU_THREADS) g0[e] = s0[e];
            if (t < NB) mine[NB * TS + t] = dinv[t];
        }
        __threadfence();
        __syncthreads();
        if (t == 0) st_release(pr.own + (G % DIST_RING), G + 1);
        store_image_all(pr, slot_off, sA, dinv);
        store_tile(gC, ld, sC);
        pivot_guard<LU_THREADS>(sC, pl, k < pr.P, k + pr.P >= ld / NB, scalars, status);
        if (rhs != nullptr && t < 32) {
            forward_rhs_warp(sC, rhs + (size_t)k * NB);
            __syncwarp();
            const double y0 = rhs[(size_t)k * NB + t], y1 = rhs[(size_t)k * NB + 32 + t];
            store_slot_all(pr, slot_off + PUB_DOUBLES + t, y0);
            store_slot_all(pr, slot_off + PUB_DOUBLES + 32 + t, y1);
        }
    } else {
        if (t == 0) {
            const int want = G + 1;
            dist_wait(pr.own + (G % DIST_RING), [want](int v) { return v == want; }, pr, status, nullptr);
        }
        __syncthreads();
        const double2* g = reinterpret_cast<const double2*>(mine);
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
        store_tile_all(pr, slot_off + DIST_HDR + (size_t)(b - 1) * NB * NB, sC);      // the U tile into every rank's slot
    }
    // arrival: every CTA's stores are fenced at system scope before it is counted; the last one releases the flags everywhere
    __threadfence_system();
    __syncthreads();
    if (t == 0) {
        const int done = atomicAdd(pr.counter, 1);
        if (done == ncols) {
            *reinterpret_cast<volatile int*>(pr.counter) = 0;
            __threadfence_system();
            for (int q = 0; q < pr.P; ++q) st_release_sys(pr.flags[q] + (G % DIST_RING), G + 1);
        }
    }
}

// L[i,k] = A[i,k] U11^-1 and b_i -= L[i,k] y_k for the local block rows lb0 .. lb0 + gridDim.x - 1 (nrows may be 0: the launch
// then only reports this rank's progress).  Block 0 first tells every rank that this rank is done with global step G - 1.
__global__ void __launch_bounds__(LU_THREADS, 3) lu_dist_panel_p2p_kernel(double* __restrict__ Mloc, int ld, int k, int lb0, int nrows,
                                                                          DistPeers pr, int slot, int G, double* __restrict__ rhs,
                                                                          int* __restrict__ status) {
    extern __shared__ __align__(16) double sm[];
    double* sA = sm;
    double* sC = sm + NB * TS;
    __shared__ __align__(16) double dinv[NB];
    __shared__ double yk[NB];
    const int t = threadIdx.x;
    if (blockIdx.x == 0 && t < pr.P) st_release_sys(pr.progress[t] + pr.rank, G);
    if ((int)blockIdx.x >= nrows) return;
    const int lb = lb0 + blockIdx.x;
    const int P = pr.P, rank = pr.rank;
    double* gC = Mloc + (size_t)lb * NB * ld + (size_t)k * NB;
    constexpr int IMG2 = NB * TS / 2;
    load_tile_async(sC, gC, ld);
    if (t == 0) {
        const int want = G + 1;
        dist_wait(pr.flags[rank] + (G % DIST_RING), [want](int v) { return v == want; }, pr, status, blockIdx.x == 0 ? pr.stats : nullptr);
    }
    __syncthreads();
    const double* recv = pr.slots[rank] + (size_t)slot * pr.slot_doubles;
    {
        const double2* g = reinterpret_cast<const double2*>(recv);
        double2* sT2 = reinterpret_cast<double2*>(sA);
#pragma unroll
        for (int u = 0; u < IMG2 / LU_THREADS; ++u) sT2[t + u * LU_THREADS] = __ldcg(g + t + u * LU_THREADS);
        if (t < NB) {
            dinv[t] = __ldcg(recv + NB * TS + t);
            yk[t] = __ldcg(recv + PUB_DOUBLES + t);
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

// Back substitution as one launch per rank: CTA c owns local block row n_loc - 1 - c (global block row bi), folds x_k
// (k = last .. bi + 1) into y_bi as the flags arrive in ITS OWN mailbox, then x_bi = inv(U_bi,bi) y_bi goes into every rank's
// xbuf.  Producers on this rank have smaller block indices, producers elsewhere run on other GPUs: nobody waits on a CTA
// that cannot be scheduled.
__global__ void __launch_bounds__(NB) lu_dist_backsolve_p2p_kernel(const double* __restrict__ Mloc, int ld, int nblk, int n_loc,
                                                                   const double* __restrict__ uinv, const double* __restrict__ y,
                                                                   DistPeers pr, int token, int G_end, int* __restrict__ status) {
    __shared__ __align__(16) double sx[NB];
    __shared__ int alive;
    const int t = threadIdx.x;
    const int lb = n_loc - 1 - (int)blockIdx.x;
    const int bi = lb * pr.P + pr.rank;
    if (blockIdx.x == 0 && t < pr.P) st_release_sys(pr.progress[t] + pr.rank, G_end);   // every step of the factorisation consumed
    double yt = y[(size_t)bi * NB + t];
    double2 ui[NB / 2];
    {
        const double2* g = reinterpret_cast<const double2*>(uinv + (size_t)bi * NB * NB + (size_t)t * NB);
#pragma unroll
        for (int c = 0; c < NB / 2; ++c) ui[c] = g[c];
    }
    const double* xloc = pr.xbuf[pr.rank];
    const int* xfl = pr.xflags[pr.rank];
    for (int k = nblk - 1; k > bi; --k) {
        double2 u[NB / 2];
        {
            const double2* g = reinterpret_cast<const double2*>(Mloc + ((size_t)lb * NB + t) * ld + (size_t)k * NB);
#pragma unroll
            for (int c = 0; c < NB / 2; ++c) u[c] = g[c];
        }
        if (t == 0) alive = dist_wait(xfl + k, [token](int v) { return v == token; }, pr, status, nullptr) ? 1 : 0;
        __syncthreads();
        if (!alive) return;
        sx[t] = __ldcg(xloc + (size_t)k * NB + t);
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
        const double xv = (s0 + s1) + (s2 + s3);
        store_x_all(pr, (size_t)bi * NB + t, xv);
    }
    __threadfence_system();
    __syncthreads();
    if (t < pr.P) st_release_sys(pr.xflags[t] + bi, token);
    // the CTA of this rank's LOWEST block row keeps the launch alive until the block rows below it (owned by other ranks, at most
    // P - 1 of them) have landed here as well: when the launch ends, the whole of x is in this rank's xbuf
    if (lb == 0 && t < bi) dist_wait(xfl + t, [token](int v) { return v == token; }, pr, status, nullptr);
}


// ---- one launch per block step and rank: trailing update with panel k, and panel k + 1 factored / solved / exchanged on the way ----
// The single-GPU lu_step_kernel spread over the ranks.  Launch k of a rank covers ITS tiles of the trailing matrix (local block
// rows with global index > k, block columns > k): every CTA does C -= L[i,k] U[k,j] (L from the rank's own rows, U from its
// mailbox slot of panel k), then, by role,
//   DIAG  (tile (k+1,k+1); only on the rank that owns block row k+1): factors it, stores the image (+ y_{k+1}) into the slot of
//         panel k+1 of EVERY rank over NVLink and releases the image flags
//   UROW  (tiles (k+1, j); owner only): waits for the image locally, solves U[k+1,j] = L11^-1 C, stores it into its own rows and
//         into every rank's slot; the last of the owner's DIAG / UROW CTAs releases the "panel k+1 complete" flags everywhere
//   LCOL  (tiles (i, k+1); every rank): waits for the image to land in ITS mailbox, solves L[i,k+1] = C U11^-1, b_i -= L[i,k+1] y_{k+1}
//   OTHER stores C.
// So the exchange of panel k+1 and every panel solve run underneath the DMMA bulk of step k; a launch waits for nothing but the
// "panel k complete" flag that was released while launch k-1 was still running.  FIRST (k = -1): no update, panel 0 only.
// Grid order: DIAG, UROW (owner only), LCOL, OTHER -- producers of a flag are resident before anything on this GPU can wait on it.

// ---- the schedule of the fused-step factorisation, shared by the launcher, the kernel and the CPU tests (lls_dist_debug_*) -----------
// One launch of a rank: tiles with block rows / columns >= base = k + passes take panels k .. k + passes - 1, the roles factor / solve
// panel `base`.  thin: the grid stops after the panel roles.
struct DistLaunch {
    int k, passes, thin;      // first panel applied (-1: the first launch, nothing to apply), panels per visit, roles only
    int base, rem;            // first block row / column of the trailing matrix handled, and its size in blocks
    int lb0, nloc;            // first local block row with global index >= base, number of such local block rows
    int own;                  // this rank owns block row `base` (its local block lb0)
    long long grid;           // CTAs
};
__host__ __device__ inline DistLaunch dist_launch(int k, int passes, int thin, int nblk, int P, int rank, int n_loc_blocks) {
    DistLaunch L;
    L.k = k; L.passes = passes; L.thin = thin;
    L.base = k + passes;
    L.rem = nblk - L.base;
    L.lb0 = (L.base - rank + P - 1) / P;
    if (L.lb0 < 0) L.lb0 = 0;
    L.nloc = n_loc_blocks - L.lb0 > 0 ? n_loc_blocks - L.lb0 : 0;
    L.own = (L.base % P == rank) ? 1 : 0;
    const int below = L.nloc - L.own;
    L.grid = L.own + (L.own ? L.rem - 1 : 0) + below + ((k >= 0 && !thin) ? (long long)below * (L.rem - 1) : 0);
    if (L.grid < 1) L.grid = 1;
    return L;
}
enum DistRole { DIST_DIAG = 0, DIST_UROW = 1, DIST_LCOL = 2, DIST_OTHER = 3, DIST_NONE = 4 };
// block index -> role and tile (local block row lb, global block column bj).  Grid order: DIAG, UROW (owner only), LCOL, OTHER.
__host__ __device__ inline int dist_role(int b, int own, int rem, int base, int lb0, int nloc, int thin, int& lb, int& bj) {
    const int nd = own ? 1 : 0, nu = own ? rem - 1 : 0;
    const int rows_below = nloc - nd;               // local block rows with global index > base
    if (b < nd) { lb = lb0; bj = base; return DIST_DIAG; }
    if (b < nd + nu) { lb = lb0; bj = base + 1 + (b - nd); return DIST_UROW; }
    if (b < nd + nu + rows_below) { lb = lb0 + nd + (b - nd - nu); bj = base; return DIST_LCOL; }
    const int o = b - nd - nu - rows_below;
    if (thin || rem <= 1 || o >= rows_below * (rem - 1)) { lb = -1; bj = -1; return DIST_NONE; }   // the launch only reported progress
    lb = lb0 + nd + o / (rem - 1);
    bj = base + 1 + o % (rem - 1);
    return DIST_OTHER;
}
// default threshold of the rank-128 pairs: ~ 64 sqrt(P) block rows; without the multicast mapping the thin launch of a pair exposes
// the owner's (P - 1)-fold panel push, which costs more than the pair returns beyond four ranks (8 B200s: 1.91 -> 1.99 s at N = 32,000)
inline int dist_default_pair_min(int P, bool multicast) {
    int pair_min = 64;
    while ((long long)pair_min * pair_min < 4096LL * P) pair_min += 8;
    if (!multicast && P > 4) pair_min = 0;
    return pair_min;
}
// the launches of one factorisation of a rank, in order; returns their number (out may be null)
inline int dist_schedule(int nblk, int P, int rank, int n_loc_blocks, int pair_min, DistLaunch* out, int max_out) {
    int n = 0;
    auto add = [&](int k, int passes, int thin) {
        if (out != nullptr && n < max_out) out[n] = dist_launch(k, passes, thin, nblk, P, rank, n_loc_blocks);
        ++n;
    };
    add(-1, 1, 0);                                      // panel 0
    for (int k = 0; k + 1 < nblk;) {
        const int rem = nblk - 1 - k;                   // block rows of the trailing matrix after step k
        if (pair_min > 0 && rem >= pair_min && k + 2 < nblk) {
            add(k, 1, 1);                               // thin: block row / column k + 1 take panel k, panel k + 1 factored and exchanged
            add(k, 2, 0);                               // the rest takes panels k and k + 1 in one visit, panel k + 2 on the way
            k += 2;
        } else {
            add(k, 1, 0);
            k += 1;
        }
    }
    return n;
}

// arrival of a whole panel: every owner CTA (1 DIAG + rem - 1 UROW) calls this from one thread AFTER fencing its stores at system
// scope and synchronising; the last one to arrive releases the "panel complete" flags in every rank's mailbox
__device__ __forceinline__ void dist_arrive(const DistPeers& pr, int rem, int Gn) {
    const int done = atomicAdd(pr.counter, 1);
    if (done == rem - 1) {
        *reinterpret_cast<volatile int*>(pr.counter) = 0;
        __threadfence_system();
        for (int q = 0; q < pr.P; ++q) st_release_sys(pr.flags[q] + (Gn % DIST_RING), Gn + 1);
    }
}

// PASSES = 2 (rank-128 visit, as in the single-GPU kernel): the tiles take the updates of panels k AND k + 1 in one visit -- half the
// traffic on C, twice the tensor work per tile visit -- and the roles factor panel k + 2; the tiles of block row / column k + 1 were
// brought up to date, and panel k + 1 factored and exchanged, by a THIN launch just before (a PASSES = 1 launch whose grid stops
// after the panel roles: `thin`).  Used while the bulk of a launch outweighs the exposed chain of the thin launch.
template <bool FIRST, int PASSES = 1>
__global__ void __launch_bounds__(LU_THREADS, 3) lu_dist_step_kernel(double* __restrict__ Mloc, int ld, int k, int nblk, int lb0, int nloc,
                                                                     double* __restrict__ rhs, DistPeers pr, int Gn,
                                                                     double* __restrict__ uinv, double* __restrict__ scalars,
                                                                     int* __restrict__ status, int thin) {
    extern __shared__ __align__(16) double sm[];
    double* sA = sm;
    double* sB = sm + NB * TS;
    __shared__ __align__(16) PivotLine pl;
    __shared__ __align__(16) double dinv[NB];
    __shared__ double yk[NB];
    const int t = threadIdx.x, b = blockIdx.x;
    const int P = pr.P, rank = pr.rank;
    static_assert(!(FIRST && PASSES != 1), "the first launch has nothing to update");
    const int base = k + PASSES, rem = nblk - base;     // trailing matrix handled here: block rows / columns base .. nblk - 1
    const int Gc = Gn - PASSES;                         // global index of panel k (not FIRST); panel k + 1 is Gc + 1
    if (b == 0 && t < P) st_release_sys(pr.progress[t] + rank, FIRST ? Gn : Gc);    // every global step below this one is consumed here
    const bool own = (base % P == rank);                // this rank owns block row base: its local block lb0
    constexpr int DIAG = DIST_DIAG, UROW = DIST_UROW, LCOL = DIST_LCOL, OTHER = DIST_OTHER;
    int lb, bj;
    const int role = dist_role(b, own ? 1 : 0, rem, base, lb0, nloc, thin, lb, bj);
    if (role == DIST_NONE) return;                      // nothing (left) for this rank: the launch only reported progress
    const int bi = lb * P + rank;                       // global block row
    double* gC = Mloc + (size_t)lb * NB * ld + (size_t)bj * NB;
    constexpr int IMG2 = NB * TS / 2;
    const size_t next_off = (size_t)(Gn % pr.S) * pr.slot_doubles;
    double* mine = pr.slots[rank] + next_off;           // this rank's slot of panel k + 1
    if (!FIRST) {
        double acc[4][4][2];
        {
            double2 cn[4][4];
            // No flag is polled here: the L-panel CTAs of the launch that factored panel k did not exit before "panel k complete" had
            // arrived in this rank's mailbox (see the end of the LCOL role), so the kernel boundary already orders these loads after it
            c_fetch(cn, gC, ld);
            load_tile_async(sA, Mloc + (size_t)lb * NB * ld + (size_t)k * NB, ld);                                   // L[i,k]
            load_tile_async(sB, pr.slots[rank] + (size_t)(Gc % pr.S) * pr.slot_doubles + DIST_HDR + (size_t)(bj - base + PASSES - 1) * NB * NB, NB);   // U[k,j]
            async_wait_all();
            c_negate(acc, cn);
        }
        __syncthreads();
        mma_tile(sA, sB, acc);
        if (PASSES == 2) {
            __syncthreads();
            load_tile_async(sA, Mloc + (size_t)lb * NB * ld + (size_t)(k + 1) * NB, ld);                             // L[i,k+1]
            load_tile_async(sB, pr.slots[rank] + (size_t)((Gc + 1) % pr.S) * pr.slot_doubles + DIST_HDR + (size_t)(bj - base) * NB * NB, NB);   // U[k+1,j]
            async_wait_all();
            __syncthreads();
            mma_tile(sA, sB, acc);
        }
        if (role == OTHER) {
            acc_io<false>(acc, gC, ld);
            return;
        }
        __syncthreads();
        acc_io<false>(acc, sB, TS);
        __syncthreads();
    } else {
        load_tile_async(sB, gC, ld);
        async_wait_all();
        __syncthreads();
    }
    double* sC = sB;
    if (role == DIAG || role == UROW) {
        // the slot of panel k + 1 may be rewritten once every rank has consumed its previous tenant (global step Gn - S)
        if (t == 0 && Gn >= pr.S) {
            const int need = Gn - pr.S + 1;
            for (int q = 0; q < P; ++q)
                if (q != rank) dist_wait(pr.progress[rank] + q, [need](int v) { return v >= need; }, pr, status, role == DIAG ? pr.stats + 1 : nullptr);
        }
        __syncthreads();
    }
    if (role == DIAG) {
        factor_block(sC, pl);
        if (t < NB) dinv[t] = 1.0 / sC[t * TS + t];
        __syncthreads();
        for (int e = t; e < NB * NB; e += LU_THREADS) {
            const int r = e >> 6, c = e & 63;
            const double v = sC[r * TS + c];
            sA[r * TS + c] = (c > r) ? v * dinv[r] : v;
        }
        __syncthreads();
        {   // own slot first: this rank's UROW CTAs are waiting for it
            double2* g0 = reinterpret_cast<double2*>(mine);
            const double2* s0 = reinterpret_cast<const double2*>(sA);
#pragma unroll
            for (int e = t; e < IMG2; e += LU_THREADS) g0[e] = s0[e];
            if (t < NB) mine[NB * TS + t] = dinv[t];
        }
        __threadfence();
        __syncthreads();
        if (t == 0) st_release(pr.own + (Gn % DIST_RING), Gn + 1);
        store_image_all(pr, next_off, sA, dinv);
        if (t < 32) {       // y_{k+1} = L11^-1 b_{k+1}, part of the image every L-panel tile reads
            forward_rhs_warp(sC, rhs + (size_t)base * NB);
            __syncwarp();
            const double y0 = rhs[(size_t)base * NB + t], y1 = rhs[(size_t)base * NB + 32 + t];
            store_slot_all(pr, next_off + PUB_DOUBLES + t, y0);
            store_slot_all(pr, next_off + PUB_DOUBLES + 32 + t, y1);
        }
        __threadfence_system();
        __syncthreads();
        if (t < P) st_release_sys(pr.iflags[t] + (Gn % DIST_RING), Gn + 1);
        if (t == 0) dist_arrive(pr, rem, Gn);         // everything this CTA sends has been sent
        // off the critical path of everybody else
        store_tile(gC, ld, sC);
        pivot_guard<LU_THREADS>(sC, pl, base < P, base + P >= nblk, scalars, status);
        {   // inverse of U11 for the back substitution
            const int ty = t >> 3, tx = t & 7;
            double a[4][8];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) a[i][j] = (ty + 16 * i == tx + 8 * j) ? 1.0 : 0.0;
            panel_solve<false>(a, sA, dinv);
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) sC[(ty + 16 * i) * TS + tx + 8 * j] = a[i][j];
        }
        __syncthreads();
        store_tile(uinv + (size_t)base * NB * NB, NB, sC);
    } else {
        // panel roles: wait for the factored diagonal block of panel k + 1 (UROW: published locally by this launch's DIAG CTA;
        // LCOL: landed in this rank's mailbox from wherever block row k + 1 lives)
        if (t == 0) {
            const int want = Gn + 1;
            if (role == UROW) dist_wait(pr.own + (Gn % DIST_RING), [want](int v) { return v == want; }, pr, status, nullptr);
            else dist_wait(pr.iflags[rank] + (Gn % DIST_RING), [want](int v) { return v == want; }, pr, status, nullptr);
        }
        __syncthreads();
        {
            const double2* g = reinterpret_cast<const double2*>(mine);
            double2* sT2 = reinterpret_cast<double2*>(sA);
            double2 v[IMG2 / LU_THREADS];
#pragma unroll
            for (int u = 0; u < IMG2 / LU_THREADS; ++u) v[u] = __ldcg(g + t + u * LU_THREADS);
#pragma unroll
            for (int u = 0; u < IMG2 / LU_THREADS; ++u) sT2[t + u * LU_THREADS] = v[u];
            if (t < NB) {
                dinv[t] = (role == UROW) ? 1.0 : __ldcg(mine + NB * TS + t);
                if (role == LCOL) yk[t] = __ldcg(mine + PUB_DOUBLES + t);
            }
        }
        __syncthreads();
        const int ty = t >> 3, tx = t & 7;
        const int sr = (role == UROW) ? 1 : TS, sc = (role == UROW) ? TS : 1;
        double a[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) a[i][j] = sC[(ty + 16 * i) * sr + (tx + 8 * j) * sc];
        if (role == UROW) panel_solve<true>(a, sA, dinv);
        else panel_solve<false>(a, sA, dinv);
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) sC[(ty + 16 * i) * sr + (tx + 8 * j) * sc] = a[i][j];
        __syncthreads();
        store_tile(gC, ld, sC);
        if (role == UROW) {
            store_tile_all(pr, next_off + DIST_HDR + (size_t)(bj - base - 1) * NB * NB, sC);      // U[k+1,j] into every rank's slot
        } else {
            if (t < NB) {
                double sdot = 0.0;
#pragma unroll 8
                for (int q = 0; q < NB; ++q) sdot = fma(sC[t * TS + q], yk[q], sdot);
                rhs[(size_t)bi * NB + t] -= sdot;
            }
            // This CTA's work is done; it stays until the WHOLE of panel k + 1 (every U tile) has landed in this rank's mailbox, so
            // that the next launch can read it without polling.  Normally the flag is long set (the owner's U-row CTAs run beside
            // the L-panel CTAs); what is waited here is the exposed part of the exchange, accounted in stats[0].
            if (t == 0) {
                const int want = Gn + 1;
                dist_wait(pr.flags[rank] + (Gn % DIST_RING), [want](int v) { return v == want; }, pr, status, lb == lb0 + (own ? 1 : 0) ? pr.stats : nullptr);
            }
        }
    }
    if (role == UROW) {
        __threadfence_system();
        __syncthreads();
        if (t == 0) dist_arrive(pr, rem, Gn);
    }
}

static size_t dist_align(size_t v) { return (v + 255) / 256 * 256; }

static int dist_slots(const Scene* s, int slots) {
    const int nblk = s->ld / NB;
    int S = slots < 2 ? 2 : slots;
    if (S > DIST_RING / 2) S = DIST_RING / 2;
    if (S > nblk) S = nblk < 2 ? 2 : nblk;
    return S;
}


// mailbox layout: [slots | xbuf | flags | progress | xflags | own | counter, abort | stats]
struct DistLayout {
    size_t slots, xbuf, flags, iflags, progress, xflags, own, misc, stats, total, slot_doubles;
};
static DistLayout dist_layout(int ld, int S) {
    const int nblk = ld / NB;
    DistLayout L;
    L.slot_doubles = (size_t)DIST_HDR + (size_t)(nblk - 1) * NB * NB;
    size_t o = 0;
    L.slots = o;    o += dist_align((size_t)S * L.slot_doubles * 8);
    L.xbuf = o;     o += dist_align((size_t)ld * 8);
    L.flags = o;    o += dist_align((size_t)DIST_RING * 4);
    L.iflags = o;   o += dist_align((size_t)DIST_RING * 4);
    L.progress = o; o += dist_align((size_t)DIST_MAX_P * 4);
    L.xflags = o;   o += dist_align((size_t)nblk * 4);
    L.own = o;      o += dist_align((size_t)DIST_RING * 4);
    L.misc = o;     o += 256;
    L.stats = o;    o += 256;
    L.total = o;
    return L;
}

static void dist_fill_peers(Scene* s, DistComm* c, int S) {
    const DistLayout L = dist_layout(s->ld, S);
    DistPeers& p = c->peers;
    p.P = s->cs.P;
    p.rank = s->cs.rank;
    p.S = S;
    p.slot_doubles = L.slot_doubles;
    for (int q = 0; q < DIST_MAX_P; ++q) {
        char* b = (q < p.P) ? c->base[q] : nullptr;
        p.slots[q] = b ? reinterpret_cast<double*>(b + L.slots) : nullptr;
        p.xbuf[q] = b ? reinterpret_cast<double*>(b + L.xbuf) : nullptr;
        p.flags[q] = b ? reinterpret_cast<int*>(b + L.flags) : nullptr;
        p.iflags[q] = b ? reinterpret_cast<int*>(b + L.iflags) : nullptr;
        p.progress[q] = b ? reinterpret_cast<int*>(b + L.progress) : nullptr;
        p.xflags[q] = b ? reinterpret_cast<int*>(b + L.xflags) : nullptr;
    }
    p.mc_slots = c->mc ? reinterpret_cast<double*>(c->mc + L.slots) : nullptr;
    p.mc_xbuf = c->mc ? reinterpret_cast<double*>(c->mc + L.xbuf) : nullptr;
    char* me = c->base[p.rank];
    p.own = reinterpret_cast<int*>(me + L.own);
    p.counter = reinterpret_cast<int*>(me + L.misc);
    p.abort = reinterpret_cast<int*>(me + L.misc) + 1;
    p.stats = reinterpret_cast<unsigned long long*>(me + L.stats);
}

int dist_comm_create(Scene* s, int slots, unsigned char* handle64) {
    LLS_REQUIRE(s->dist == nullptr, LLS_ERR_STATE, "lls_dist_comm_create: the mailbox exists already");
    LLS_REQUIRE(s->cs.P <= DIST_MAX_P, LLS_ERR_UNSUPPORTED, "lls_dist_comm_create: at most 8 ranks (one box)");
    const int nblk = s->ld / NB;
    const int S = dist_slots(s, slots);
    (void)nblk;
    DistComm* c = new (std::nothrow) DistComm();
    LLS_REQUIRE(c != nullptr, LLS_ERR_ARG, "lls_dist_comm_create: out of host memory");
    const DistLayout L = dist_layout(s->ld, S);
    c->bytes = L.total;
    void* p = nullptr;
    if (cudaMalloc(&p, L.total) != cudaSuccess || cudaMemset(p, 0, L.total) != cudaSuccess) {
        delete c;
        set_last_error(__FILE__, __LINE__, "lls_dist_comm_create: cudaMalloc of the mailbox failed");
        return LLS_ERR_CUDA;
    }
    c->base[s->cs.rank] = static_cast<char*>(p);
    if (handle64 != nullptr) {
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        cudaIpcMemHandle_t h;
        if (s->cs.P > 1) {
            if (cudaIpcGetMemHandle(&h, p) != cudaSuccess) {
                cudaFree(p);
                delete c;
                set_last_error(__FILE__, __LINE__, "lls_dist_comm_create: cudaIpcGetMemHandle failed");
                return LLS_ERR_CUDA;
            }
            memcpy(handle64, &h, 64);
        } else {
            memset(handle64, 0, 64);
        }
    }
    s->dist = c;
    dist_fill_peers(s, c, S);
    c->connected = (s->cs.P == 1);
    static bool attr = false;
    if (!attr) {
        LLS_CUDA_TRY(cudaFuncSetAttribute(lu_dist_owner_p2p_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * TILE_BYTES));
        LLS_CUDA_TRY(cudaFuncSetAttribute(lu_dist_panel_p2p_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * TILE_BYTES));
        LLS_CUDA_TRY(cudaFuncSetAttribute(lu_dist_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * TILE_BYTES));
        LLS_CUDA_TRY(cudaFuncSetAttribute(lu_dist_step_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * TILE_BYTES));
        LLS_CUDA_TRY(cudaFuncSetAttribute(lu_dist_step_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * TILE_BYTES));
        LLS_CUDA_TRY(cudaFuncSetAttribute(lu_dist_step_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * TILE_BYTES));
        attr = true;
    }
    return LLS_OK;
}

size_t dist_comm_bytes(const Scene* s, int slots) { return dist_layout(s->ld, dist_slots(s, slots)).total; }

// Mailboxes that belong to the caller: `bases` = the mailbox of every rank as mapped into this process (symmetric memory: same
// size on every rank, zero-filled, bases[rank] = the own one), `mc` = the multicast mapping of all of them (or null).
int dist_comm_attach(Scene* s, int slots, void* const* bases, void* mc) {
    LLS_REQUIRE(s->dist == nullptr, LLS_ERR_STATE, "lls_dist_comm_attach: the mailbox exists already");
    LLS_REQUIRE(s->cs.P <= DIST_MAX_P, LLS_ERR_UNSUPPORTED, "lls_dist_comm_attach: at most 8 ranks (one box)");
    DistComm* c = new (std::nothrow) DistComm();
    LLS_REQUIRE(c != nullptr, LLS_ERR_ARG, "lls_dist_comm_attach: out of host memory");
    const int S = dist_slots(s, slots);
    c->bytes = dist_layout(s->ld, S).total;
    c->external = true;
    for (int q = 0; q < s->cs.P; ++q) {
        if (bases[q] == nullptr) {
            delete c;
            set_last_error(__FILE__, __LINE__, "lls_dist_comm_attach: null mailbox pointer");
            return LLS_ERR_ARG;
        }
        c->base[q] = static_cast<char*>(bases[q]);
    }
    c->mc = static_cast<char*>(mc);
    s->dist = c;
    dist_fill_peers(s, c, S);
    c->connected = true;
    LLS_CUDA_TRY(cudaFuncSetAttribute(lu_dist_owner_p2p_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * TILE_BYTES));
    LLS_CUDA_TRY(cudaFuncSetAttribute(lu_dist_panel_p2p_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * TILE_BYTES));
    LLS_CUDA_TRY(cudaFuncSetAttribute(lu_dist_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * TILE_BYTES));
    LLS_CUDA_TRY(cudaFuncSetAttribute(lu_dist_step_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * TILE_BYTES));
    LLS_CUDA_TRY(cudaFuncSetAttribute(lu_dist_step_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * TILE_BYTES));
    LLS_CUDA_TRY(cudaFuncSetAttribute(lu_dist_step_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * TILE_BYTES));
    return LLS_OK;
}

int dist_comm_multicast(const Scene* s) {
    const DistComm* c = static_cast<const DistComm*>(s->dist);
    return (c != nullptr && c->mc != nullptr) ? 1 : 0;
}

int dist_comm_connect(Scene* s, const unsigned char* handles) {
    DistComm* c = static_cast<DistComm*>(s->dist);
    LLS_REQUIRE(c != nullptr, LLS_ERR_STATE, "lls_dist_comm_connect: call lls_dist_comm_create first");
    for (int q = 0; q < s->cs.P; ++q) {
        if (q == s->cs.rank || c->opened[q]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)q * 64, 64);
        void* p = nullptr;
        if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            set_last_error(__FILE__, __LINE__, "lls_dist_comm_connect: cudaIpcOpenMemHandle failed (no peer access between the GPUs?)");
            cudaGetLastError();
            return LLS_ERR_CUDA;
        }
        c->base[q] = static_cast<char*>(p);
        c->opened[q] = true;
    }
    dist_fill_peers(s, c, c->peers.S);
    c->connected = true;
    return LLS_OK;
}

int dist_comm_destroy(Scene* s) {
    DistComm* c = static_cast<DistComm*>(s->dist);
    if (c == nullptr) return LLS_OK;
    for (int q = 0; q < DIST_MAX_P; ++q)
        if (c->opened[q]) cudaIpcCloseMemHandle(c->base[q]);
    if (!c->external && c->base[s->cs.rank]) cudaFree(c->base[s->cs.rank]);
    delete c;
    s->dist = nullptr;
    return LLS_OK;
}

int dist_comm_stats(Scene* s, double* out4, cudaStream_t st) {
    DistComm* c = static_cast<DistComm*>(s->dist);
    LLS_REQUIRE(c != nullptr, LLS_ERR_STATE, "lls_dist_comm_stats: no mailbox");
    unsigned long long h[4] = {0, 0, 0, 0};
    LLS_CUDA_TRY(cudaMemcpyAsync(h, c->peers.stats, sizeof h, cudaMemcpyDeviceToHost, st));
    LLS_CUDA_TRY(cudaStreamSynchronize(st));
    out4[0] = (double)h[0] * 1e-6;       // ms block 0 of the panel kernels waited for a panel to arrive
    out4[1] = (double)h[1] * 1e-6;       // ms the owner kernels waited for a slot to be free
    out4[2] = (double)c->gstep;
    out4[3] = (double)c->solves;
    LLS_CUDA_TRY(cudaMemsetAsync(c->peers.stats, 0, 32, st));
    return LLS_OK;
}

// One factorisation with the right-hand side carried along, then the back substitution: x (ld doubles, replicated) on return.
// Everything is enqueued here; between two block steps there is neither a host synchronisation nor a collective call.
int lu_dist_factor_solve_p2p(Scene* s, double* rhs, double* x, cudaStream_t st) {
    DistComm* c = static_cast<DistComm*>(s->dist);
    LLS_REQUIRE(c != nullptr && c->connected, LLS_ERR_STATE, "lls_dist_lu_solve: mailboxes not connected (lls_dist_comm_create / _connect)");
    const int nblk = s->ld / NB, P = s->cs.P, rank = s->cs.rank, S = c->peers.S;
    const DistPeers& pr = c->peers;
    const long long G0 = c->gstep;
    LLS_REQUIRE(G0 + nblk < 2000000000LL, LLS_ERR_STATE, "lls_dist_lu_solve: global step counter exhausted; recreate the scene");
    auto owner = [&](int k) {
        const int G = (int)(G0 + k), ncols = nblk - 1 - k;
        lu_dist_owner_p2p_kernel<<<1 + ncols, LU_THREADS, 2 * TILE_BYTES, st>>>(s->M, s->ld, k, k / P, ncols, rhs, pr, G % S, G, s->uinv,
                                                                              s->scalars, s->status);
        count_launch();
    };
    auto update = [&](int k, int lb0, int nrows) {
        const int ncols = nblk - 1 - k, G = (int)(G0 + k);
        if (nrows <= 0 || ncols <= 0) return;
        lu_dist_update_kernel<<<nrows * ncols, LU_THREADS, 2 * TILE_BYTES, st>>>(s->M, s->ld, k, lb0, ncols,
                                                                               pr.slots[rank] + (size_t)(G % S) * pr.slot_doubles + DIST_HDR);
        count_launch();
    };
    static int fused_steps = -1;        // LLS_PARTITIONED_FUSED_STEPS=0: separate owner / panel / update launches with launch-order look-ahead
    if (fused_steps < 0) {
        const char* env = getenv("LLS_PARTITIONED_FUSED_STEPS");
        fused_steps = (env != nullptr && env[0] == '0') ? 0 : 1;
    }
    static int pair_min_rem = -1;       // LLS_PARTITIONED_PAIR_MIN_REM: rank-128 visits while the trailing matrix has at least this many block rows (0 = never); default 64 sqrt(P)
    if (pair_min_rem < 0) {
        const char* env = getenv("LLS_PARTITIONED_PAIR_MIN_REM");
        pair_min_rem = env != nullptr ? atoi(env) : -2;
    }
    const int pair_min = (pair_min_rem == -2) ? dist_default_pair_min(P, c->mc != nullptr) : pair_min_rem;
    if (fused_steps) {
        const int nlaunch = dist_schedule(nblk, P, rank, s->n_loc_blocks, pair_min, nullptr, 0);
        std::vector<DistLaunch> plan((size_t)nlaunch);
        dist_schedule(nblk, P, rank, s->n_loc_blocks, pair_min, plan.data(), nlaunch);
        for (const DistLaunch& L : plan) {
            const int Gn = (int)(G0 + L.base);
            if (L.k < 0) {
                lu_dist_step_kernel<true, 1><<<(unsigned)L.grid, LU_THREADS, 2 * TILE_BYTES, st>>>(s->M, s->ld, L.k, nblk, L.lb0, L.nloc, rhs, pr, Gn, s->uinv,
                                                                                                  s->scalars, s->status, L.thin);
            } else if (L.passes == 1) {
                lu_dist_step_kernel<false, 1><<<(unsigned)L.grid, LU_THREADS, 2 * TILE_BYTES, st>>>(s->M, s->ld, L.k, nblk, L.lb0, L.nloc, rhs, pr, Gn, s->uinv,
                                                                                                   s->scalars, s->status, L.thin);
            } else {
                lu_dist_step_kernel<false, 2><<<(unsigned)L.grid, LU_THREADS, 2 * TILE_BYTES, st>>>(s->M, s->ld, L.k, nblk, L.lb0, L.nloc, rhs, pr, Gn, s->uinv,
                                                                                                   s->scalars, s->status, L.thin);
            }
            count_launch();
        }
    } else {
    if (0 % P == rank) owner(0);
    for (int k = 0; k < nblk; ++k) {
        const int G = (int)(G0 + k);
        int lb0 = (k + 1 - rank + P - 1) / P;          // first local block whose global index lb * P + rank exceeds k
        if (lb0 < 0) lb0 = 0;
        const int nloc = s->n_loc_blocks - lb0 > 0 ? s->n_loc_blocks - lb0 : 0;
        lu_dist_panel_p2p_kernel<<<nloc > 0 ? nloc : 1, LU_THREADS, 2 * TILE_BYTES, st>>>(s->M, s->ld, k, lb0, nloc, pr, G % S, G, rhs, s->status);
        count_launch();
        if (k + 1 < nblk && (k + 1) % P == rank) {
            // look-ahead: this rank owns block row k + 1 (its local block lb0): bring that row up to date, factor it and push
            // panel k + 1 to everybody before the rest of the local rows take panel k
            update(k, lb0, 1);
            owner(k + 1);
            update(k, lb0 + 1, nloc - 1);
        } else {
            update(k, lb0, nloc);
        }
    }
    }
    c->gstep = G0 + nblk;
    const int token = ++c->solves;
    lu_dist_backsolve_p2p_kernel<<<s->n_loc_blocks, NB, 0, st>>>(s->M, s->ld, nblk, s->n_loc_blocks, s->uinv, rhs, pr, token, (int)c->gstep, s->status);
    count_launch();
    LLS_CUDA_TRY(cudaMemcpyAsync(x, pr.xbuf[rank], sizeof(double) * (size_t)s->ld, cudaMemcpyDeviceToDevice, st));
    LLS_CUDA_TRY(cudaGetLastError());
    return LLS_OK;
}

// ---- the schedule as data, for the CPU tests: no GPU is touched ---------------------------------------------------------------------
int dist_debug_schedule(int nblk, int P, int rank, int pair_min, int* out, int max_launches) {
    if (nblk < 1 || P < 1 || rank < 0 || rank >= P || nblk % P != 0) return -1;
    const int nloc = nblk / P;
    if (pair_min < 0) pair_min = dist_default_pair_min(P, pair_min == -1);        // -1: default with multicast, -2: default without
    const int n = dist_schedule(nblk, P, rank, nloc, pair_min, nullptr, 0);
    if (out == nullptr) return n;
    std::vector<DistLaunch> plan((size_t)n);
    dist_schedule(nblk, P, rank, nloc, pair_min, plan.data(), n);
    for (int e = 0; e < n && e < max_launches; ++e) {
        const DistLaunch& L = plan[(size_t)e];
        int* o = out + 9 * e;
        o[0] = L.k; o[1] = L.passes; o[2] = L.thin; o[3] = L.base; o[4] = L.rem; o[5] = L.lb0; o[6] = L.nloc; o[7] = L.own;
        o[8] = (int)L.grid;
    }
    return n;
}
int dist_debug_role(int b, int own, int rem, int base, int lb0, int nloc, int thin, int* lb, int* bj) {
    return dist_role(b, own, rem, base, lb0, nloc, thin, *lb, *bj);
}

}  // namespace lls
