// Internal declarations shared by the CUDA translation units of liblls (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <type_traits>
#include "../../include/lls.h"

#define LLS_CUDA_TRY(expr)                                                                    \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            lls::set_last_error(__FILE__, __LINE__, cudaGetErrorString(_e));                  \
            return LLS_ERR_CUDA;                                                              \
        }                                                                                     \
    } while (0)

#define LLS_REQUIRE(cond, code, msg)                                                          \
    do {                                                                                      \
        if (!(cond)) {                                                                        \
            lls::set_last_error(__FILE__, __LINE__, msg);                                     \
            return code;                                                                      \
        }                                                                                     \
    } while (0)

#define LLS_TRY(expr)                                                                         \
    do {                                                                                      \
        int _rc = (expr);                                                                     \
        if (_rc != LLS_OK) return _rc;                                                        \
    } while (0)

namespace lls {

void set_last_error(const char* file, int line, const char* msg);
extern unsigned long long g_launches;  // kernels launched by this library since load (bench.py's gpu_launches)
inline void count_launch(int k = 1) { g_launches += (unsigned long long)k; }

constexpr int LD_ALIGN = 64;   // leading dimension granularity (elements); equals the LU block size
constexpr int LU_NB = 64;      // LU block size
constexpr size_t LU_PUB_BYTES = 2 * (size_t)(2 * LU_NB * (LU_NB + 4) + LU_NB) * 8;   // two parities of the published diagonal-block image (lu.cu)

// Per-control-point scratch ("row state"), field-major with the scene's ld.  Produced by the
// flow-state and residual kernels, consumed by the assembly kernels.
enum WField {
    W_VINF = 0,      // 3  v_inf                         (scene.py:571)
    W_VIR = 3,       // 3  v_inf_and_rot                 (scene.py:572)
    W_VINF_MAG = 6,  //    |v_inf|                       (scene.py:585)
    W_VIR_MAG = 7,   //    |v_inf_and_rot|               (scene.py:587)
    W_VINF_IP = 8,   //    |P v_inf|                     (scene.py:639)
    W_CLA0 = 9,      //    CLa of the linear pre-pass    (scene.py:659)
    W_CL0 = 10,      //    CL of the linear pre-pass incl. sweep correction (scene.py:660, 666)
    W_RHS = 11,      //    right-hand side / Newton step (scene.py:818, 896)
    W_DIAG = 12,     //    diagonal addend               (scene.py:824, 893)
    W_G = 13,        // 3  row vector g_i of the assembly kernels
    W_WS = 16,       // 3  scaled w_i (2 gamma_i/|w_i|) w_i   (scene.py:867)
    W_E = 19,        // 3  CLRe row vector (scene.py:873)
    W_X = 22,        //    solution of the triangular solves
    W_UINF = 23,     // 3  v_inf / |v_inf|               (scene.py:586)
    W_NF = 26
};

// Batches: `ncases` independent flight states of ONE geometry (alpha / beta / rate / control sweeps) are solved by the same
// launches with the case index in blockIdx.z.  Geometry tables are shared; every per-case buffer is the single-scene
// buffer repeated with these element strides (all zero for a single scene).  `active` masks cases out of the Newton phase.
struct CaseStrides {
    size_t out = 0, V = 0, M = 0, ac = 0, seg = 0;   // doubles
    size_t tab = 0;                                   // doubles between the per-case copies of the double table (0: one table shared by every case)
    size_t work = 0;                                  // BYTES between the work regions of consecutive cases
    const int* active = nullptr;
    // row partition of ONE large scene over the GPUs of a box: 64-row blocks dealt round-robin, rank r owns global blocks
    // r, r + P, r + 2P ... stored consecutively (local block lb <-> global block lb * P + rank).  P = 1: everything local.
    int P = 1, rank = 0;
};
__host__ __device__ __forceinline__ int global_row(const CaseStrides& cs, int local_row) {
    return ((local_row >> 6) * cs.P + cs.rank) * 64 + (local_row & 63);
}
__host__ __device__ __forceinline__ bool owns_row(const CaseStrides& cs, int global_row_index) {
    return ((global_row_index >> 6) % cs.P) == cs.rank;
}
template <class T>
__device__ __forceinline__ T* case_ptr(T* p, size_t stride) { return p + (size_t)blockIdx.z * stride; }
template <class T>
__device__ __forceinline__ T* case_work_ptr(T* p, size_t stride_bytes) {
    return reinterpret_cast<T*>(reinterpret_cast<char*>(const_cast<typename std::remove_const<T>::type*>(p)) + (size_t)blockIdx.z * stride_bytes);
}
#define LLS_CASE_INACTIVE(cs) ((cs).active != nullptr && (cs).active[blockIdx.z] == 0)

struct Scene {
    int n = 0, ld = 0, n_ac = 0, n_seg = 0, n_af = 0, flags = 0;
    int ncases = 1;             // > 1 after lls_bind_batch
    int n_loc_blocks = 0;       // 64-row blocks of V / M stored on this rank (all of them unless lls_set_partition was called)
    CaseStrides cs;             // strides of the bound batch (active == nullptr)
    int* ctl = nullptr;         // batch control block (device): active[ncases], iterations[ncases], n_active
    int* h_ctl = nullptr;       // pinned host mirror of n_active
    const double* tab = nullptr;
    const int32_t* itab = nullptr;
    const double* af = nullptr;
    const double* ac = nullptr;
    double* out = nullptr;
    double* V = nullptr;
    double* M = nullptr;
    // carved from the caller's work buffer
    double* W = nullptr;        // W_NF * ld
    double* inv = nullptr;      // LU_PUB_BYTES: factored diagonal block images published by the LU's diag CTA (double-buffered)
    double* uinv = nullptr;     // ld * LU_NB: inverse of every diagonal U block, left by the factorisation for the back-substitution chain
    int* lu_flags = nullptr;    // 6 * (ld / LU_NB + 1) ints ([3 (nblk + 1)] = SM announced by the diag CTA, [4 (nblk + 1) + 2 k] role claims of step k): [k] == lu_epoch once diagonal block k is published; [nblk + 1 + k] once x_k is; [2 (nblk + 1) + k] work counter of block step k
    int lu_epoch = 0;           // factorisation counter (host)
    double* partial = nullptr;  // reduction scratch (>= 4096 doubles)
    double* scalars = nullptr;  // [0] ||R||^2  [1] min |pivot| ...   (device)
    int* status = nullptr;      // device status word
    double* h_pinned = nullptr; // pinned host mirror for scalar read-back (8 doubles)
    bool timing = false;
    double ms[6] = {0, 0, 0, 0, 0, 0};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    void* dist = nullptr;       // DistComm of the row-partitioned LU over peer memory (lu_dist.cu), created by lls_dist_comm_create
};

inline CaseStrides strides(const Scene* s, bool masked) {
    CaseStrides c = s->cs;
    c.active = (masked && s->ncases > 1) ? s->ctl : nullptr;
    return c;
}
inline const double* tabf(const Scene* s, int f) { return s->tab + (size_t)f * s->ld; }
inline const int32_t* itabf(const Scene* s, int f) { return s->itab + (size_t)f * s->ld; }
inline double* outf(const Scene* s, int f) { return s->out + (size_t)f * s->ld; }
inline double* wf(const Scene* s, int f) { return s->W + (size_t)f * s->ld; }

// kernels' host launchers (each returns LLS_OK or an error code)
int launch_flow_state(Scene* s, cudaStream_t st);
int launch_influence(Scene* s, double impingement_threshold, cudaStream_t st);
int launch_assemble_linear(Scene* s, cudaStream_t st);
int launch_residual(Scene* s, int vel_only, cudaStream_t st, bool masked = false);  // matvec + section lift + R, ||R||^2 -> scalars[0]
int launch_newton_control(Scene* s, int k, double convergence, cudaStream_t st);      // batch: active mask / iteration counters of iteration k
int launch_assemble_jacobian(Scene* s, cudaStream_t st);
int launch_newton_update(Scene* s, double relaxation, cudaStream_t st);
int launch_integrate(Scene* s, double* d_seg_fm, cudaStream_t st);
int launch_scale_gamma_pro(Scene* s, cudaStream_t st);
int lu_factor_solve(Scene* s, double* d_mat, double* d_rhs, double* d_x, cudaStream_t st, bool masked = false);  // rhs may be null
int lu_solve_only(Scene* s, const double* d_mat, double* d_rhs_inout, cudaStream_t st);
// row-partitioned LU, one block step at a time (the host owns the broadcasts between the steps)
size_t lu_dist_step_doubles(const Scene* s, int k);
int lu_dist_begin(Scene* s, cudaStream_t st);
int lu_dist_owner(Scene* s, int k, double* rhs, double* send, cudaStream_t st);
int lu_dist_update(Scene* s, int k, double* rhs, const double* recv, cudaStream_t st);
int lu_dist_back(Scene* s, int k, const double* y, double* x, cudaStream_t st);
// row-partitioned LU with the exchange inside the kernels (mailboxes in peer memory, CUDA IPC between the processes)
int dist_comm_create(Scene* s, int slots, unsigned char* handle64);
int dist_comm_connect(Scene* s, const unsigned char* handles);
size_t dist_comm_bytes(const Scene* s, int slots);
int dist_comm_attach(Scene* s, int slots, void* const* bases, void* mc);
int dist_comm_multicast(const Scene* s);
int dist_debug_schedule(int nblk, int P, int rank, int pair_min, int* out, int max_launches);
int dist_debug_role(int b, int own, int rem, int base, int lb0, int nloc, int thin, int* lb, int* bj);
int dist_comm_destroy(Scene* s);
int dist_comm_stats(Scene* s, double* out4, cudaStream_t st);
int lu_dist_factor_solve_p2p(Scene* s, double* rhs, double* x, cudaStream_t st);

}  // namespace lls

struct lls_scene : lls::Scene {};
