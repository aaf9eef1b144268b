// Handle life cycle, buffer binding and the solver drivers of liblls (the C ABI in include/lls.h).
#include "lls_internal.h"
#include <cstring>
#include <new>

namespace lls {

static thread_local char g_err[512] = "";
unsigned long long g_launches = 0;
void set_last_error(const char* file, int line, const char* msg) {
    snprintf(g_err, sizeof g_err, "%s:%d: %s", file, line, msg);
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static size_t work_bytes_for(int ld) {
    size_t b = 0;
    b += align_up((size_t)W_NF * ld * 8, 256);
    b += align_up((size_t)LU_PUB_BYTES, 256);               // published diagonal LU block images (x2)
    b += align_up((size_t)ld * LU_NB * 8, 256);             // inverse diagonal U blocks
    b += align_up((size_t)6 * (ld / LU_NB + 1) * 4, 256);   // LU flags (factorisation, back-substitution), work counters, diag-SM word, role claims
    b += align_up((size_t)ld * 8, 256);  // partial
    b += 256;                            // scalars
    b += 256;                            // status
    return b;
}

struct PhaseTimer {
    Scene* s;
    cudaStream_t st;
    int slot;
    PhaseTimer(Scene* s_, cudaStream_t st_, int slot_) : s(s_), st(st_), slot(slot_) {
        if (s->timing) cudaEventRecord(s->ev[0], st);
    }
    ~PhaseTimer() {
        if (s->timing) {
            cudaEventRecord(s->ev[1], st);
            cudaEventSynchronize(s->ev[1]);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, s->ev[0], s->ev[1]);
            s->ms[slot] += ms;
        }
    }
};

static int read_scalars(Scene* s, cudaStream_t st, int count) {
    LLS_CUDA_TRY(cudaMemcpyAsync(s->h_pinned, s->scalars, sizeof(double) * count, cudaMemcpyDeviceToHost, st));
    LLS_CUDA_TRY(cudaStreamSynchronize(st));
    return LLS_OK;
}

}  // namespace lls

using namespace lls;

extern "C" const char* lls_last_error(void) { return lls::g_err; }
extern "C" int lls_abi_version(void) { return 3; }
extern "C" unsigned long long lls_launch_count(void) { return lls::g_launches; }

extern "C" int lls_query_sizes(int n, int n_segments, lls_sizes* out) {
    LLS_REQUIRE(n > 0 && out != nullptr, LLS_ERR_ARG, "lls_query_sizes: bad arguments");
    (void)n_segments;
    const size_t ld = align_up((size_t)n, LD_ALIGN);
    out->ld = ld;
    out->tab_bytes = (size_t)LLS_NF * ld * 8;
    out->itab_bytes = (size_t)LLS_NI * ld * 4;
    out->out_bytes = (size_t)LLS_NO * ld * 8;
    out->vji_bytes = (size_t)n * 3 * ld * 8;
    out->mat_bytes = ld * ld * 8;
    out->work_bytes = work_bytes_for((int)ld);
    return LLS_OK;
}

extern "C" int lls_query_sizes_partitioned(int n, int n_segments, int world, lls_sizes* out) {
    LLS_REQUIRE(n > 0 && out != nullptr && world >= 1, LLS_ERR_ARG, "lls_query_sizes_partitioned: bad arguments");
    (void)n_segments;
    const size_t ld = align_up((size_t)n, (size_t)LD_ALIGN * world);     // every rank owns the same number of 64-row blocks
    const size_t local_rows = ld / world;
    out->ld = ld;
    out->tab_bytes = (size_t)LLS_NF * ld * 8;
    out->itab_bytes = (size_t)LLS_NI * ld * 4;
    out->out_bytes = (size_t)LLS_NO * ld * 8;
    out->vji_bytes = local_rows * 3 * ld * 8;
    out->mat_bytes = local_rows * ld * 8;
    out->work_bytes = work_bytes_for((int)ld);
    return LLS_OK;
}

extern "C" int lls_set_partition(lls_scene* h, int world, int rank) {
    LLS_REQUIRE(h != nullptr && world >= 1 && rank >= 0 && rank < world, LLS_ERR_ARG, "lls_set_partition: bad arguments");
    LLS_REQUIRE(h->tab == nullptr, LLS_ERR_STATE, "lls_set_partition: call it before lls_bind");
    h->ld = (int)align_up((size_t)h->n, (size_t)LD_ALIGN * world);
    h->n_loc_blocks = h->ld / LU_NB / world;
    h->cs.P = world;
    h->cs.rank = rank;
    return LLS_OK;
}

extern "C" int lls_create(lls_scene** out, int n, int n_aircraft, int n_segments, int n_airfoils, int flags) {
    LLS_REQUIRE(out != nullptr && n > 0 && n_aircraft > 0 && n_segments > 0 && n_airfoils > 0, LLS_ERR_ARG, "lls_create: bad arguments");
    const bool total = flags & LLS_FLAG_TOTAL_VELOCITY, pro = flags & LLS_FLAG_MATCH_MACHUP_PRO;
    LLS_REQUIRE(total || pro, LLS_ERR_UNSUPPORTED,
                "use_total_velocity=False without match_machup_pro is not supported (the reference itself fails there: scene.py:959)");
    lls_scene* h = new (std::nothrow) lls_scene();
    LLS_REQUIRE(h != nullptr, LLS_ERR_ARG, "lls_create: out of host memory");
    h->n = n;
    h->ld = (int)align_up((size_t)n, LD_ALIGN);
    h->n_loc_blocks = h->ld / LU_NB;
    h->n_ac = n_aircraft;
    h->n_seg = n_segments;
    h->n_af = n_airfoils;
    h->flags = flags;
    if (cudaMallocHost(&h->h_pinned, 64 * sizeof(double)) != cudaSuccess || cudaMallocHost(&h->h_ctl, 16 * sizeof(int)) != cudaSuccess ||
        cudaEventCreate(&h->ev[0]) != cudaSuccess ||
        cudaEventCreate(&h->ev[1]) != cudaSuccess) {
        set_last_error(__FILE__, __LINE__, "lls_create: CUDA resource creation failed (is a GPU visible?)");
        delete h;
        return LLS_ERR_CUDA;
    }
    *out = h;
    return LLS_OK;
}

extern "C" int lls_destroy(lls_scene* h) {
    if (h == nullptr) return LLS_OK;
    dist_comm_destroy(h);
    if (h->h_pinned) cudaFreeHost(h->h_pinned);
    if (h->h_ctl) cudaFreeHost(h->h_ctl);
    if (h->ev[0]) cudaEventDestroy(h->ev[0]);
    if (h->ev[1]) cudaEventDestroy(h->ev[1]);
    delete h;
    return LLS_OK;
}

extern "C" int lls_bind(lls_scene* h, const double* d_tab, const int32_t* d_itab, const double* d_airfoils, double* d_out,
                        double* d_vji, double* d_mat, void* d_work, size_t work_bytes) {
    LLS_REQUIRE(h != nullptr, LLS_ERR_ARG, "lls_bind: null handle");
    LLS_REQUIRE(d_tab && d_itab && d_airfoils && d_out && d_vji && d_mat && d_work, LLS_ERR_ARG, "lls_bind: null buffer");
    LLS_REQUIRE(work_bytes >= work_bytes_for(h->ld), LLS_ERR_ARG, "lls_bind: work buffer too small");
    LLS_REQUIRE(((uintptr_t)d_vji % 16 == 0) && ((uintptr_t)d_mat % 16 == 0) && ((uintptr_t)d_out % 16 == 0) && ((uintptr_t)d_work % 256 == 0),
                LLS_ERR_ARG, "lls_bind: buffers must be 16-byte (work: 256-byte) aligned");
    h->tab = d_tab;
    h->itab = d_itab;
    h->af = d_airfoils;
    h->out = d_out;
    h->V = d_vji;
    h->M = d_mat;
    char* p = (char*)d_work;
    h->W = (double*)p;
    p += align_up((size_t)W_NF * h->ld * 8, 256);
    h->inv = (double*)p;
    p += align_up((size_t)LU_PUB_BYTES, 256);
    h->uinv = (double*)p;
    p += align_up((size_t)h->ld * LU_NB * 8, 256);
    h->lu_flags = (int*)p;
    p += align_up((size_t)6 * (h->ld / LU_NB + 1) * 4, 256);
    h->partial = (double*)p;
    p += align_up((size_t)h->ld * 8, 256);
    h->scalars = (double*)p;
    p += 256;
    h->status = (int*)p;
    // the LU's publish flags compare against a per-handle epoch that starts at 0: clear them once here
    LLS_CUDA_TRY(cudaMemset(h->lu_flags, 0, (size_t)6 * (h->ld / LU_NB + 1) * 4));
    h->lu_epoch = 0;
    LLS_CUDA_TRY(cudaMemset(h->inv, 0, LU_PUB_BYTES));       // includes the accumulator / ticket of the partitioned back substitution
    h->ncases = 1;
    {
        const int P = h->cs.P, rank = h->cs.rank;
        h->cs = CaseStrides();
        h->cs.P = P;
        h->cs.rank = rank;
    }
    h->ctl = nullptr;
    return LLS_OK;
}

extern "C" int lls_bind_batch(lls_scene* h, int ncases, const double* d_tab, const int32_t* d_itab, const double* d_airfoils,
                              double* d_out, double* d_vji, double* d_mat, void* d_work, size_t work_bytes_per_case, int* d_ctl) {
    LLS_REQUIRE(h != nullptr && ncases >= 1 && ncases <= 65535, LLS_ERR_ARG, "lls_bind_batch: 1 <= ncases <= 65535 (case index is blockIdx.z)");
    LLS_REQUIRE(h->cs.P == 1, LLS_ERR_STATE, "lls_bind_batch: batches and row partitions do not combine");
    LLS_REQUIRE(d_ctl != nullptr, LLS_ERR_ARG, "lls_bind_batch: null control block");
    LLS_REQUIRE(work_bytes_per_case % 256 == 0, LLS_ERR_ARG, "lls_bind_batch: work_bytes_per_case must be a multiple of 256");
    LLS_TRY(lls_bind(h, d_tab, d_itab, d_airfoils, d_out, d_vji, d_mat, d_work, work_bytes_per_case));   // case 0 carves the layout
    h->ncases = ncases;
    h->cs.out = (size_t)LLS_NO * h->ld;
    h->cs.V = (size_t)h->n * 3 * h->ld;
    h->cs.M = (size_t)h->ld * h->ld;
    h->cs.ac = (size_t)h->n_ac * LLS_AC_STRIDE;
    h->cs.seg = (size_t)h->n_seg * 12;
    h->cs.work = work_bytes_per_case;
    h->ctl = d_ctl;
    // LU publish flags of every case start below any epoch
    LLS_CUDA_TRY(cudaMemset2D(h->lu_flags, work_bytes_per_case, 0, (size_t)6 * (h->ld / LU_NB + 1) * 4, (size_t)ncases));
    LLS_CUDA_TRY(cudaMemset(d_ctl, 0, sizeof(int) * (2 * (size_t)ncases + 4)));
    return LLS_OK;
}

extern "C" int lls_set_case_tables(lls_scene* h, const double* d_tab, size_t case_stride_doubles) {
    LLS_REQUIRE(h != nullptr && d_tab != nullptr, LLS_ERR_ARG, "lls_set_case_tables: null argument");
    LLS_REQUIRE(h->ctl != nullptr, LLS_ERR_STATE, "lls_set_case_tables: no batch bound (lls_bind_batch)");
    LLS_REQUIRE(case_stride_doubles == 0 || case_stride_doubles >= (size_t)LLS_NF * h->ld, LLS_ERR_ARG,
                "lls_set_case_tables: stride shorter than one table");
    h->tab = d_tab;
    h->cs.tab = case_stride_doubles;
    return LLS_OK;
}

extern "C" int lls_set_state(lls_scene* h, const double* d_ac_state) {
    LLS_REQUIRE(h != nullptr && d_ac_state != nullptr, LLS_ERR_ARG, "lls_set_state: null argument");
    h->ac = d_ac_state;
    return LLS_OK;
}

#define LLS_CHECK_BOUND(h)                                                                     \
    LLS_REQUIRE((h) != nullptr && (h)->tab != nullptr && (h)->ac != nullptr, LLS_ERR_STATE,     \
                "scene is not bound (call lls_bind and lls_set_state first)")

extern "C" int lls_flow_state(lls_scene* h, void* stream) {
    LLS_CHECK_BOUND(h);
    return launch_flow_state(h, (cudaStream_t)stream);
}

extern "C" int lls_build_influence(lls_scene* h, double impingement_threshold, void* stream) {
    LLS_CHECK_BOUND(h);
    PhaseTimer t(h, (cudaStream_t)stream, 0);
    return launch_influence(h, impingement_threshold, (cudaStream_t)stream);
}

extern "C" int lls_solve_linear(lls_scene* h, void* stream) {
    LLS_CHECK_BOUND(h);
    cudaStream_t st = (cudaStream_t)stream;
    {
        PhaseTimer t(h, st, 1);
        LLS_TRY(launch_assemble_linear(h, st));
    }
    {
        PhaseTimer t(h, st, 2);
        LLS_TRY(lu_factor_solve(h, h->M, wf(h, W_RHS), wf(h, W_X), st));
    }
    return launch_newton_update(h, -1.0, st);  // gamma = x
}

extern "C" int lls_solve_nonlinear(lls_scene* h, double convergence, double relaxation, int max_iterations, void* stream,
                                   int* iterations, double* final_error, int* status, double* err_hist) {
    LLS_CHECK_BOUND(h);
    LLS_REQUIRE(iterations && final_error && status, LLS_ERR_ARG, "lls_solve_nonlinear: null output");
    LLS_REQUIRE(h->ncases == 1, LLS_ERR_STATE, "lls_solve_nonlinear: a batch is bound, use lls_solve_nonlinear_batch");
    cudaStream_t st = (cudaStream_t)stream;
    int it = 0;
    double error = 100.0;                                                        // scene.py:851-852
    int flags_out = 0;
    while (error > convergence) {                                                // scene.py:853
        ++it;
        {
            PhaseTimer t(h, st, 4);
            LLS_TRY(launch_residual(h, 0, st));                                  // scene.py:857
        }
        LLS_TRY(read_scalars(h, st, 1));
        error = sqrt(h->h_pinned[0]);                                            // scene.py:858
        if (err_hist) err_hist[it - 1] = error;
        if (!(error == error)) {  // NaN: the reference's while-test fails on NaN and falls out of the loop
            flags_out |= LLS_STATUS_NAN;
        }
        {
            PhaseTimer t(h, st, 1);
            LLS_TRY(launch_assemble_jacobian(h, st));                            // scene.py:862-893
        }
        {
            PhaseTimer t(h, st, 2);
            LLS_TRY(lu_factor_solve(h, h->M, wf(h, W_RHS), wf(h, W_X), st));     // scene.py:896
        }
        LLS_TRY(launch_newton_update(h, relaxation, st));                        // scene.py:899
        if (it >= max_iterations) {                                              // scene.py:905-908
            flags_out |= LLS_STATUS_NOT_CONVERGED;
            break;
        }
    }
    {
        PhaseTimer t(h, st, 4);
        LLS_TRY(launch_residual(h, 0, st));                                      // scene.py:906-907, 912-913
    }
    LLS_TRY(read_scalars(h, st, 1));
    *final_error = sqrt(h->h_pinned[0]);
    *iterations = it;
    int dev_status = 0;
    LLS_CUDA_TRY(cudaMemcpyAsync(&dev_status, h->status, sizeof(int), cudaMemcpyDeviceToHost, st));
    LLS_CUDA_TRY(cudaStreamSynchronize(st));
    *status = flags_out | dev_status;
    return LLS_OK;
}

extern "C" int lls_solve_nonlinear_batch(lls_scene* h, double convergence, double relaxation, int max_iterations, void* stream,
                                         int* iterations, double* final_error, int* status) {
    LLS_CHECK_BOUND(h);
    LLS_REQUIRE(iterations && final_error && status, LLS_ERR_ARG, "lls_solve_nonlinear_batch: null output");
    LLS_REQUIRE(h->ctl != nullptr, LLS_ERR_STATE, "lls_solve_nonlinear_batch: no batch bound (lls_bind_batch)");
    cudaStream_t st = (cudaStream_t)stream;
    const int nc = h->ncases;
    // Every case runs the reference's loop (scene.py:851-913) in lock step: iteration k is executed for the cases whose
    // residual norm of iteration k-1 was above the threshold; kernels of masked-out cases return at once.
    for (int k = 1; k <= max_iterations; ++k) {
        LLS_TRY(launch_newton_control(h, k, convergence, st));
        LLS_CUDA_TRY(cudaMemcpyAsync(h->h_ctl, h->ctl + 2 * nc, sizeof(int), cudaMemcpyDeviceToHost, st));
        LLS_CUDA_TRY(cudaStreamSynchronize(st));
        if (h->h_ctl[0] == 0) break;
        {
            PhaseTimer t(h, st, 4);
            LLS_TRY(launch_residual(h, 0, st, true));                            // scene.py:857-858
        }
        {
            PhaseTimer t(h, st, 1);
            LLS_TRY(launch_assemble_jacobian(h, st));                            // scene.py:862-893
        }
        {
            PhaseTimer t(h, st, 2);
            LLS_TRY(lu_factor_solve(h, h->M, wf(h, W_RHS), wf(h, W_X), st, true)); // scene.py:896
        }
        LLS_TRY(launch_newton_update(h, relaxation, st));                        // scene.py:899
    }
    {
        PhaseTimer t(h, st, 4);
        LLS_TRY(launch_residual(h, 0, st, false));                               // scene.py:906-907, 912-913
    }
    LLS_CUDA_TRY(cudaMemcpy2DAsync(final_error, sizeof(double), h->scalars, h->cs.work, sizeof(double), (size_t)nc, cudaMemcpyDeviceToHost, st));
    LLS_CUDA_TRY(cudaMemcpyAsync(iterations, h->ctl + nc, sizeof(int) * (size_t)nc, cudaMemcpyDeviceToHost, st));
    LLS_CUDA_TRY(cudaMemcpy2DAsync(status, sizeof(int), h->status, h->cs.work, sizeof(int), (size_t)nc, cudaMemcpyDeviceToHost, st));
    LLS_CUDA_TRY(cudaStreamSynchronize(st));
    for (int c = 0; c < nc; ++c) {
        final_error[c] = sqrt(final_error[c]);
        if (iterations[c] >= max_iterations) status[c] |= LLS_STATUS_NOT_CONVERGED;   // scene.py:905-908
        if (!(final_error[c] == final_error[c])) status[c] |= LLS_STATUS_NAN;
    }
    return LLS_OK;
}

extern "C" int lls_get_status_batch(lls_scene* h, void* stream, int* status) {
    LLS_CHECK_BOUND(h);
    LLS_REQUIRE(status != nullptr, LLS_ERR_ARG, "lls_get_status_batch: null output");
    cudaStream_t st = (cudaStream_t)stream;
    LLS_CUDA_TRY(cudaMemcpy2DAsync(status, sizeof(int), h->status, h->cs.work, sizeof(int), (size_t)h->ncases, cudaMemcpyDeviceToHost, st));
    LLS_CUDA_TRY(cudaStreamSynchronize(st));
    return LLS_OK;
}

extern "C" int lls_integrate(lls_scene* h, double* d_seg_fm, void* stream) {
    LLS_CHECK_BOUND(h);
    LLS_REQUIRE(d_seg_fm != nullptr, LLS_ERR_ARG, "lls_integrate: null output");
    cudaStream_t st = (cudaStream_t)stream;
    PhaseTimer t(h, st, 5);
    if (h->flags & LLS_FLAG_MATCH_MACHUP_PRO) LLS_TRY(launch_scale_gamma_pro(h, st));   // scene.py:938-939
    LLS_TRY(launch_residual(h, 1, st, false));                                   // v_i with the final gamma, scene.py:943
    return launch_integrate(h, d_seg_fm, st);
}

// ---- one large scene, rows partitioned over the GPUs of a box: the solve in steps, collectives owned by the host ------
#define LLS_CHECK_PART(h) LLS_REQUIRE((h)->ncases == 1, LLS_ERR_STATE, "partitioned entry points need a single scene")

extern "C" int lls_dist_assemble(lls_scene* h, int newton, void* stream) {
    LLS_CHECK_BOUND(h);
    LLS_CHECK_PART(h);
    PhaseTimer t(h, (cudaStream_t)stream, 1);
    return newton ? launch_assemble_jacobian(h, (cudaStream_t)stream) : launch_assemble_linear(h, (cudaStream_t)stream);
}

extern "C" int lls_dist_residual(lls_scene* h, void* stream, double* local_sum) {
    LLS_CHECK_BOUND(h);
    LLS_CHECK_PART(h);
    LLS_REQUIRE(local_sum != nullptr, LLS_ERR_ARG, "lls_dist_residual: null output");
    cudaStream_t st = (cudaStream_t)stream;
    {
        PhaseTimer t(h, st, 4);
        LLS_TRY(launch_residual(h, 0, st, false));
    }
    LLS_TRY(read_scalars(h, st, 1));
    *local_sum = h->h_pinned[0];        // sum of R_i^2 over the rows this rank owns
    return LLS_OK;
}

extern "C" int lls_dist_update_gamma(lls_scene* h, double relaxation, void* stream) {
    LLS_CHECK_BOUND(h);
    LLS_CHECK_PART(h);
    return launch_newton_update(h, relaxation, (cudaStream_t)stream);
}

extern "C" int lls_dist_vectors(lls_scene* h, double** d_rhs, double** d_x, double** d_gamma) {
    LLS_CHECK_BOUND(h);
    if (d_rhs) *d_rhs = wf(h, W_RHS);
    if (d_x) *d_x = wf(h, W_X);
    if (d_gamma) *d_gamma = outf(h, LLS_O_GAMMA);
    return LLS_OK;
}

extern "C" int lls_dist_lu_begin(lls_scene* h, size_t* max_step_doubles, void* stream) {
    LLS_CHECK_BOUND(h);
    LLS_CHECK_PART(h);
    if (max_step_doubles) *max_step_doubles = lu_dist_step_doubles(h, 0);
    return lu_dist_begin(h, (cudaStream_t)stream);
}

extern "C" int lls_dist_lu_step_doubles(lls_scene* h, int k, size_t* doubles) {
    LLS_REQUIRE(h != nullptr && doubles != nullptr && k >= 0 && k < h->ld / LU_NB, LLS_ERR_ARG, "lls_dist_lu_step_doubles: bad arguments");
    *doubles = lu_dist_step_doubles(h, k);
    return LLS_OK;
}

extern "C" int lls_dist_lu_owner(lls_scene* h, int k, int with_rhs, double* d_step, void* stream) {
    LLS_CHECK_BOUND(h);
    LLS_REQUIRE(d_step != nullptr && k >= 0 && k < h->ld / LU_NB, LLS_ERR_ARG, "lls_dist_lu_owner: bad arguments");
    PhaseTimer t(h, (cudaStream_t)stream, 2);
    return lu_dist_owner(h, k, with_rhs ? wf(h, W_RHS) : nullptr, d_step, (cudaStream_t)stream);
}

extern "C" int lls_dist_lu_update(lls_scene* h, int k, int with_rhs, const double* d_step, void* stream) {
    LLS_CHECK_BOUND(h);
    LLS_REQUIRE(d_step != nullptr && k >= 0 && k < h->ld / LU_NB, LLS_ERR_ARG, "lls_dist_lu_update: bad arguments");
    PhaseTimer t(h, (cudaStream_t)stream, 2);
    return lu_dist_update(h, k, with_rhs ? wf(h, W_RHS) : nullptr, d_step, (cudaStream_t)stream);
}

extern "C" int lls_dist_back(lls_scene* h, int k, void* stream) {
    LLS_CHECK_BOUND(h);
    LLS_REQUIRE(k >= 0 && k < h->ld / LU_NB, LLS_ERR_ARG, "lls_dist_back: bad arguments");
    PhaseTimer t(h, (cudaStream_t)stream, 3);
    return lu_dist_back(h, k, wf(h, W_RHS), wf(h, W_X), (cudaStream_t)stream);
}

extern "C" int lls_dist_comm_create(lls_scene* h, int slots, unsigned char* ipc_handle_out64) {
    LLS_CHECK_BOUND(h);
    LLS_CHECK_PART(h);
    return dist_comm_create(h, slots, ipc_handle_out64);
}

extern "C" int lls_dist_comm_bytes(lls_scene* h, int slots, size_t* bytes) {
    LLS_CHECK_BOUND(h);
    LLS_REQUIRE(bytes != nullptr, LLS_ERR_ARG, "lls_dist_comm_bytes: null output");
    *bytes = dist_comm_bytes(h, slots);
    return LLS_OK;
}

extern "C" int lls_dist_comm_attach(lls_scene* h, int slots, void* const* d_mailboxes, void* d_multicast) {
    LLS_CHECK_BOUND(h);
    LLS_CHECK_PART(h);
    LLS_REQUIRE(d_mailboxes != nullptr, LLS_ERR_ARG, "lls_dist_comm_attach: null mailbox list");
    return dist_comm_attach(h, slots, d_mailboxes, d_multicast);
}

extern "C" int lls_dist_comm_multicast(lls_scene* h) { return (h != nullptr) ? dist_comm_multicast(h) : 0; }

extern "C" int lls_dist_debug_schedule(int nblk, int world, int rank, int pair_min_rem, int* out9, int max_launches) {
    return dist_debug_schedule(nblk, world, rank, pair_min_rem, out9, max_launches);
}

extern "C" int lls_dist_debug_role(int block, int own, int rem, int base, int lb0, int nloc, int thin, int* local_block_row, int* block_col) {
    if (local_block_row == nullptr || block_col == nullptr) return -1;
    return dist_debug_role(block, own, rem, base, lb0, nloc, thin, local_block_row, block_col);
}

extern "C" int lls_dist_comm_connect(lls_scene* h, const unsigned char* all_handles) {
    LLS_CHECK_BOUND(h);
    LLS_REQUIRE(all_handles != nullptr || h->cs.P == 1, LLS_ERR_ARG, "lls_dist_comm_connect: null handles");
    return dist_comm_connect(h, all_handles);
}

extern "C" int lls_dist_lu_solve(lls_scene* h, void* stream) {
    LLS_CHECK_BOUND(h);
    LLS_CHECK_PART(h);
    PhaseTimer t(h, (cudaStream_t)stream, 2);
    return lu_dist_factor_solve_p2p(h, wf(h, W_RHS), wf(h, W_X), (cudaStream_t)stream);
}

extern "C" int lls_dist_comm_stats(lls_scene* h, double* out4, void* stream) {
    LLS_REQUIRE(h != nullptr && out4 != nullptr, LLS_ERR_ARG, "lls_dist_comm_stats: null argument");
    return dist_comm_stats(h, out4, (cudaStream_t)stream);
}

extern "C" int lls_dist_comm_destroy(lls_scene* h) {
    LLS_REQUIRE(h != nullptr, LLS_ERR_ARG, "lls_dist_comm_destroy: null handle");
    return dist_comm_destroy(h);
}

extern "C" int lls_get_status(lls_scene* h, void* stream, int* status) {
    LLS_CHECK_BOUND(h);
    LLS_REQUIRE(status != nullptr, LLS_ERR_ARG, "lls_get_status: null output");
    cudaStream_t st = (cudaStream_t)stream;
    LLS_CUDA_TRY(cudaMemcpyAsync(status, h->status, sizeof(int), cudaMemcpyDeviceToHost, st));
    LLS_CUDA_TRY(cudaStreamSynchronize(st));
    return LLS_OK;
}

extern "C" int lls_status_offset(lls_scene* h, size_t* offset_bytes) {
    LLS_REQUIRE(h != nullptr && h->W != nullptr && offset_bytes != nullptr, LLS_ERR_STATE, "lls_status_offset: scene not bound");
    *offset_bytes = (size_t)((const char*)h->status - (const char*)h->W);     // W is the start of the caller's work buffer
    return LLS_OK;
}

extern "C" int lls_lu_factor(lls_scene* h, double* d_mat, void* stream) {
    LLS_REQUIRE(h != nullptr && h->inv != nullptr && d_mat != nullptr, LLS_ERR_STATE, "lls_lu_factor: scene not bound");
    return lu_factor_solve(h, d_mat, nullptr, nullptr, (cudaStream_t)stream);
}

extern "C" int lls_lu_solve(lls_scene* h, const double* d_mat, double* d_rhs_inout, void* stream) {
    LLS_REQUIRE(h != nullptr && h->inv != nullptr && d_mat != nullptr && d_rhs_inout != nullptr, LLS_ERR_STATE, "lls_lu_solve: scene not bound");
    return lu_solve_only(h, d_mat, d_rhs_inout, (cudaStream_t)stream);
}

extern "C" int lls_set_timing(lls_scene* h, int enabled) {
    LLS_REQUIRE(h != nullptr, LLS_ERR_ARG, "null handle");
    h->timing = enabled != 0;
    for (double& m : h->ms) m = 0.0;
    return LLS_OK;
}

extern "C" int lls_get_timings(lls_scene* h, double* ms6) {
    LLS_REQUIRE(h != nullptr && ms6 != nullptr, LLS_ERR_ARG, "null argument");
    for (int i = 0; i < 6; ++i) ms6[i] = h->ms[i];
    return LLS_OK;
}
