// K2, K4-K8: linear-system assembly, induced-velocity matvec fused with the section-lift residual,
// Jacobian assembly and the Newton update.  All HBM-streaming kernels over V[i][c][j] (24 B/pair).
//
//   assemble_linear   A, b of Scene._solve_linear                           scene.py:811-824
//   residual          _calc_v_i + _lifting_line_residual + _get_section_lift  scene.py:705-797
//   assemble_jacobian J of Scene._solve_nonlinear                            scene.py:862-893
//   newton_update     gamma += relaxation * dGamma                            scene.py:899
#include "lls_internal.h"
#include "section.cuh"

namespace lls {

// ---------------------------------------------------------------------------------------------
// linear system rows: g_i = -(V_inf CLa dS)_i u_n_i, diag_i = 2 |v_inf_rot x dl|, b_i     (scene.py:811-824)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) linear_rows_kernel(int n, int ld, int flags, const double* __restrict__ tab, double* __restrict__ W,
                                                          CaseStrides cs) {
    tab = case_ptr(tab, cs.tab);
    W = case_work_ptr(W, cs.work);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double Vref = (flags & LLS_FLAG_IN_PLANE) ? W[(size_t)W_VINF_IP * ld + i] : W[(size_t)W_VINF_MAG * ld + i];
    const double CLa = W[(size_t)W_CLA0 * ld + i], CL = W[(size_t)W_CL0 * ld + i];
    const double dS = tab[(size_t)LLS_F_DS * ld + i];
    const Vec3 un = ld3(tab, ld, LLS_F_UN, i);
    st3(W, ld, W_G, i, (-(Vref * CLa * dS)) * un);
    const Vec3 w = cross(ld3(W, ld, W_VIR, i), ld3(tab, ld, LLS_F_DL, i));
    W[(size_t)W_DIAG * ld + i] = 2.0 * norm(w);
    W[(size_t)W_RHS * ld + i] = Vref * Vref * dS * CL;
}

// ---------------------------------------------------------------------------------------------
// N^2 assembly.  Thread <-> column j (coalesced over V's component planes and over the output row),
// CTA walks ASM_ROWS rows whose row vectors sit in shared memory.
//   LINEAR : M_ij = V_ij . g_i + delta_ij diag_i
//   NEWTON : Vt = V_ij - u_s_j (u_s_j . V_ij)            (projector of the VORTEX j: scene.py:847 broadcasting)
//            M_ij = ws_i . (Vt x dl_j) - Vt . g_i - cbar_j (Vt . e_i) + delta_ij diag_i     (scene.py:867-893;
//            np.cross(V_ji, dl) pairs dl with j, c_bar in the CLRe term pairs with j — both as in the reference)
// Rows n..ld of M are set to the identity so the LU runs on whole blocks.
// ---------------------------------------------------------------------------------------------
constexpr int ASM_ROWS = 16;

template <bool NEWTON>
__global__ void __launch_bounds__(256) assemble_kernel(int n, int ld, int flags, const double* __restrict__ tab,
                                                       const double* __restrict__ W, const double* __restrict__ V,
                                                       double* __restrict__ M, CaseStrides cs) {
    if (LLS_CASE_INACTIVE(cs)) return;
    tab = case_ptr(tab, cs.tab);
    W = case_work_ptr(W, cs.work);
    V = case_ptr(V, cs.V);
    M = case_ptr(M, cs.M);
    __shared__ double sg[3][ASM_ROWS], sws[3][ASM_ROWS], se[3][ASM_ROWS], sdiag[ASM_ROWS];
    const int lrow0 = blockIdx.y * ASM_ROWS;          // local (this rank's V / M) and global first row of the chunk
    const int row0 = global_row(cs, lrow0);
    if (threadIdx.x < ASM_ROWS) {
        const int i = row0 + threadIdx.x;
        const bool ok = i < n;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            sg[c][threadIdx.x] = ok ? W[(size_t)(W_G + c) * ld + i] : 0.0;
            sws[c][threadIdx.x] = (ok && NEWTON) ? W[(size_t)(W_WS + c) * ld + i] : 0.0;
            se[c][threadIdx.x] = (ok && NEWTON) ? W[(size_t)(W_E + c) * ld + i] : 0.0;
        }
        sdiag[threadIdx.x] = ok ? W[(size_t)W_DIAG * ld + i] : 1.0;
    }
    __syncthreads();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ld) return;
    Vec3 us{0, 0, 0}, dl{0, 0, 0};
    double cbar = 0.0;
    const bool in_plane = (flags & LLS_FLAG_IN_PLANE) != 0;
    if (NEWTON && j < n) {
        us = ld3(tab, ld, LLS_F_US, j);
        dl = ld3(tab, ld, LLS_F_DL, j);
        cbar = tab[(size_t)LLS_F_CBAR * ld + j];
    }
#pragma unroll 4
    for (int r = 0; r < ASM_ROWS; ++r) {
        const int i = row0 + r;
        if (i >= ld) break;
        double m;
        if (i < n) {
            const double* v = V + (size_t)(lrow0 + r) * 3 * ld + j;
            Vec3 vt{v[0], v[(size_t)ld], v[(size_t)2 * ld]};
            const Vec3 g{sg[0][r], sg[1][r], sg[2][r]};
            if (NEWTON) {
                if (in_plane) vt = vt - dot(us, vt) * us;
                const Vec3 ws{sws[0][r], sws[1][r], sws[2][r]};
                const Vec3 e{se[0][r], se[1][r], se[2][r]};
                m = dot(ws, cross(vt, dl)) - dot(vt, g) - cbar * dot(vt, e);
            } else {
                m = dot(vt, g);
            }
            if (i == j) m += sdiag[r];
        } else {
            m = (i == j) ? 1.0 : 0.0;
        }
        M[(size_t)(lrow0 + r) * ld + j] = m;
    }
}

// ---------------------------------------------------------------------------------------------
// residual: one CTA per control point i.  v_i = v_inf_rot + sum_j V_ij gamma_j (scene.py:728), then the whole
// section-lift evaluation (scene.py:732-791), R_i (scene.py:723) and the row vectors of the Jacobian.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// everything of a residual row after v_i: section lift (scene.py:732-791), R_i (scene.py:723), the row vectors of the Jacobian.
// Per-case pointers already applied by the caller.
template <bool DB>
__device__ __forceinline__ void residual_row_tail(int i, const Vec3& vi, int ld, int flags, const double* __restrict__ tab,
                                                  const int32_t* __restrict__ itab, const double* __restrict__ af, double* __restrict__ W,
                                                  double* __restrict__ out, double* __restrict__ partial, int* __restrict__ status_case) {
    const double* gamma = out + (size_t)LLS_O_GAMMA * ld;
    const double gam = gamma[i];
    const Vec3 dl = ld3(tab, ld, LLS_F_DL, i);
    const Vec3 w = cross(vi, dl);                                                 // scene.py:715-717
    const double wmag = norm(w);
    const double L_vortex = 2.0 * wmag * gam;
    const Vec3 us = ld3(tab, ld, LLS_F_US, i), ua = ld3(tab, ld, LLS_F_UA, i), un = ld3(tab, ld, LLS_F_UN, i);
    const bool in_plane = flags & LLS_FLAG_IN_PLANE, total = flags & LLS_FLAG_TOTAL_VELOCITY;
    const bool swept = flags & LLS_FLAG_SWEPT_SECTIONS, pro = flags & LLS_FLAG_MATCH_MACHUP_PRO;
    const Vec3 vref = in_plane ? (vi - dot(us, vi) * us) : vi;                    // scene.py:737-745
    const double V2 = dot(vref, vref), Vmag = sqrt(V2);
    const double cbar = tab[(size_t)LLS_F_CBAR * ld + i], nu = tab[(size_t)LLS_F_NU * ld + i], sos = tab[(size_t)LLS_F_SOS * ld + i];
    const double Re = Vmag * cbar / nu, Mach = Vmag / sos;
    const double va = dot(vi, ua), vn = dot(vi, un);                              // scene.py:751-753
    const double alpha = atan2(vn, va);
    const SectionParams sp = load_section(tab, itab, af, ld, i);
    bool oob = false;
    const double CLa = section_CLa<DB>(sp, alpha, Re, Mach);                      // scene.py:765-768
    double CL = section_CL<DB>(sp, alpha, Re, Mach, oob);
    const double CLRe = section_CLRe<DB>(sp, alpha, Re, Mach), CLM = section_CLM<DB>(sp, alpha, Re, Mach);
    if (oob) atomicOr(status_case, LLS_STATUS_SECTION_BOUNDS);
    const double dS = tab[(size_t)LLS_F_DS * ld + i];
    const double Vinf = W[(size_t)W_VINF_MAG * ld + i];
    const double Vfix = in_plane ? W[(size_t)W_VINF_IP * ld + i] : Vinf;
    double L_section, scale;
    if (pro) {                                                                    // scene.py:774-775
        L_section = Vinf * Vinf * CL * dS;
    } else {
        if (swept) {                                                              // scene.py:778-779, 797
            const double aL0 = out[(size_t)LLS_O_AL0 * ld + i];
            CL += CLa * (aL0 - aL0 * tab[(size_t)LLS_F_CSWI * ld + i]);
        }
        L_section = (total ? V2 : Vfix * Vfix) * CL * dS;                         // scene.py:782-791
    }
    scale = (total ? V2 : Vfix * Vfix) * dS;                                      // scene.py:881-890
    const double R = L_vortex - L_section;                                        // scene.py:723
    // per-control-point results other code reads after a solve
    st3(out, ld, LLS_O_VI, i, vi);
    out[(size_t)LLS_O_ALPHA * ld + i] = alpha;
    out[(size_t)LLS_O_CL * ld + i] = CL;
    out[(size_t)LLS_O_RE * ld + i] = Re;
    out[(size_t)LLS_O_MACH * ld + i] = Mach;
    out[(size_t)LLS_O_R * ld + i] = R;
    partial[i] = R * R;
    // Jacobian row data (scene.py:867-893)
    st3(W, ld, W_WS, i, (2.0 * gam / wmag) * w);
    const double c2 = total ? 2.0 * dS * CL : 0.0;
    const double c3 = scale * CLa / (vn * vn + va * va);
    const double c4 = scale * CLRe / (nu * Vmag), c5 = scale * CLM / (sos * Vmag);
    st3(W, ld, W_G, i, (c2 + c5) * vref + c3 * (va * un - vn * ua));
    st3(W, ld, W_E, i, c4 * vref);
    W[(size_t)W_DIAG * ld + i] = 2.0 * wmag;
    W[(size_t)W_RHS * ld + i] = -R;                                               // scene.py:896
}

// THREADS threads share a row; ROWS rows per CTA.  ROWS > 1 needs THREADS == 32 (one warp per row, reduction by shuffles only): the
// shape for the small scenes of large batches, where a row is shorter than a CTA and the scalar tail of the row (section lift,
// Jacobian row vectors) is what there is to hide.
template <int THREADS, int ROWS = 1, bool TAIL = true, bool DB = false>
__global__ void __launch_bounds__(THREADS * ROWS) residual_kernel(int n, int ld, int flags, const double* __restrict__ tab,
                                                           const int32_t* __restrict__ itab, const double* __restrict__ af,
                                                           const double* __restrict__ V, double* __restrict__ W,
                                                           double* __restrict__ out, double* __restrict__ partial, int vel_only,
                                                           int* __restrict__ status, CaseStrides cs) {
    if (LLS_CASE_INACTIVE(cs)) return;
    tab = case_ptr(tab, cs.tab);
    V = case_ptr(V, cs.V);
    W = case_work_ptr(W, cs.work);
    out = case_ptr(out, cs.out);
    partial = case_work_ptr(partial, cs.work);
    static_assert(ROWS == 1 || THREADS == 32, "several rows per CTA: one warp per row");
    const int tid = (ROWS == 1) ? threadIdx.x : (threadIdx.x & 31);
    const int li = (ROWS == 1) ? blockIdx.x : blockIdx.x * ROWS + (threadIdx.x >> 5);    // row of this rank's V
    const int i = global_row(cs, li);
    if (i >= n) return;
    const double* gamma = out + (size_t)LLS_O_GAMMA * ld;
    const double2* v0 = reinterpret_cast<const double2*>(V + (size_t)li * 3 * ld);
    const double2* v1 = reinterpret_cast<const double2*>(V + (size_t)li * 3 * ld + ld);
    const double2* v2 = reinterpret_cast<const double2*>(V + (size_t)li * 3 * ld + 2 * (size_t)ld);
    const double2* g2 = reinterpret_cast<const double2*>(gamma);
    double sx = 0.0, sy = 0.0, sz = 0.0;
    for (int j = tid; j < ld / 2; j += THREADS) {
        const double2 g = g2[j], a = v0[j], b = v1[j], c = v2[j];
        sx = fma(a.x, g.x, fma(a.y, g.y, sx));
        sy = fma(b.x, g.x, fma(b.y, g.y, sy));
        sz = fma(c.x, g.x, fma(c.y, g.y, sz));
    }
    sx = warp_sum(sx);
    sy = warp_sum(sy);
    sz = warp_sum(sz);
    if (THREADS > 32) {
        __shared__ double red[3][THREADS / 32];
        if ((threadIdx.x & 31) == 0) {
            red[0][threadIdx.x >> 5] = sx;
            red[1][threadIdx.x >> 5] = sy;
            red[2][threadIdx.x >> 5] = sz;
        }
        __syncthreads();
        if (threadIdx.x != 0) return;
        sx = sy = sz = 0.0;
        for (int w = 0; w < THREADS / 32; ++w) {
            sx += red[0][w];
            sy += red[1][w];
            sz += red[2][w];
        }
    } else if (tid != 0) {
        return;
    }
    const Vec3 vi = ld3(W, ld, W_VIR, i) + Vec3{sx, sy, sz};                       // scene.py:728
    if (vel_only || !TAIL) {  // integration pre-pass of a linear solve (scene.py:943), or the tail runs as a kernel of its own
        st3(out, ld, LLS_O_VI, i, vi);
        return;
    }
    residual_row_tail<DB>(i, vi, ld, flags, tab, itab, af, W, out, partial, case_work_ptr(status, cs.work));
}

// The tail of every row as one THREAD per row (batches of small scenes): all table reads are coalesced along i and no lane idles,
// where the fused kernel leaves the whole tail to one lane of the row's warp.
template <bool DB>
__global__ void __launch_bounds__(128) residual_tail_kernel(int n, int ld, int flags, const double* __restrict__ tab,
                                                            const int32_t* __restrict__ itab, const double* __restrict__ af,
                                                            double* __restrict__ W, double* __restrict__ out, double* __restrict__ partial,
                                                            int* __restrict__ status, CaseStrides cs) {
    if (LLS_CASE_INACTIVE(cs)) return;
    tab = case_ptr(tab, cs.tab);
    W = case_work_ptr(W, cs.work);
    out = case_ptr(out, cs.out);
    partial = case_work_ptr(partial, cs.work);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    residual_row_tail<DB>(i, ld3(out, ld, LLS_O_VI, i), ld, flags, tab, itab, af, W, out, partial, case_work_ptr(status, cs.work));
}

// deterministic sum of partial[0..n) -> scalars[slot]
__global__ void __launch_bounds__(256) sum_kernel(int n, const double* __restrict__ partial, double* __restrict__ scalars, int slot,
                                                  CaseStrides cs) {
    if (LLS_CASE_INACTIVE(cs)) return;
    partial = case_work_ptr(partial, cs.work);
    scalars = case_work_ptr(scalars, cs.work);
    __shared__ double red[8];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256)
        if (owns_row(cs, i)) s += partial[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        scalars[slot] = t;
    }
}

__global__ void __launch_bounds__(128) newton_update_kernel(int n, int ld, double relaxation, const double* __restrict__ W, double* __restrict__ out,
                                                            CaseStrides cs) {
    if (LLS_CASE_INACTIVE(cs)) return;
    W = case_work_ptr(W, cs.work);
    out = case_ptr(out, cs.out);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[(size_t)LLS_O_GAMMA * ld + i] += relaxation * W[(size_t)W_X * ld + i];       // scene.py:899
}

__global__ void __launch_bounds__(128) copy_gamma_kernel(int n, int ld, const double* __restrict__ W, double* __restrict__ out, CaseStrides cs) {
    W = case_work_ptr(W, cs.work);
    out = case_ptr(out, cs.out);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[(size_t)LLS_O_GAMMA * ld + i] = W[(size_t)W_X * ld + i];                     // scene.py:827
}

int launch_assemble_linear(Scene* s, cudaStream_t st) {
    linear_rows_kernel<<<dim3((s->n + 127) / 128, 1, s->ncases), 128, 0, st>>>(s->n, s->ld, s->flags, s->tab, s->W, strides(s, false));
    count_launch();
    dim3 grid((s->ld + 255) / 256, s->n_loc_blocks * (64 / ASM_ROWS), s->ncases);
    assemble_kernel<false><<<grid, 256, 0, st>>>(s->n, s->ld, s->flags, s->tab, s->W, s->V, s->M, strides(s, false));
    count_launch();
    LLS_CUDA_TRY(cudaGetLastError());
    return LLS_OK;
}

int launch_residual(Scene* s, int vel_only, cudaStream_t st, bool masked) {
    const int rows = s->cs.P == 1 ? s->n : s->n_loc_blocks * 64;
    const bool db = (s->flags & LLS_FLAG_DATABASE_AIRFOILS) != 0;   // section model with table look-ups: the <DB = true> kernels
    if (s->ld <= 512 && s->ncases > 1) {
        // small scenes in large batches: one warp per row, 8 rows per CTA
        // ... and the tail of the rows as a kernel of its own, one thread per row
        residual_kernel<32, 8, false><<<dim3((rows + 7) / 8, 1, s->ncases), 256, 0, st>>>(s->n, s->ld, s->flags, s->tab, s->itab, s->af, s->V, s->W,
                                                                                       s->out, s->partial, vel_only, s->status, strides(s, masked));
        if (!vel_only) {
            auto tail = db ? residual_tail_kernel<true> : residual_tail_kernel<false>;
            tail<<<dim3((s->n + 127) / 128, 1, s->ncases), 128, 0, st>>>(s->n, s->ld, s->flags, s->tab, s->itab, s->af, s->W, s->out, s->partial,
                                                                      s->status, strides(s, masked));
            count_launch();
        }
    } else {
        auto fused = db ? residual_kernel<128, 1, true, true> : residual_kernel<128, 1, true, false>;
        fused<<<dim3(rows, 1, s->ncases), 128, 0, st>>>(s->n, s->ld, s->flags, s->tab, s->itab, s->af, s->V, s->W, s->out, s->partial, vel_only,
                                                        s->status, strides(s, masked));
    }
    count_launch();
    if (!vel_only) {
        sum_kernel<<<dim3(1, 1, s->ncases), 256, 0, st>>>(s->n, s->partial, s->scalars, 0, strides(s, masked));
        count_launch();
    }
    LLS_CUDA_TRY(cudaGetLastError());
    return LLS_OK;
}

int launch_assemble_jacobian(Scene* s, cudaStream_t st) {
    dim3 grid((s->ld + 255) / 256, s->n_loc_blocks * (64 / ASM_ROWS), s->ncases);
    assemble_kernel<true><<<grid, 256, 0, st>>>(s->n, s->ld, s->flags, s->tab, s->W, s->V, s->M, strides(s, true));
    count_launch();
    LLS_CUDA_TRY(cudaGetLastError());
    return LLS_OK;
}

int launch_newton_update(Scene* s, double relaxation, cudaStream_t st) {
    dim3 grid((s->n + 127) / 128, 1, s->ncases);
    if (relaxation < 0.0) {
        copy_gamma_kernel<<<grid, 128, 0, st>>>(s->n, s->ld, s->W, s->out, strides(s, false));
        count_launch();
    } else {
        newton_update_kernel<<<grid, 128, 0, st>>>(s->n, s->ld, relaxation, s->W, s->out, strides(s, true));
        count_launch();
    }
    LLS_CUDA_TRY(cudaGetLastError());
    return LLS_OK;
}

// ---------------------------------------------------------------------------------------------
// batch control: which cases take part in Newton iteration k (reference loop test `while error > convergence`,
// scene.py:853; error of iteration 0 is 100, scene.py:852), their iteration counters and the number still active
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) newton_control_kernel(int ncases, int k, double convergence, const double* __restrict__ scalars,
                                                             size_t work_stride, int* __restrict__ ctl) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncases) return;
    int* active = ctl;
    int* iters = ctl + ncases;
    int a = 1;
    if (k > 1) {
        const double err2 = *reinterpret_cast<const double*>(reinterpret_cast<const char*>(scalars) + (size_t)c * work_stride);
        a = active[c] && (sqrt(err2) > convergence);   // NaN compares false: the case stops, like the reference's loop test
    } else {
        iters[c] = 0;
    }
    active[c] = a;
    if (a) {
        iters[c] = k;
        atomicAdd(ctl + 2 * ncases, 1);
    }
}

int launch_newton_control(Scene* s, int k, double convergence, cudaStream_t st) {
    LLS_CUDA_TRY(cudaMemsetAsync(s->ctl + 2 * s->ncases, 0, sizeof(int), st));
    newton_control_kernel<<<(s->ncases + 255) / 256, 256, 0, st>>>(s->ncases, k, convergence, s->scalars, s->cs.work, s->ctl);
    count_launch();
    LLS_CUDA_TRY(cudaGetLastError());
    return LLS_OK;
}

}  // namespace lls
