// K9+K10: force and moment elements at every control point and their per-segment sums.
// Replaces the numeric part of Scene._integrate_forces_and_moments (scene.py:941-1046) and the
// per-segment np.sum(axis=0) reductions (scene.py:1108-1117).  One CTA per wing segment: its control
// points are a contiguous index range, so the sums are warp-shuffle reductions over that range —
// fixed order, hence bit-reproducible run to run.
#include "lls_internal.h"
#include "section.cuh"

namespace lls {

__global__ void __launch_bounds__(128) scale_gamma_pro_kernel(int n, int ld, const double* __restrict__ W, double* __restrict__ out,
                                                              CaseStrides cs) {
    W = case_work_ptr(W, cs.work);
    out = case_ptr(out, cs.out);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double r = W[(size_t)W_VIR_MAG * ld + i] / W[(size_t)W_VINF_MAG * ld + i];   // scene.py:938-939
    out[(size_t)LLS_O_GAMMA * ld + i] *= r * r;
}

__device__ __forceinline__ int lower_bound_seg(const int32_t* __restrict__ seg, int n, int key) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (seg[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

constexpr int INT_THREADS = 128;

template <bool DB>
__global__ void __launch_bounds__(INT_THREADS) integrate_kernel(int n, int ld, int flags, const double* __restrict__ tab,
                                                                const int32_t* __restrict__ itab, const double* __restrict__ af,
                                                                const double* __restrict__ W, double* __restrict__ out,
                                                                double* __restrict__ seg_fm, int* __restrict__ status, CaseStrides cs) {
    tab = case_ptr(tab, cs.tab);
    status = case_work_ptr(status, cs.work);
    W = case_work_ptr(W, cs.work);
    out = case_ptr(out, cs.out);
    seg_fm = case_ptr(seg_fm, cs.seg);
    const int s = blockIdx.x;
    const int32_t* segid = itab + (size_t)LLS_I_SEGMENT * ld;
    const int begin = lower_bound_seg(segid, n, s), end = lower_bound_seg(segid, n, s + 1);
    const bool in_plane = flags & LLS_FLAG_IN_PLANE, total = flags & LLS_FLAG_TOTAL_VELOCITY;
    const bool swept = flags & LLS_FLAG_SWEPT_SECTIONS, pro = flags & LLS_FLAG_MATCH_MACHUP_PRO;
    double acc[12];
#pragma unroll
    for (int c = 0; c < 12; ++c) acc[c] = 0.0;

    for (int i = begin + threadIdx.x; i < end; i += INT_THREADS) {
        if (!owns_row(cs, i)) continue;               // row-partitioned scene: the other ranks add their rows (all-reduce on the host side)
        const Vec3 vi = ld3(out, ld, LLS_O_VI, i);                                   // scene.py:943
        const double Vi2 = dot(vi, vi), Vi = sqrt(Vi2);                              // scene.py:944-946
        const Vec3 us = ld3(tab, ld, LLS_F_US, i), ua = ld3(tab, ld, LLS_F_UA, i), un = ld3(tab, ld, LLS_F_UN, i);
        const Vec3 dl = ld3(tab, ld, LLS_F_DL, i), rcg = ld3(tab, ld, LLS_F_RCG, i);
        const double rho = tab[(size_t)LLS_F_RHO * ld + i], nu = tab[(size_t)LLS_F_NU * ld + i], sos = tab[(size_t)LLS_F_SOS * ld + i];
        const double cbar = tab[(size_t)LLS_F_CBAR * ld + i], dS = tab[(size_t)LLS_F_DS * ld + i], cswi = tab[(size_t)LLS_F_CSWI * ld + i];
        const double gam = out[(size_t)LLS_O_GAMMA * ld + i];
        double Vip2 = Vi2;
        if (in_plane) {                                                              // scene.py:947-949
            const Vec3 vip = vi - dot(us, vi) * us;
            Vip2 = dot(vip, vip);
        }
        const Vec3 dF_inv = (rho * gam) * cross(vi, dl);                             // scene.py:952
        const double alpha = atan2(dot(vi, un), dot(vi, ua));                        // scene.py:955-957
        double Re_unswept, alpha_unswept;
        if (swept) {                                                                 // scene.py:958-970
            Re_unswept = Vi * cbar * cswi / nu;
            alpha_unswept = atan2(dot(vi, ld3(tab, ld, LLS_F_UNU, i)), dot(vi, ld3(tab, ld, LLS_F_UAU, i)));
        } else {
            Re_unswept = Vi * cbar / nu;
            alpha_unswept = alpha;
        }
        const double Vsec = in_plane ? sqrt(Vip2) : Vi;                              // scene.py:975-981
        const double Re = Vsec * cbar / nu, Mach = Vsec / sos;
        double redim_full, redim_ip;
        if (total) {                                                                 // scene.py:984-991
            redim_full = 0.5 * rho * Vi2 * dS;
            redim_ip = 0.5 * rho * Vip2 * dS;
        } else {
            const double Vinf = W[(size_t)W_VINF_MAG * ld + i], Vinf_ip = W[(size_t)W_VINF_IP * ld + i];
            redim_full = 0.5 * rho * Vinf * Vinf * dS;
            redim_ip = 0.5 * rho * Vinf_ip * Vinf_ip * dS;
        }
        const SectionParams sp = load_section(tab, itab, af, ld, i);
        bool oob = false;
        const double CD = section_CD<DB>(sp, alpha_unswept, Re_unswept, Mach, oob);  // scene.py:1014-1017
        double Cm = section_Cm<DB>(sp, alpha, Re, Mach, oob);                        // scene.py:1020
        if (oob) atomicOr(status, LLS_STATUS_SECTION_BOUNDS);
        if (swept) Cm *= cswi;                                                       // scene.py:1026
        const Vec3 dM_section = ((in_plane ? redim_ip : redim_full) * cbar * Cm) * us;   // scene.py:1029-1032
        const Vec3 dM_inv = cross(rcg, dF_inv) + dM_section;                         // scene.py:1035-1036
        const double dD = redim_full * CD;                                           // scene.py:1039
        const Vec3 udir = (total || pro) ? Vec3{vi.x / Vi, vi.y / Vi, vi.z / Vi} : ld3(W, ld, W_UINF, i);   // scene.py:946, 1040-1043
        const Vec3 dF_visc = dD * udir;
        const Vec3 dM_visc = cross(rcg, dF_visc);                                    // scene.py:1046

        out[(size_t)LLS_O_ALPHA * ld + i] = alpha;
        out[(size_t)LLS_O_CD * ld + i] = CD;
        out[(size_t)LLS_O_CM * ld + i] = Cm;
        out[(size_t)LLS_O_RE * ld + i] = Re;
        out[(size_t)LLS_O_MACH * ld + i] = Mach;
        out[(size_t)LLS_O_VI2 * ld + i] = Vi2;
        out[(size_t)LLS_O_REDIM_IP * ld + i] = redim_ip;
        out[(size_t)LLS_O_REDIM_FULL * ld + i] = redim_full;
        st3(out, ld, LLS_O_DFINV, i, dF_inv);
        st3(out, ld, LLS_O_DMINV, i, dM_inv);
        st3(out, ld, LLS_O_DFVISC, i, dF_visc);
        st3(out, ld, LLS_O_DMVISC, i, dM_visc);
        acc[0] += dF_inv.x;  acc[1] += dF_inv.y;  acc[2] += dF_inv.z;
        acc[3] += dM_inv.x;  acc[4] += dM_inv.y;  acc[5] += dM_inv.z;
        acc[6] += dF_visc.x; acc[7] += dF_visc.y; acc[8] += dF_visc.z;
        acc[9] += dM_visc.x; acc[10] += dM_visc.y; acc[11] += dM_visc.z;
    }
    __shared__ double red[INT_THREADS / 32][12];
#pragma unroll
    for (int c = 0; c < 12; ++c) {
        double v = acc[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][c] = v;
    }
    __syncthreads();
    if (threadIdx.x < 12) {
        double t = 0.0;
        for (int w = 0; w < INT_THREADS / 32; ++w) t += red[w][threadIdx.x];
        seg_fm[(size_t)s * 12 + threadIdx.x] = t;
    }
}

int launch_integrate(Scene* s, double* d_seg_fm, cudaStream_t st) {
    auto kernel = (s->flags & LLS_FLAG_DATABASE_AIRFOILS) ? integrate_kernel<true> : integrate_kernel<false>;
    kernel<<<dim3(s->n_seg, 1, s->ncases), INT_THREADS, 0, st>>>(s->n, s->ld, s->flags, s->tab, s->itab, s->af, s->W, s->out, d_seg_fm, s->status,
                                                                 strides(s, false));
    count_launch();
    LLS_CUDA_TRY(cudaGetLastError());
    return LLS_OK;
}

int launch_scale_gamma_pro(Scene* s, cudaStream_t st) {
    scale_gamma_pro_kernel<<<dim3((s->n + 127) / 128, 1, s->ncases), 128, 0, st>>>(s->n, s->ld, s->W, s->out, strides(s, false));
    count_launch();
    LLS_CUDA_TRY(cudaGetLastError());
    return LLS_OK;
}

}  // namespace lls
