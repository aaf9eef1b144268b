// K-state: O(N) flight-state dependent quantities at control points and joints.
// Replaces the per-aircraft loops and O(N) numpy of Scene._calc_invariant_flow_properties
// (scene.py:557-592, 636-666).  One thread per control point; pure latency-bound streaming.
#include "lls_internal.h"
#include "section.cuh"

namespace lls {

template <bool DB>
__global__ void __launch_bounds__(128) flow_state_kernel(int n, int ld, int flags, const double* __restrict__ tab,
                                                        const int32_t* __restrict__ itab, const double* __restrict__ af,
                                                        const double* __restrict__ ac, double* __restrict__ W,
                                                        double* __restrict__ out, int* __restrict__ status, CaseStrides cs) {
    tab = case_ptr(tab, cs.tab);
    ac = case_ptr(ac, cs.ac);
    W = case_work_ptr(W, cs.work);
    out = case_ptr(out, cs.out);
    status = case_work_ptr(status, cs.work);
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ld) return;
    if (i >= n) {  // benign padding so full-ld kernels never see garbage
        for (int f = 0; f < W_NF; ++f) W[(size_t)f * ld + i] = 0.0;
        out[(size_t)LLS_O_GAMMA * ld + i] = 0.0;
        return;
    }
    const double* a = ac + (size_t)itab[(size_t)LLS_I_AIRCRAFT * ld + i] * LLS_AC_STRIDE;
    const Vec3 vtrans{a[LLS_AC_VTRANS], a[LLS_AC_VTRANS + 1], a[LLS_AC_VTRANS + 2]};
    const Vec3 w{a[LLS_AC_W], a[LLS_AC_W + 1], a[LLS_AC_W + 2]};
    const Vec3 cgr{a[LLS_AC_CGR], a[LLS_AC_CGR + 1], a[LLS_AC_CGR + 2]};

    const Vec3 vinf = vtrans + ld3(tab, ld, LLS_F_VWIND, i);                     // scene.py:571
    const Vec3 vir = vinf - cross(w, ld3(tab, ld, LLS_F_RCG, i));                // scene.py:569, 572
    Vec3 j0, j1;
    if (flags & LLS_FLAG_MATCH_MACHUP_PRO) {                                     // scene.py:575-577
        j0 = vinf;
        j1 = vinf;
    } else {                                                                     // scene.py:579-582
        j0 = vinf - cross(w, ld3(tab, ld, LLS_F_P0J, i) - cgr);
        j1 = vinf - cross(w, ld3(tab, ld, LLS_F_P1J, i) - cgr);
    }
    const double vinf_mag = norm(vinf), vir_mag = norm(vir);
    st3(W, ld, W_VINF, i, vinf);
    st3(W, ld, W_VIR, i, vir);
    W[(size_t)W_VINF_MAG * ld + i] = vinf_mag;
    W[(size_t)W_VIR_MAG * ld + i] = vir_mag;
    st3(W, ld, W_UINF, i, (1.0 / vinf_mag) * vinf);
    {   // trailing directions (scene.py:591-592); a true division like the reference
        double m0 = norm(j0), m1 = norm(j1);
        Vec3 u0{j0.x / m0, j0.y / m0, j0.z / m0}, u1{j1.x / m1, j1.y / m1, j1.z / m1};
        if (flags & LLS_FLAG_CONSTRAIN_VORTEX_SHEET) {                           // scene.py:595-611
            // the sheet is kept in the aircraft's x-y plane: project the body z axis (earth axes) out, then renormalise
            const Vec3 zf{a[LLS_AC_ZF], a[LLS_AC_ZF + 1], a[LLS_AC_ZF + 2]};
            u0 = u0 - dot(zf, u0) * zf;
            u1 = u1 - dot(zf, u1) * zf;
            m0 = norm(u0);
            m1 = norm(u1);
            u0 = Vec3{u0.x / m0, u0.y / m0, u0.z / m0};
            u1 = Vec3{u1.x / m1, u1.y / m1, u1.z / m1};
        }
        st3(out, ld, LLS_O_UTRAIL0, i, u0);
        st3(out, ld, LLS_O_UTRAIL1, i, u1);
    }
    const Vec3 us = ld3(tab, ld, LLS_F_US, i), ua = ld3(tab, ld, LLS_F_UA, i), un = ld3(tab, ld, LLS_F_UN, i);
    double vinf_ip = vinf_mag, vir_ip = vir_mag;
    if (flags & LLS_FLAG_IN_PLANE) {                                             // scene.py:637-641
        const Vec3 p = vinf - dot(us, vinf) * us;
        vinf_ip = norm(p);
        const Vec3 pr = vir - dot(us, vir) * us;
        vir_ip = norm(pr);
    }
    W[(size_t)W_VINF_IP * ld + i] = vinf_ip;
    // Reynolds and Mach number of the linear pre-pass (scene.py:642-646; the in-plane branch really mixes the two speeds)
    const double cbar = tab[(size_t)LLS_F_CBAR * ld + i];
    const double Re = ((flags & LLS_FLAG_IN_PLANE) ? vir_ip : vinf_mag) * cbar / tab[(size_t)LLS_F_NU * ld + i];
    const double Mach = ((flags & LLS_FLAG_IN_PLANE) ? vinf_ip : vinf_mag) / tab[(size_t)LLS_F_SOS * ld + i];
    const double alpha_inf = atan2(dot(vir, un), dot(vir, ua));                  // scene.py:649-651
    SectionParams sp = load_section(tab, itab, af, ld, i);
    bool oob = false;
    const double CLa = section_CLa<DB>(sp, alpha_inf, Re, Mach);                 // scene.py:659-661
    double CL = section_CL<DB>(sp, alpha_inf, Re, Mach, oob);
    const double aL0 = section_aL0<DB>(sp, Re, Mach);
    if (oob) atomicOr(status, LLS_STATUS_SECTION_BOUNDS);
    if (flags & LLS_FLAG_SWEPT_SECTIONS) {                                       // scene.py:665-666, 797
        CL += CLa * (aL0 - aL0 * tab[(size_t)LLS_F_CSWI * ld + i]);
    }
    W[(size_t)W_CLA0 * ld + i] = CLa;
    W[(size_t)W_CL0 * ld + i] = CL;
    out[(size_t)LLS_O_AL0 * ld + i] = aL0;
    // what the reference's _CL holds after a LINEAR solve (scene.py:659-666; distributions() reports it as section_CL); every
    // residual evaluation of a nonlinear solve overwrites it (scene.py:765-779)
    out[(size_t)LLS_O_CL * ld + i] = CL;
}

int launch_flow_state(Scene* s, cudaStream_t st) {
    int block = 128, grid = (s->ld + block - 1) / block;
    // the status word of every case is cleared in stream order BEFORE the launch: blocks of this launch set bits in it (section
    // bounds), and a reset from inside the kernel could wipe a bit another block had already set
    if (s->ncases > 1) LLS_CUDA_TRY(cudaMemset2DAsync(s->status, s->cs.work, 0, sizeof(int), (size_t)s->ncases, st));
    else LLS_CUDA_TRY(cudaMemsetAsync(s->status, 0, sizeof(int), st));
    auto kernel = (s->flags & LLS_FLAG_DATABASE_AIRFOILS) ? flow_state_kernel<true> : flow_state_kernel<false>;
    kernel<<<dim3(grid, 1, s->ncases), block, 0, st>>>(s->n, s->ld, s->flags, s->tab, s->itab, s->af, s->ac, s->W, s->out, s->status,
                                                       strides(s, false));
    count_launch();
    LLS_CUDA_TRY(cudaGetLastError());
    return LLS_OK;
}

}  // namespace lls
