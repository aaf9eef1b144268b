// K0+K1: horseshoe-vortex influence matrix V_ji with the effective (blended, jointed) lifting-line
// geometry recomputed per pair in registers.
//
// Replaces, in one pass and without any N x N x 3 geometry arrays:
//   * Airplane._calculate_geometry effective-LAC loop            airplane.py:569-689
//   * Scene._perform_geometry_and_atmos_calcs cross blocks + V_ji_const   scene.py:469-541
//   * trailing legs of Scene._calc_invariant_flow_properties     scene.py:621-634
//
// Mapping: one warp owns a window of 32 consecutive vortices j (lanes), of which the inner 30 are
// written; the two halo lanes exist so the three-point tangent stencil of the blended line
// (numpy.gradient, airplane.py:628/640) is served by warp shuffles.  The warp walks down a chunk
// of control points i whose row constants sit in shared memory (broadcast reads).  Vortex-node
// data of lane j stay in registers for the whole chunk.  Output rows are component-planar
// V[i][c][j], so every store instruction writes one contiguous run per row.
//
// FP64-pipe bound by design: 24 B written per pair against ~0.5 kflop of DFMA-class work.
#include <cfloat>
#include <cstdlib>
#include "lls_internal.h"
#include "section.cuh"

namespace lls {

constexpr int INF_ROWS = 64;      // control points per CTA chunk
constexpr int INF_WARPS = 4;      // j-windows per CTA
constexpr int INF_OUT = 30;       // written lanes per window

struct RowSmem {
    double pc[3][INF_ROWS];
    double pcd[3][INF_ROWS];
    double uau[3][INF_ROWS];
    double pos[3][INF_ROWS];
    double sigma[INF_ROWS];
    double span[INF_ROWS];
    int wb[INF_ROWS];
    int we[INF_ROWS];
    int reid[INF_ROWS];
    int ac[INF_ROWS];
};

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ Vec3 shfl_v(const Vec3& v, int src) { return Vec3{shfl_d(v.x, src), shfl_d(v.y, src), shfl_d(v.z, src)}; }

#ifndef INF_FAST
#define INF_FAST 1     // 0: IEEE divisions and FP64-pipe nan_to_num as first written (kept for A/B timing, scripts/influence_bench.py)
#endif

// 1 / d to about 1 ulp: hardware seed (reciprocal of the high word, ~2^-20) and one cubic Newton step, i.e. 3 FP64-pipe
// instructions instead of the ~8 plus a guarded slow path of the IEEE division.  Zero, infinity and NaN come out as the seed gives
// them (inf, 0, NaN), as the division would; the test for that is on the integer pipe.
__device__ __forceinline__ double rcp_fast(double d) {
#if INF_FAST
    double x0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x0) : "d"(d));
    const double e = fma(-d, x0, 1.0);
    const double q = fma(e, e, e);
    const double r = fma(x0, q, x0);
    return ((__double2hiint(e) & 0x7ff00000) == 0x7ff00000) ? x0 : r;
#else
    return 1.0 / d;
#endif
}

// finite vortex segment from a to b seen from the control point (ra = PC - a, rb = PC - b):
//   (|ra|+|rb|) (ra x rb) / (|ra||rb| (|ra||rb| + ra.rb))          scene.py:525-527, 531-533, 536-538
__device__ __forceinline__ Vec3 segment_term(const Vec3& ra, double ram, const Vec3& rb, double rbm) {
    const double mm = ram * rbm;
    const double inv = rcp_fast(mm * (mm + dot(ra, rb)));
    return ((ram + rbm) * inv) * cross(ra, rb);
}

// nan_to_num of the reference (scene.py:624, 630): NaN -> 0, +-inf -> +-DBL_MAX
__device__ __forceinline__ double nan_to_num(double v) {
#if INF_FAST
    if ((__double2hiint(v) & 0x7ff00000) != 0x7ff00000) return v;      // finite: integer pipe only
#endif
    return (v != v) ? 0.0 : fmin(fmax(v, -DBL_MAX), DBL_MAX);
}

// semi-infinite trailing leg from the joint along u; sign = -1 for the inbound leg, +1 for the outbound leg
// `guard` = max(impingement threshold, 1e-13): one comparison settles both tests of the reference for every ordinary pair
__device__ __forceinline__ Vec3 trailing_term(const Vec3& u, const Vec3& rj, double rjm, double sign, double thr, double guard, bool& imp) {
    const double d = rjm * (rjm - dot(u, rj));                                  // scene.py:621, 627
#if INF_FAST
    if (!(d >= guard)) {
        imp |= (fabs(d) < thr);
        if (!(d > 1e-13)) return Vec3{0.0, 0.0, 0.0};
    }
#else
    imp |= (fabs(d) < thr);
    if (!(d > 1e-13)) return Vec3{0.0, 0.0, 0.0};
#endif
    const double t = sign * rcp_fast(d);
    const Vec3 c = cross(u, rj);
    return Vec3{nan_to_num(c.x * t), nan_to_num(c.y * t), nan_to_num(c.z * t)};
}

// one point of the blended lifting line seen from control point i (airplane.py:587-596)
__device__ __forceinline__ Vec3 blend_point(const Vec3& pc, const Vec3& pcd, double sigma, double span_i, const Vec3& node, double span_node) {
    const double ds = span_node - span_i;
    const double b = exp(-sigma * ds * ds);
    const Vec3 straight = pc + ds * pcd;
    return b * straight + (1.0 - b) * node;
}

// tangent direction (unnormalised) of numpy.gradient's three-point formulas, multiplied through by the
// positive common denominator:  interior  hs^2 d_next + hd^2 d_prev
__device__ __forceinline__ Vec3 tangent_interior(const Vec3& d_prev, double hs, const Vec3& d_next, double hd) {
    return (hs * hs) * d_next + (hd * hd) * d_prev;
}
//   left end   h1 (2 h0 + h1) d01 - h0^2 d12
__device__ __forceinline__ Vec3 tangent_left(const Vec3& d01, double h0, const Vec3& d12, double h1) {
    return (h1 * (2.0 * h0 + h1)) * d01 - (h0 * h0) * d12;
}
//   right end  hs (2 hd + hs) d_last - hd^2 d_prev
__device__ __forceinline__ Vec3 tangent_right(const Vec3& d_prev, double hs, const Vec3& d_last, double hd) {
    return (hs * (2.0 * hd + hs)) * d_last - (hd * hd) * d_prev;
}

// joint corner: u_j = normalise(u_a - (T.u_a) T), T the unit tangent (airplane.py:629-635; the c1 scale cancels)
__device__ __forceinline__ Vec3 joint_corner(const Vec3& Pe, const Vec3& tangent, const Vec3& ua, double cdj) {
    // u_j is parallel to u_a - (T.u_a) T with T = t/|t|; multiplied through by |t|^2 this is |t|^2 u_a - (t.u_a) t,
    // so one reciprocal square root (of the result) is enough
    const Vec3 uj = dot(tangent, tangent) * ua - dot(tangent, ua) * tangent;
    return Pe + (cdj * rsqrt(dot(uj, uj))) * uj;
}

#ifndef INF_UNROLL
#define INF_UNROLL 1
#endif
constexpr int kInfUnroll = INF_UNROLL;   // rows per trip of the control-point loop; 2 spills more and is slower (0.543 -> 0.595 ms at C2)
#ifndef INF_MIN_BLOCKS
#define INF_MIN_BLOCKS 4   // 128 registers: 16 warps per SM; measured 0.61 ms vs 0.76 ms at 2 blocks (C2), slower again at 5-6 (spills)
#endif
#ifndef INF_PLAIN_MIN_BLOCKS
#define INF_PLAIN_MIN_BLOCKS 5   // the plain specialisation carries no blend state: 102 registers -> 20 warps per SM
#endif
// Two specialisations share the grid (INF_SPLIT, the default).  A tile (64 control points x 4 windows of vortices) whose rows and
// vortices cannot lie on a common wing -- wings are contiguous index ranges, so that is an interval test on two table entries --
// is PLAIN: cross-wing / cross-aircraft pairs, about two thirds of a refined aeroplane and nearly all of a swarm.  Its kernel has no
// blend code at all, hence no blend registers: more resident warps for the dependent FP64 chains of the reciprocals and square roots.
// The other tiles take the BLEND kernel (the general one).  Each launch covers the whole grid and a CTA whose tile belongs to the other
// specialisation returns at once.  MODE: 0 = general kernel for every tile (INF_SPLIT=0), 1 = blend tiles only, 2 = plain tiles only.
template <int MODE>
__global__ void __launch_bounds__(INF_WARPS * 32, MODE == 2 ? INF_PLAIN_MIN_BLOCKS : INF_MIN_BLOCKS) influence_kernel(
    int n, int ld, const double* __restrict__ tab, const int32_t* __restrict__ itab, const double* __restrict__ ac,
    const double* __restrict__ out, double* __restrict__ V, double thr, int* __restrict__ status, CaseStrides cs) {
    tab = case_ptr(tab, cs.tab);
    ac = case_ptr(ac, cs.ac);
    out = case_ptr(out, cs.out);
    V = case_ptr(V, cs.V);
    status = case_work_ptr(status, cs.work);
    __shared__ RowSmem rs;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    static_assert(INF_ROWS == 64, "a chunk of control points is one 64-row block of the row partition");
    const int lrow0 = blockIdx.y * INF_ROWS;          // first row of the chunk in this rank's V
    const int row0 = global_row(cs, lrow0);           // ... and in the scene
    const int nrows = min(INF_ROWS, n - row0);
    if (nrows <= 0) return;
    if (MODE != 0) {
        // wings spanned by the rows of this chunk: [wing begin of the first row, wing end of the last row); vortices of this CTA:
        // [first window's first lane, last window's last lane]
        const int wlo = itab[(size_t)LLS_I_WING_BEGIN * ld + row0], whi = itab[(size_t)LLS_I_WING_END * ld + row0 + nrows - 1];
        const int jlo = blockIdx.x * INF_WARPS * INF_OUT - 1, jhi = (blockIdx.x * INF_WARPS + INF_WARPS) * INF_OUT;
        const bool may_blend = (jlo < whi) && (jhi >= wlo);
        if (may_blend != (MODE == 1)) return;
    }

    // stage the row constants of this chunk
    for (int r = threadIdx.x; r < nrows; r += blockDim.x) {
        const int i = row0 + r;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            rs.pc[c][r] = tab[(size_t)(LLS_F_PC + c) * ld + i];
            rs.pcd[c][r] = tab[(size_t)(LLS_F_PCD + c) * ld + i];
            rs.uau[c][r] = tab[(size_t)(LLS_F_UAU + c) * ld + i];
        }
        rs.sigma[r] = tab[(size_t)LLS_F_SIGMA * ld + i];
        rs.span[r] = tab[(size_t)LLS_F_SPC * ld + i];
        rs.wb[r] = itab[(size_t)LLS_I_WING_BEGIN * ld + i];
        rs.we[r] = itab[(size_t)LLS_I_WING_END * ld + i];
        rs.reid[r] = itab[(size_t)LLS_I_REID * ld + i];
        const int a = itab[(size_t)LLS_I_AIRCRAFT * ld + i];
        rs.ac[r] = a;
#pragma unroll
        for (int c = 0; c < 3; ++c) rs.pos[c][r] = ac[(size_t)a * LLS_AC_STRIDE + LLS_AC_POS + c];
    }
    __syncthreads();

    // this lane's vortex
    const int window = blockIdx.x * INF_WARPS + warp;
    const int j_raw = window * INF_OUT - 1 + lane;
    const bool j_real = (j_raw >= 0) && (j_raw < n);
    const bool writes = (lane >= 1) && (lane <= INF_OUT) && (j_raw < ld) && (j_raw >= 0);
    if (window * INF_OUT >= ld) return;  // whole warp beyond the padded width
    const int j = min(max(j_raw, 0), n - 1);

    const Vec3 P0 = ld3(tab, ld, LLS_F_P0, j), P1 = ld3(tab, ld, LLS_F_P1, j);
    const Vec3 P0J = ld3(tab, ld, LLS_F_P0J, j), P1J = ld3(tab, ld, LLS_F_P1J, j);
    const Vec3 uau_j = ld3(tab, ld, LLS_F_UAU, j);
    const Vec3 u0 = ld3(out, ld, LLS_O_UTRAIL0, j), u1 = ld3(out, ld, LLS_O_UTRAIL1, j);
    const double s0 = tab[(size_t)LLS_F_SP0 * ld + j], s1 = tab[(size_t)LLS_F_SP1 * ld + j], sc = tab[(size_t)LLS_F_SPC * ld + j];
    const double cdj0 = tab[(size_t)LLS_F_CDJ0 * ld + j], cdj1 = tab[(size_t)LLS_F_CDJ1 * ld + j];
    const int wb_j = j_real ? itab[(size_t)LLS_I_WING_BEGIN * ld + j] : -2;
    const int ac_j = itab[(size_t)LLS_I_AIRCRAFT * ld + j];
    const Vec3 pos_j{ac[(size_t)ac_j * LLS_AC_STRIDE + LLS_AC_POS], ac[(size_t)ac_j * LLS_AC_STRIDE + LLS_AC_POS + 1],
                     ac[(size_t)ac_j * LLS_AC_STRIDE + LLS_AC_POS + 2]};

    bool imp = false;
    const double guard = fmax(thr, 1e-13);
    const double inv4pi = 0.25 / 3.14159265358979323846;

#pragma unroll kInfUnroll
    for (int r = 0; r < nrows; ++r) {
        const int i = row0 + r;
        const Vec3 pc{rs.pc[0][r], rs.pc[1][r], rs.pc[2][r]};
        const int wb_i = rs.wb[r];
        const bool reid_i = rs.reid[r] != 0;
        const bool same_ac = (rs.ac[r] == ac_j);
        const bool blended = reid_i && (wb_i == wb_j);

        // plain geometry (other wing / other aircraft): scene.py:469-475, airplane.py:530-531, 555-558
        Vec3 n0 = P0, n1 = P1, q0 = P0J, q1 = P1J;
        if (same_ac && !reid_i) {  // airplane.py:665-669: joints collapse onto the nodes
            q0 = P0;
            q1 = P1;
        }

        if (MODE != 2 && __any_sync(0xffffffffu, blended)) {
            const Vec3 pcd{rs.pcd[0][r], rs.pcd[1][r], rs.pcd[2][r]};
            const double sigma = rs.sigma[r], span_i = rs.span[r];
            const int we_i = rs.we[r];
            // blended nodes of this lane (airplane.py:587-596)
            const Vec3 e0 = blend_point(pc, pcd, sigma, span_i, P0, s0);
            const Vec3 e1 = blend_point(pc, pcd, sigma, span_i, P1, s1);
            // blended unswept axial vector (airplane.py:599-603)
            const double dsc = sc - span_i;
            const double bc = exp(-sigma * dsc * dsc);
            Vec3 ua = bc * Vec3{rs.uau[0][r], rs.uau[1][r], rs.uau[2][r]} + (1.0 - bc) * uau_j;
            ua = rsqrt(dot(ua, ua)) * ua;
            // forward differences along both lines, neighbours by shuffle (airplane.py:626-627, 638-639)
            const Vec3 dn0 = shfl_v(e0, lane + 1) - e0;
            const Vec3 dn1 = shfl_v(e1, lane + 1) - e1;
            const double hn0 = norm(dn0), hn1 = norm(dn1);
            const Vec3 dp0 = shfl_v(dn0, lane - 1), dp1 = shfl_v(dn1, lane - 1);
            const double hp0 = shfl_d(hn0, lane - 1), hp1 = shfl_d(hn1, lane - 1);
            Vec3 t0 = tangent_interior(dp0, hp0, dn0, hn0);
            Vec3 t1 = tangent_interior(dp1, hp1, dn1, hn1);
            const bool left_edge = blended && (j == wb_i);
            const bool right_edge = blended && (j == we_i - 1);
            if (__any_sync(0xffffffffu, left_edge || right_edge)) {
                // one-sided second-order ends (numpy.gradient edge_order=2): evaluate the two further points directly
                if (left_edge) {
                    const int ja = min(j + 1, n - 1), jb = min(j + 2, n - 1);
                    const Vec3 a0 = blend_point(pc, pcd, sigma, span_i, ld3(tab, ld, LLS_F_P0, ja), tab[(size_t)LLS_F_SP0 * ld + ja]);
                    const Vec3 b0 = blend_point(pc, pcd, sigma, span_i, ld3(tab, ld, LLS_F_P0, jb), tab[(size_t)LLS_F_SP0 * ld + jb]);
                    const Vec3 a1 = blend_point(pc, pcd, sigma, span_i, ld3(tab, ld, LLS_F_P1, ja), tab[(size_t)LLS_F_SP1 * ld + ja]);
                    const Vec3 b1 = blend_point(pc, pcd, sigma, span_i, ld3(tab, ld, LLS_F_P1, jb), tab[(size_t)LLS_F_SP1 * ld + jb]);
                    const Vec3 d01_0 = a0 - e0, d12_0 = b0 - a0, d01_1 = a1 - e1, d12_1 = b1 - a1;
                    t0 = tangent_left(d01_0, norm(d01_0), d12_0, norm(d12_0));
                    t1 = tangent_left(d01_1, norm(d01_1), d12_1, norm(d12_1));
                } else if (right_edge) {
                    const int ja = max(j - 1, 0), jb = max(j - 2, 0);
                    const Vec3 a0 = blend_point(pc, pcd, sigma, span_i, ld3(tab, ld, LLS_F_P0, ja), tab[(size_t)LLS_F_SP0 * ld + ja]);
                    const Vec3 b0 = blend_point(pc, pcd, sigma, span_i, ld3(tab, ld, LLS_F_P0, jb), tab[(size_t)LLS_F_SP0 * ld + jb]);
                    const Vec3 a1 = blend_point(pc, pcd, sigma, span_i, ld3(tab, ld, LLS_F_P1, ja), tab[(size_t)LLS_F_SP1 * ld + ja]);
                    const Vec3 b1 = blend_point(pc, pcd, sigma, span_i, ld3(tab, ld, LLS_F_P1, jb), tab[(size_t)LLS_F_SP1 * ld + jb]);
                    const Vec3 dl_0 = e0 - a0, dp_0 = a0 - b0, dl_1 = e1 - a1, dp_1 = a1 - b1;
                    t0 = tangent_right(dp_0, norm(dp_0), dl_0, norm(dl_0));
                    t1 = tangent_right(dp_1, norm(dp_1), dl_1, norm(dl_1));
                }
            }
            if (blended) {
                n0 = e0;
                n1 = e1;
                q0 = joint_corner(e0, t0, ua, cdj0);                              // airplane.py:635
                q1 = joint_corner(e1, t1, ua, cdj1);                              // airplane.py:647
            }
        }

        // vectors from the four corners to the control point (airplane.py:672-675, scene.py:501-504); the offset between the
        // two aircraft origins is exactly zero inside one aircraft, so same-aircraft differences keep every digit
        const Vec3 pcs{pc.x + (rs.pos[0][r] - pos_j.x), pc.y + (rs.pos[1][r] - pos_j.y), pc.z + (rs.pos[2][r] - pos_j.z)};
        const Vec3 r0 = pcs - n0, r1 = pcs - n1, r0j = pcs - q0, r1j = pcs - q1;
        const double r0m = norm(r0), r1m = norm(r1), r0jm = norm(r0j), r1jm = norm(r1j);

        Vec3 v = segment_term(r0j, r0jm, r0, r0m) + segment_term(r1, r1m, r1j, r1jm);   // joints, scene.py:531-538
        if (i != j_raw) v = v + segment_term(r0, r0m, r1, r1m);                       // bound, zero on the diagonal (scene.py:528)
        v = v + trailing_term(u0, r0j, r0jm, -1.0, thr, guard, imp) + trailing_term(u1, r1j, r1jm, 1.0, thr, guard, imp);
        if (!j_real) v = Vec3{0.0, 0.0, 0.0};
        if (writes) {
            double* row = V + (size_t)(lrow0 + r) * 3 * ld + j_raw;
            row[0] = inv4pi * v.x;
            row[(size_t)ld] = inv4pi * v.y;
            row[(size_t)2 * ld] = inv4pi * v.z;
        }
    }
    if (imp && j_real && lane >= 1 && lane <= INF_OUT) atomicOr(status, LLS_STATUS_IMPINGEMENT);
}

int launch_influence(Scene* s, double impingement_threshold, cudaStream_t st) {
    const int windows = (s->ld + INF_OUT - 1) / INF_OUT;
    dim3 grid((windows + INF_WARPS - 1) / INF_WARPS, s->n_loc_blocks, s->ncases);
    static int split = -1;      // INF_SPLIT=0: one general kernel for every tile
    if (split < 0) {
        const char* env = getenv("INF_SPLIT");
        split = (env != nullptr && env[0] == '0') ? 0 : 1;
    }
    if (split) {
        influence_kernel<2><<<grid, INF_WARPS * 32, 0, st>>>(s->n, s->ld, s->tab, s->itab, s->ac, s->out, s->V, impingement_threshold, s->status,
                                                             strides(s, false));
        influence_kernel<1><<<grid, INF_WARPS * 32, 0, st>>>(s->n, s->ld, s->tab, s->itab, s->ac, s->out, s->V, impingement_threshold, s->status,
                                                             strides(s, false));
        count_launch(2);
    } else {
        influence_kernel<0><<<grid, INF_WARPS * 32, 0, st>>>(s->n, s->ld, s->tab, s->itab, s->ac, s->out, s->V, impingement_threshold, s->status,
                                                             strides(s, false));
        count_launch();
    }
    LLS_CUDA_TRY(cudaGetLastError());
    return LLS_OK;
}

}  // namespace lls
