// Device section model: the airfoil_db "linear", "poly_fit" and "database" airfoils (third-party dependency of the reference,
// airfoil_db>=1.4.3, setup.py:13), evaluated per control point inside the Newton loop.
// Two bracketing airfoils blended linearly in span (wing_segment.py:1062-1072, 1104-1124).
// Host mirror: machupx_b200/airfoil.py; oracle mirror: oracle/lls_oracle.py section_*().
#pragma once
#include "lls_internal.h"
#include "db_airfoil.cuh"

namespace lls {

struct SectionParams {
    const double* base;   // start of the airfoil buffer (blocks, then the poly_fit pool)
    const double* a;      // airfoil A block
    const double* b;      // airfoil B block
    double wb;            // blend weight toward B
    double flap_da;       // delta * eps           (linear airfoils)
    double flap_dcm;
    double flap_dcd;
    double defl, cf;      // raw flap deflection [rad] and chord fraction (poly_fit variables)
};

__device__ __forceinline__ SectionParams load_section(const double* __restrict__ tab, const int32_t* __restrict__ itab,
                                                      const double* __restrict__ af, int ld, int i) {
    SectionParams p;
    p.base = af;
    p.a = af + (size_t)itab[(size_t)LLS_I_AFA * ld + i] * LLS_AF_STRIDE;
    p.b = af + (size_t)itab[(size_t)LLS_I_AFB * ld + i] * LLS_AF_STRIDE;
    p.wb = tab[(size_t)LLS_F_WB * ld + i];
    p.flap_da = tab[(size_t)LLS_F_FLAP_DA * ld + i];
    p.flap_dcm = tab[(size_t)LLS_F_FLAP_DCM * ld + i];
    p.flap_dcd = tab[(size_t)LLS_F_FLAP_DCD * ld + i];
    p.defl = tab[(size_t)LLS_F_FLAP_DEFL * ld + i];
    p.cf = tab[(size_t)LLS_F_FLAP_CF * ld + i];
    return p;
}

__device__ __forceinline__ double clipd(double x, double m) { return fmin(fmax(x, -m), m); }

__device__ __forceinline__ double blendd(double wb, double fa, double fb) { return (1.0 - wb) * fa + wb * fb; }

// ---- poly_fit airfoils: multivariate polynomials in (alpha, Rey, Mach, flap deflection, flap fraction) ------------------
// Third-party model (airfoil_db "poly_fit" type, not in the reference tree): coefficient layout follows the fits MachUpX
// ships (dev/NACA_0012_fits.json, machupX/poly_fits.py:398-431): exponent tuple row-major, last variable fastest.
// CLa, CLRe, CLM are the analytic partial derivatives of the CL polynomial; aL0 is its root in alpha (Newton).
// PARITY UNPINNED: no reference test evaluates a poly_fit airfoil; pinned here only through the
// linear-equivalent fits of tests/test_polyfit_*.py.
struct PolyPoint {
    double x[5];   // value of every declared variable at this control point
    bool out_of_bounds;
};
enum { POLY_CL = 0, POLY_CD = 1, POLY_CM = 2 };

__device__ __forceinline__ bool is_poly(const double* blk) { return blk[LLS_AF_TYPE] > 0.5; }   // tested after is_db() where both occur
__device__ __forceinline__ bool is_db(const double* blk) { return blk[LLS_AF_TYPE] > 1.5; }

// ---- database airfoils (db_airfoil.cuh).  The block of a database airfoil uses the poly_fit fields (pool offset, V, variable
// codes).  Every kernel that evaluates the section model exists twice: the <DB = true> instantiations, launched for scenes with
// LLS_FLAG_DATABASE_AIRFOILS, carry the table walk (one out-of-line function); the <DB = false> ones are the kernels as they
// were, instruction for instruction, so the analytic models pay neither registers nor code for it.
static __device__ __noinline__ DbResult db_section(const double* base, const double* blk, double alpha, double Re, double Mach,
                                            double defl, double cf, int what) {
    return db_evaluate(base + (size_t)blk[LLS_AF_POLY_OFF], (int)blk[LLS_AF_POLY_NDOF], blk + LLS_AF_POLY_DOF0, alpha, Re, Mach,
                       defl, cf, what);
}
__device__ __forceinline__ double db_value(const SectionParams& p, const double* blk, double alpha, double Re, double Mach, int what,
                                           bool& oob) {
    const DbResult r = db_section(p.base, blk, alpha, Re, Mach, p.defl, p.cf, what);
    oob |= r.oob != 0;
    return r.value;
}

__device__ inline PolyPoint poly_point(const SectionParams& p, const double* blk, double alpha, double Re, double Mach) {
    PolyPoint pt;
    const int V = (int)blk[LLS_AF_POLY_NDOF];
    const double* lim = p.base + (size_t)blk[LLS_AF_POLY_OFF];
    pt.out_of_bounds = false;
    for (int v = 0; v < V; ++v) {
        const int kind = (int)blk[LLS_AF_POLY_DOF0 + v];
        const double x = kind == 0 ? alpha : kind == 1 ? Re : kind == 2 ? Mach : kind == 3 ? p.defl : p.cf;
        pt.x[v] = x;
        pt.out_of_bounds |= !(x >= lim[2 * v] && x <= lim[2 * v + 1]);
    }
    return pt;
}

// value (dvar < 0) or partial derivative with respect to the variable of KIND dvar (0 alpha, 1 Rey, 2 Mach) of polynomial `which`
__device__ inline double poly_eval(const SectionParams& p, const double* blk, const PolyPoint& pt, int which, int dkind) {
    const int V = (int)blk[LLS_AF_POLY_NDOF];
    const double* rec = p.base + (size_t)blk[LLS_AF_POLY_OFF] + 2 * V;
    for (int w = 0; w < which; ++w) rec += 1 + V + (int)rec[0];
    const int ncoef = (int)rec[0];
    const double* deg = rec + 1;
    const double* coef = rec + 1 + V;
    int dv = -1;
    if (dkind >= 0) {
        for (int v = 0; v < V; ++v)
            if ((int)blk[LLS_AF_POLY_DOF0 + v] == dkind) dv = v;
        if (dv < 0) return 0.0;        // the fit does not depend on that variable
    }
    double sum = 0.0;
    for (int j = 0; j < ncoef; ++j) {
        int rest = j;
        double term = coef[j];
        for (int v = V - 1; v >= 0; --v) {
            const int base = (int)deg[v] + 1;
            const int n = rest % base;
            rest /= base;
            if (v == dv) {
                if (n == 0) { term = 0.0; continue; }
                term *= (double)n;
                for (int e = 1; e < n; ++e) term *= pt.x[v];
            } else {
                for (int e = 0; e < n; ++e) term *= pt.x[v];
            }
        }
        sum += term;
    }
    return sum;
}

// one airfoil: CL, its alpha / Rey / Mach derivatives
template <bool DB>
__device__ inline double af_CL(const SectionParams& p, const double* blk, double alpha, double Re, double Mach, bool& oob) {
    if (DB && is_db(blk)) return db_value(p, blk, alpha, Re, Mach, DB_CL, oob);
    if (is_poly(blk)) {
        const PolyPoint pt = poly_point(p, blk, alpha, Re, Mach);
        oob |= pt.out_of_bounds;
        return poly_eval(p, blk, pt, POLY_CL, -1);
    }
    return clipd(blk[LLS_AF_CLA] * (alpha - blk[LLS_AF_AL0] + p.flap_da), blk[LLS_AF_CLMAX]);
}
template <bool DB>
__device__ inline double af_CL_d(const SectionParams& p, const double* blk, double alpha, double Re, double Mach, int dkind) {
    if (DB && is_db(blk)) {
        bool ignored = false;
        return db_value(p, blk, alpha, Re, Mach, DB_CLA + dkind, ignored);
    }
    if (is_poly(blk)) return poly_eval(p, blk, poly_point(p, blk, alpha, Re, Mach), POLY_CL, dkind);
    return dkind == 0 ? blk[LLS_AF_CLA] : 0.0;
}
template <bool DB>
__device__ inline double af_aL0(const SectionParams& p, const double* blk, double Re, double Mach) {
    if (DB && is_db(blk)) {
        bool ignored = false;
        return db_value(p, blk, 0.0, Re, Mach, DB_AL0, ignored);
    }
    if (!is_poly(blk)) return blk[LLS_AF_AL0] - p.flap_da;
    double a0 = 0.0;
    for (int it = 0; it < 30; ++it) {
        const PolyPoint pt = poly_point(p, blk, a0, Re, Mach);
        const double f = poly_eval(p, blk, pt, POLY_CL, -1), df = poly_eval(p, blk, pt, POLY_CL, 0);
        const double step = f / df;
        a0 -= step;
        if (!(fabs(step) > 1e-15)) break;
    }
    return a0;
}

// CL = clip(CLa (alpha - aL0 + delta eps), +-CL_max)   (linear)   |   polynomial (poly_fit)
template <bool DB>
__device__ __forceinline__ double section_CL(const SectionParams& p, double alpha, double Re, double Mach, bool& oob) {
    return blendd(p.wb, af_CL<DB>(p, p.a, alpha, Re, Mach, oob), af_CL<DB>(p, p.b, alpha, Re, Mach, oob));
}
template <bool DB>
__device__ __forceinline__ double section_CLa(const SectionParams& p, double alpha, double Re, double Mach) {
    return blendd(p.wb, af_CL_d<DB>(p, p.a, alpha, Re, Mach, 0), af_CL_d<DB>(p, p.b, alpha, Re, Mach, 0));
}
template <bool DB>
__device__ __forceinline__ double section_CLRe(const SectionParams& p, double alpha, double Re, double Mach) {
    return blendd(p.wb, af_CL_d<DB>(p, p.a, alpha, Re, Mach, 1), af_CL_d<DB>(p, p.b, alpha, Re, Mach, 1));
}
template <bool DB>
__device__ __forceinline__ double section_CLM(const SectionParams& p, double alpha, double Re, double Mach) {
    return blendd(p.wb, af_CL_d<DB>(p, p.a, alpha, Re, Mach, 2), af_CL_d<DB>(p, p.b, alpha, Re, Mach, 2));
}
template <bool DB>
__device__ __forceinline__ double section_aL0(const SectionParams& p, double Re, double Mach) {
    return blendd(p.wb, af_aL0<DB>(p, p.a, Re, Mach), af_aL0<DB>(p, p.b, Re, Mach));
}

// CD = CD0 + CD1 CL + CD2 CL^2 (CL without flap) + 0.002 |delta| 180/pi   (linear)
template <bool DB>
__device__ inline double af_CD(const SectionParams& p, const double* blk, double alpha, double Re, double Mach, bool& oob) {
    if (DB && is_db(blk)) return db_value(p, blk, alpha, Re, Mach, DB_CD, oob);
    if (is_poly(blk)) {
        const PolyPoint pt = poly_point(p, blk, alpha, Re, Mach);
        oob |= pt.out_of_bounds;
        return poly_eval(p, blk, pt, POLY_CD, -1);
    }
    const double cl = clipd(blk[LLS_AF_CLA] * (alpha - blk[LLS_AF_AL0]), blk[LLS_AF_CLMAX]);
    return blk[LLS_AF_CD0] + blk[LLS_AF_CD1] * cl + blk[LLS_AF_CD2] * cl * cl + p.flap_dcd;
}
template <bool DB>
__device__ __forceinline__ double section_CD(const SectionParams& p, double alpha, double Re, double Mach, bool& oob) {
    return blendd(p.wb, af_CD<DB>(p, p.a, alpha, Re, Mach, oob), af_CD<DB>(p, p.b, alpha, Re, Mach, oob));
}

// Cm = Cma (alpha - aL0) + CmL0 + delta dCm   (linear)
template <bool DB>
__device__ inline double af_Cm(const SectionParams& p, const double* blk, double alpha, double Re, double Mach, bool& oob) {
    if (DB && is_db(blk)) return db_value(p, blk, alpha, Re, Mach, DB_CM, oob);
    if (is_poly(blk)) {
        const PolyPoint pt = poly_point(p, blk, alpha, Re, Mach);
        oob |= pt.out_of_bounds;
        return poly_eval(p, blk, pt, POLY_CM, -1);
    }
    return blk[LLS_AF_CMA] * (alpha - blk[LLS_AF_AL0]) + blk[LLS_AF_CML0] + p.flap_dcm;
}
template <bool DB>
__device__ __forceinline__ double section_Cm(const SectionParams& p, double alpha, double Re, double Mach, bool& oob) {
    return blendd(p.wb, af_Cm<DB>(p, p.a, alpha, Re, Mach, oob), af_Cm<DB>(p, p.b, alpha, Re, Mach, oob));
}

struct Vec3 {
    double x, y, z;
};
__device__ __forceinline__ Vec3 ld3(const double* __restrict__ base, int ld, int f, int i) {
    return Vec3{base[(size_t)f * ld + i], base[(size_t)(f + 1) * ld + i], base[(size_t)(f + 2) * ld + i]};
}
__device__ __forceinline__ void st3(double* __restrict__ base, int ld, int f, int i, const Vec3& v) {
    base[(size_t)f * ld + i] = v.x;
    base[(size_t)(f + 1) * ld + i] = v.y;
    base[(size_t)(f + 2) * ld + i] = v.z;
}
__device__ __forceinline__ Vec3 operator+(const Vec3& a, const Vec3& b) { return Vec3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ Vec3 operator-(const Vec3& a, const Vec3& b) { return Vec3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ Vec3 operator*(double s, const Vec3& a) { return Vec3{s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ double dot(const Vec3& a, const Vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ Vec3 cross(const Vec3& a, const Vec3& b) {
    return Vec3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ double norm(const Vec3& a) { return sqrt(dot(a, a)); }

}  // namespace lls
