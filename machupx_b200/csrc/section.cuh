// Device section model: the airfoil_db "linear" airfoil (third-party dependency of the reference,
// airfoil_db>=1.4.3, setup.py:13), evaluated per control point inside the Newton loop.
// Two bracketing airfoils blended linearly in span (wing_segment.py:1062-1072, 1104-1124).
// Host mirror: machupx_b200/airfoil.py; oracle mirror: oracle/lls_oracle.py section_*().
#pragma once
#include "lls_internal.h"

namespace lls {

struct SectionParams {
    const double* a;  // airfoil A block
    const double* b;  // airfoil B block
    double wb;        // blend weight toward B
    double flap_da;   // delta * eps
    double flap_dcm;
    double flap_dcd;
};

__device__ __forceinline__ SectionParams load_section(const double* __restrict__ tab, const int32_t* __restrict__ itab,
                                                      const double* __restrict__ af, int ld, int i) {
    SectionParams p;
    p.a = af + (size_t)itab[(size_t)LLS_I_AFA * ld + i] * LLS_AF_STRIDE;
    p.b = af + (size_t)itab[(size_t)LLS_I_AFB * ld + i] * LLS_AF_STRIDE;
    p.wb = tab[(size_t)LLS_F_WB * ld + i];
    p.flap_da = tab[(size_t)LLS_F_FLAP_DA * ld + i];
    p.flap_dcm = tab[(size_t)LLS_F_FLAP_DCM * ld + i];
    p.flap_dcd = tab[(size_t)LLS_F_FLAP_DCD * ld + i];
    return p;
}

__device__ __forceinline__ double clipd(double x, double m) { return fmin(fmax(x, -m), m); }

__device__ __forceinline__ double blendd(double wb, double fa, double fb) { return (1.0 - wb) * fa + wb * fb; }

// CL = clip(CLa (alpha - aL0 + delta eps), +-CL_max)
__device__ __forceinline__ double section_CL(const SectionParams& p, double alpha) {
    double ca = clipd(p.a[LLS_AF_CLA] * (alpha - p.a[LLS_AF_AL0] + p.flap_da), p.a[LLS_AF_CLMAX]);
    double cb = clipd(p.b[LLS_AF_CLA] * (alpha - p.b[LLS_AF_AL0] + p.flap_da), p.b[LLS_AF_CLMAX]);
    return blendd(p.wb, ca, cb);
}

__device__ __forceinline__ double section_CLa(const SectionParams& p) {
    return blendd(p.wb, p.a[LLS_AF_CLA], p.b[LLS_AF_CLA]);
}

__device__ __forceinline__ double section_aL0(const SectionParams& p) {
    return blendd(p.wb, p.a[LLS_AF_AL0] - p.flap_da, p.b[LLS_AF_AL0] - p.flap_da);
}

// CD = CD0 + CD1 CL + CD2 CL^2 (CL without flap) + 0.002 |delta| 180/pi
__device__ __forceinline__ double section_CD(const SectionParams& p, double alpha) {
    double cla = clipd(p.a[LLS_AF_CLA] * (alpha - p.a[LLS_AF_AL0]), p.a[LLS_AF_CLMAX]);
    double clb = clipd(p.b[LLS_AF_CLA] * (alpha - p.b[LLS_AF_AL0]), p.b[LLS_AF_CLMAX]);
    double cda = p.a[LLS_AF_CD0] + p.a[LLS_AF_CD1] * cla + p.a[LLS_AF_CD2] * cla * cla + p.flap_dcd;
    double cdb = p.b[LLS_AF_CD0] + p.b[LLS_AF_CD1] * clb + p.b[LLS_AF_CD2] * clb * clb + p.flap_dcd;
    return blendd(p.wb, cda, cdb);
}

// Cm = Cma (alpha - aL0) + CmL0 + delta dCm
__device__ __forceinline__ double section_Cm(const SectionParams& p, double alpha) {
    double ma = p.a[LLS_AF_CMA] * (alpha - p.a[LLS_AF_AL0]) + p.a[LLS_AF_CML0] + p.flap_dcm;
    double mb = p.b[LLS_AF_CMA] * (alpha - p.b[LLS_AF_AL0]) + p.b[LLS_AF_CML0] + p.flap_dcm;
    return blendd(p.wb, ma, mb);
}

// linear airfoils have no Reynolds / Mach dependence
__device__ __forceinline__ double section_CLRe(const SectionParams&, double, double, double) { return 0.0; }
__device__ __forceinline__ double section_CLM(const SectionParams&, double, double, double) { return 0.0; }

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
