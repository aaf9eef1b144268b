// "database" airfoils (airfoil_db's table look-up type; third-party dependency of the reference, airfoil_db>=1.4.3, setup.py:13;
// call sites wing_segment.py:1081-1085, 1111-1120): CL, CD, Cm tabulated over any of (alpha, Rey, Mach, flap deflection, flap
// fraction), interpolated barycentrically on the Delaunay triangulation of the data points in variables scaled to [0, 1] (what
// scipy.interpolate.griddata(method="linear") does inside airfoil_db), CLa / CLRe / CLM as central differences of that
// interpolant (steps 0.001 rad / 1000 / 0.05), aL0 as its root in alpha (secant).  PINNED by the reference's golden of
// test/scene_tests/test_solve_forces.py:453-494 (tests/test_database_airfoils.py).  The triangulation is an input: the host
// builds it with scipy (machupx_b200/airfoil.py) and ships it as the pool record described in include/lls.h.
//
// Plain C++ without CUDA intrinsics on purpose: tests/test_db_airfoil_source.py compiles THIS file with g++ and checks it against
// the numpy restatement of the oracle, so the arithmetic below is verified on CPU as well as on the device.
#pragma once
#include <cmath>
#include <cstddef>

#ifdef __CUDACC__
#define LLS_HD __host__ __device__
#else
#define LLS_HD
#endif

namespace lls {

constexpr double DB_EPS = 1e-12;   // barycentric tolerance of the point location (normalised variables are O(1))
enum { DB_CL = 0, DB_CD = 1, DB_CM = 2, DB_CLA = 3, DB_CLRE = 4, DB_CLM = 5, DB_AL0 = 6 };

struct DbTable {
    int V, P, S, w;       // variables, points, simplices, doubles per simplex record
    const double* lim;    // V x (lower limit, span)
    const double* vals;   // P x (CL, CD, Cm)
    const double* simp;   // S x [V+1 vertex ids | V x V map to barycentric coordinates | V origin (= last vertex)]
};

LLS_HD inline DbTable db_table(const double* rec, int V) {
    DbTable t;
    t.V = V;
    t.P = (int)rec[0];
    t.S = (int)rec[1];
    t.w = V + 1 + V * V + V;
    t.lim = rec + 2;
    t.vals = t.lim + 2 * V;
    t.simp = t.vals + 3 * (size_t)t.P;
    return t;
}

// barycentric coordinates c[0..V] of the normalised point xn in simplex s; returns the smallest (-1e300 for a flat simplex,
// whose map the host fills with NaN)
LLS_HD inline double db_bary(const DbTable& t, int s, const double* xn, double* c) {
    const double* T = t.simp + (size_t)s * t.w + t.V + 1;
    const double* r = T + t.V * t.V;
    double sum = 0.0;
    for (int i = 0; i < t.V; ++i) {
        double ci = 0.0;
        for (int j = 0; j < t.V; ++j) ci += T[i * t.V + j] * (xn[j] - r[j]);
        c[i] = ci;
        sum += ci;
    }
    c[t.V] = 1.0 - sum;
    double cmin = 1e300;
    for (int i = 0; i <= t.V; ++i) {
        if (!(c[i] == c[i])) return -1e300;
        cmin = c[i] < cmin ? c[i] : cmin;
    }
    return cmin;
}

// coefficient `which` (DB_CL / DB_CD / DB_CM) at the raw variables x[0..V).  hint: simplex tried first (in) / used (out).
// A point outside the table sets oob and extrapolates from the simplex whose smallest coordinate is largest.
LLS_HD inline double db_lookup(const DbTable& t, int which, const double* x, int& hint, bool& oob) {
    double xn[5], c[6];
    for (int v = 0; v < t.V; ++v) xn[v] = (x[v] - t.lim[2 * v]) / t.lim[2 * v + 1];
    int found = -1;
    if (hint >= 0 && hint < t.S && db_bary(t, hint, xn, c) >= -DB_EPS) found = hint;
    if (found < 0) {
        double best = -2e300;
        int sbest = 0;
        for (int s = 0; s < t.S; ++s) {
            const double m = db_bary(t, s, xn, c);
            if (m >= -DB_EPS) {
                found = s;
                break;
            }
            if (m > best) {
                best = m;
                sbest = s;
            }
        }
        if (found < 0) {
            oob = true;
            found = sbest;
            db_bary(t, found, xn, c);
        }
    }
    hint = found;
    const double* verts = t.simp + (size_t)found * t.w;
    double out = 0.0;
    for (int i = 0; i <= t.V; ++i) out += c[i] * t.vals[3 * (size_t)(int)verts[i] + which];
    return out;
}

// central difference of CL in variable v with the two samples kept inside the table's limits (one-sided at its edges)
LLS_HD inline double db_slope(const DbTable& t, int v, double step, double* x, int& hint) {
    const double lo = t.lim[2 * v], hi = lo + t.lim[2 * v + 1], x0 = x[v];
    const double up = fmin(x0 + step, fmax(hi, x0)) - x0;
    const double dn = fmax(x0 - step, fmin(lo, x0)) - x0;
    bool ignored = false;
    x[v] = x0 + up;
    const double f1 = db_lookup(t, DB_CL, x, hint, ignored);
    x[v] = x0 + dn;
    const double f0 = db_lookup(t, DB_CL, x, hint, ignored);
    x[v] = x0;
    return (f1 - f0) / (up - dn);
}

// root of CL in alpha (variable va): secant from (0, 0.01) until two iterates agree to 1e-10, 50 steps at most
LLS_HD inline double db_root(const DbTable& t, int va, double* x, int& hint) {
    bool ignored = false;
    double a0 = 0.0, a1 = 0.01;
    x[va] = a0;
    double f0 = db_lookup(t, DB_CL, x, hint, ignored);
    x[va] = a1;
    double f1 = db_lookup(t, DB_CL, x, hint, ignored);
    for (int it = 0; it < 50; ++it) {
        if (!(fabs(a1 - a0) > 1e-10)) break;
        const double a2 = a1 - f1 * (a0 - a1) / (f0 - f1);
        x[va] = a2;
        const double f2 = db_lookup(t, DB_CL, x, hint, ignored);
        a0 = a1;
        f0 = f1;
        a1 = a2;
        f1 = f2;
    }
    return a1;
}

struct DbResult {
    double value;
    int oob;
};

// One quantity of one database airfoil at one control point.  rec: pool record; kinds: the V variable codes of the airfoil block
// (0 alpha, 1 Rey, 2 Mach, 3 flap deflection, 4 flap fraction).
LLS_HD inline DbResult db_evaluate(const double* rec, int V, const double* kinds, double alpha, double Re, double Mach,
                                   double defl, double cf, int what) {
    const DbTable t = db_table(rec, V);
    double x[5];
    int where[5] = {-1, -1, -1, -1, -1};   // position of each kind among the table's variables
    for (int v = 0; v < V; ++v) {
        const int kind = (int)kinds[v];
        x[v] = kind == 0 ? alpha : kind == 1 ? Re : kind == 2 ? Mach : kind == 3 ? defl : cf;
        where[kind] = v;
    }
    DbResult r;
    r.oob = 0;
    int hint = -1;
    if (what <= DB_CM) {
        bool oob = false;
        r.value = db_lookup(t, what, x, hint, oob);
        r.oob = oob ? 1 : 0;
    } else if (what == DB_AL0) {
        r.value = where[0] < 0 ? 0.0 : db_root(t, where[0], x, hint);
    } else {
        const int dkind = what - DB_CLA;
        const double step = dkind == 0 ? 0.001 : dkind == 1 ? 1000.0 : 0.05;
        r.value = where[dkind] < 0 ? 0.0 : db_slope(t, where[dkind], step, x, hint);
    }
    return r;
}

}  // namespace lls
