// TEST INFRASTRUCTURE: compiles the device source of the database-airfoil model (machupx_b200/csrc/db_airfoil.cuh, plain C++ by
// design) for the HOST, so tests/test_db_airfoil_source.py can check the very code the kernels run against the numpy restatement
// of the oracle without a GPU.  Never part of liblls.so.
#include "../../machupx_b200/csrc/db_airfoil.cuh"

extern "C" void db_host_evaluate(const double* rec, int V, const double* kinds, int n, const double* alpha, const double* Re,
                                 const double* Mach, const double* defl, const double* cf, int what, double* out, int* oob) {
    for (int i = 0; i < n; ++i) {
        const lls::DbResult r = lls::db_evaluate(rec, V, kinds, alpha[i], Re[i], Mach[i], defl[i], cf[i], what);
        out[i] = r.value;
        oob[i] = r.oob;
    }
}
