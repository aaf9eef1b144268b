"""ctypes binding of liblls (include/lls.h).  The product has NO CPU fallback: if the shared
library is missing or cannot be loaded this module raises, it never routes anywhere else."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblls.so")
ABI_VERSION = 3


class LlsError(RuntimeError):
    pass


class Sizes(ctypes.Structure):
    _fields_ = [("ld", ctypes.c_size_t), ("tab_bytes", ctypes.c_size_t), ("itab_bytes", ctypes.c_size_t),
                ("out_bytes", ctypes.c_size_t), ("vji_bytes", ctypes.c_size_t), ("mat_bytes", ctypes.c_size_t),
                ("work_bytes", ctypes.c_size_t)]


_lib = None

_vp, _i, _d = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
_pi, _pd = ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double)

# name -> (restype, argtypes); every symbol include/lls.h declares
SIGNATURES = {
    "lls_query_sizes": (_i, [_i, _i, ctypes.POINTER(Sizes)]),
    "lls_create": (_i, [ctypes.POINTER(_vp), _i, _i, _i, _i, _i]),
    "lls_destroy": (_i, [_vp]),
    "lls_bind": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_size_t]),
    "lls_bind_batch": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_size_t, _vp]),
    "lls_set_case_tables": (_i, [_vp, _vp, ctypes.c_size_t]),
    "lls_solve_nonlinear_batch": (_i, [_vp, _d, _d, _i, _vp, _pi, _pd, _pi]),
    "lls_get_status_batch": (_i, [_vp, _vp, _pi]),
    "lls_query_sizes_partitioned": (_i, [_i, _i, _i, ctypes.POINTER(Sizes)]),
    "lls_set_partition": (_i, [_vp, _i, _i]),
    "lls_dist_assemble": (_i, [_vp, _i, _vp]),
    "lls_dist_residual": (_i, [_vp, _vp, _pd]),
    "lls_dist_update_gamma": (_i, [_vp, _d, _vp]),
    "lls_dist_vectors": (_i, [_vp, ctypes.POINTER(_vp), ctypes.POINTER(_vp), ctypes.POINTER(_vp)]),
    "lls_dist_lu_begin": (_i, [_vp, ctypes.POINTER(ctypes.c_size_t), _vp]),
    "lls_dist_lu_step_doubles": (_i, [_vp, _i, ctypes.POINTER(ctypes.c_size_t)]),
    "lls_dist_lu_owner": (_i, [_vp, _i, _i, _vp, _vp]),
    "lls_dist_lu_update": (_i, [_vp, _i, _i, _vp, _vp]),
    "lls_dist_back": (_i, [_vp, _i, _vp]),
    "lls_dist_comm_create": (_i, [_vp, _i, _vp]),
    "lls_dist_comm_connect": (_i, [_vp, _vp]),
    "lls_dist_comm_bytes": (_i, [_vp, _i, ctypes.POINTER(ctypes.c_size_t)]),
    "lls_dist_comm_attach": (_i, [_vp, _i, ctypes.POINTER(_vp), _vp]),
    "lls_dist_comm_multicast": (_i, [_vp]),
    "lls_dist_lu_solve": (_i, [_vp, _vp]),
    "lls_dist_comm_stats": (_i, [_vp, _pd, _vp]),
    "lls_dist_comm_destroy": (_i, [_vp]),
    "lls_dist_debug_schedule": (_i, [_i, _i, _i, _i, _pi, _i]),
    "lls_dist_debug_role": (_i, [_i, _i, _i, _i, _i, _i, _i, _pi, _pi]),
    "lls_set_state": (_i, [_vp, _vp]),
    "lls_flow_state": (_i, [_vp, _vp]),
    "lls_build_influence": (_i, [_vp, _d, _vp]),
    "lls_solve_linear": (_i, [_vp, _vp]),
    "lls_solve_nonlinear": (_i, [_vp, _d, _d, _i, _vp, _pi, _pd, _pi, _pd]),
    "lls_integrate": (_i, [_vp, _vp, _vp]),
    "lls_get_status": (_i, [_vp, _vp, _pi]),
    "lls_status_offset": (_i, [_vp, ctypes.POINTER(ctypes.c_size_t)]),
    "lls_lu_factor": (_i, [_vp, _vp, _vp]),
    "lls_lu_solve": (_i, [_vp, _vp, _vp, _vp]),
    "lls_set_timing": (_i, [_vp, _i]),
    "lls_get_timings": (_i, [_vp, _pd]),
    "lls_probe_peaks": (_i, [_pd, _vp]),
    "lls_last_error": (ctypes.c_char_p, []),
    "lls_abi_version": (_i, []),
    "lls_launch_count": (ctypes.c_ulonglong, []),
}


def load():
    """Loads liblls.so (built in-tree by build.sh / __graft_entry__.build()) and checks its ABI version."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LlsError("liblls.so not found at {0}: build it with ./build.sh (nvcc, sm_100a). "
                       "machupx_b200 has no CPU fallback.".format(LIB_PATH))
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.lls_abi_version() != ABI_VERSION:
        raise LlsError("liblls.so ABI version {0} != expected {1}; rebuild".format(lib.lls_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().lls_last_error()
        raise LlsError("liblls call failed (code {0}): {1}".format(rc, msg.decode() if msg else "?"))
