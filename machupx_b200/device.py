"""Device-resident lifting-line scene: PyTorch tensors own the HBM buffers, liblls does the work.

``DeviceScene`` is the thin layer between the Python ``Scene`` (API mirror of the reference) and
the C ABI in include/lls.h: it uploads the O(N) tables, keeps V_ji / J / gamma resident on the
GPU between solves (the reference's ``initial_guess="previous"`` warm start, scene.py:1437-1438)
and returns results as numpy views.  There is no CPU code path here.
"""
import ctypes

import numpy as np
import torch

from . import _native
from . import layout as L


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


class DeviceScene:
    def __init__(self, tables, device=None):
        if not torch.cuda.is_available():
            raise _native.LlsError("machupx_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = _native.load()
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.n = int(tables["n"])
        self.ld = int(tables["ld"])
        self.n_aircraft = int(tables["n_aircraft"])
        self.n_segments = int(tables["n_segments"])
        self.flags = int(tables["flags"])
        af = np.ascontiguousarray(tables["airfoils"], dtype=np.float64).reshape(-1, L.AF_STRIDE)
        self.n_airfoils = af.shape[0]
        sizes = _native.Sizes()
        _native.check(self.lib.lls_query_sizes(self.n, self.n_segments, ctypes.byref(sizes)))
        assert sizes.ld == self.ld, "table ld %d != library ld %d" % (self.ld, sizes.ld)
        self.sizes = sizes
        with torch.cuda.device(self.device):
            self.handle = ctypes.c_void_p()
            _native.check(self.lib.lls_create(ctypes.byref(self.handle), self.n, self.n_aircraft, self.n_segments,
                                              self.n_airfoils, self.flags))
            dev = self.device
            self.d_tab = torch.empty((L.NF, self.ld), dtype=torch.float64, device=dev)
            self.d_itab = torch.empty((L.NI, self.ld), dtype=torch.int32, device=dev)
            self.d_af = torch.empty((self.n_airfoils, L.AF_STRIDE), dtype=torch.float64, device=dev)
            self.d_out = torch.zeros((L.NO, self.ld), dtype=torch.float64, device=dev)
            self.d_V = torch.empty((self.n, 3, self.ld), dtype=torch.float64, device=dev)
            self.d_M = torch.empty((self.ld, self.ld), dtype=torch.float64, device=dev)
            self.d_work = torch.zeros(sizes.work_bytes, dtype=torch.uint8, device=dev)
            self.d_ac = torch.zeros((self.n_aircraft, L.AC_STRIDE), dtype=torch.float64, device=dev)
            self.d_seg_fm = torch.zeros((self.n_segments, 12), dtype=torch.float64, device=dev)
            self.h_ac = torch.zeros((self.n_aircraft, L.AC_STRIDE), dtype=torch.float64).pin_memory()
            self.h_seg_fm = torch.zeros((self.n_segments, 12), dtype=torch.float64).pin_memory()
            self.h_status = torch.zeros(1, dtype=torch.int32).pin_memory()
            _native.check(self.lib.lls_bind(self.handle, _ptr(self.d_tab), _ptr(self.d_itab), _ptr(self.d_af), _ptr(self.d_out),
                                            _ptr(self.d_V), _ptr(self.d_M), _ptr(self.d_work), ctypes.c_size_t(sizes.work_bytes)))
            _native.check(self.lib.lls_set_state(self.handle, _ptr(self.d_ac)))
            off = ctypes.c_size_t(0)
            _native.check(self.lib.lls_status_offset(self.handle, ctypes.byref(off)))
            self.d_status = self.d_work[off.value:off.value + 4].view(torch.int32)
        self.upload_tables(tables)
        self.bytes_h2d = 0
        self.bytes_d2h = 0
        self.last_status = 0

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.lls_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # -- uploads ---------------------------------------------------------------------------------
    def upload_tables(self, tables):
        """Geometry / control / atmosphere tables (changes with geometry, orientation, position or controls)."""
        tab = np.ascontiguousarray(tables["tab"], dtype=np.float64)
        itab = np.ascontiguousarray(tables["itab"], dtype=np.int32)
        af = np.ascontiguousarray(tables["airfoils"], dtype=np.float64).reshape(-1, L.AF_STRIDE)
        assert tab.shape == (L.NF, self.ld) and itab.shape == (L.NI, self.ld) and af.shape[0] == self.n_airfoils
        self.d_tab.copy_(torch.from_numpy(tab))
        self.d_itab.copy_(torch.from_numpy(itab))
        self.d_af.copy_(torch.from_numpy(af))

    def upload_rows(self, tab, fields):
        """Re-uploads single rows of the double table (control deflections change only the flap rows)."""
        for f in fields:
            self.d_tab[f].copy_(torch.from_numpy(np.ascontiguousarray(tab[f], dtype=np.float64)))

    def upload_state(self, ac_state):
        self.h_ac.copy_(torch.from_numpy(np.ascontiguousarray(L.pad_ac_state(ac_state)).reshape(self.n_aircraft, L.AC_STRIDE)))
        self.d_ac.copy_(self.h_ac, non_blocking=True)
        self.bytes_h2d += self.h_ac.numel() * 8

    def set_gamma(self, gamma):
        g = torch.zeros(self.ld, dtype=torch.float64)
        g[:self.n] = torch.from_numpy(np.asarray(gamma, dtype=np.float64))
        self.d_out[L.O_GAMMA].copy_(g)

    # -- hot path ----------------------------------------------------------------------------------
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def flow_state(self):
        with torch.cuda.device(self.device):
            _native.check(self.lib.lls_flow_state(self.handle, self._stream()))

    def build_influence(self, impingement_threshold=1e-10):
        with torch.cuda.device(self.device):
            _native.check(self.lib.lls_build_influence(self.handle, float(impingement_threshold), self._stream()))

    def solve_linear(self):
        with torch.cuda.device(self.device):
            _native.check(self.lib.lls_solve_linear(self.handle, self._stream()))

    def solve_nonlinear(self, convergence, relaxation, max_iterations):
        it = ctypes.c_int(0)
        err = ctypes.c_double(0.0)
        status = ctypes.c_int(0)
        hist = (ctypes.c_double * max(1, int(max_iterations)))()
        with torch.cuda.device(self.device):
            _native.check(self.lib.lls_solve_nonlinear(self.handle, float(convergence), float(relaxation), int(max_iterations),
                                                       self._stream(), ctypes.byref(it), ctypes.byref(err), ctypes.byref(status), hist))
        return it.value, err.value, status.value, [hist[k] for k in range(it.value)]

    def integrate(self):
        with torch.cuda.device(self.device):
            _native.check(self.lib.lls_integrate(self.handle, _ptr(self.d_seg_fm), self._stream()))
            self.h_seg_fm.copy_(self.d_seg_fm, non_blocking=True)
            self.h_status.copy_(self.d_status, non_blocking=True)      # status word after the integration, same synchronisation
            torch.cuda.current_stream(self.device).synchronize()
        self.bytes_d2h += self.h_seg_fm.numel() * 8 + 4
        self.last_status = int(self.h_status[0])
        return self.h_seg_fm.numpy().copy()

    def status(self):
        st = ctypes.c_int(0)
        with torch.cuda.device(self.device):
            _native.check(self.lib.lls_get_status(self.handle, self._stream(), ctypes.byref(st)))
        return st.value

    def set_timing(self, enabled):
        _native.check(self.lib.lls_set_timing(self.handle, 1 if enabled else 0))

    def timings(self):
        ms = (ctypes.c_double * 6)()
        _native.check(self.lib.lls_get_timings(self.handle, ms))
        return dict(zip(("influence", "assemble", "lu", "trisolve", "residual", "integrate"), list(ms)))

    # -- results -----------------------------------------------------------------------------------
    def out(self):
        """(NO, n) numpy copy of the per-control-point solution table."""
        return self.d_out[:, :self.n].cpu().numpy()

    def gamma(self):
        return self.d_out[L.O_GAMMA, :self.n].cpu().numpy()

    def V_ji(self):
        """(n, n, 3) numpy copy in the reference's index order (control point, vortex, component)."""
        return self.d_V[:, :, :self.n].permute(0, 2, 1).contiguous().cpu().numpy()

    # -- standalone LU (tests, bench) ------------------------------------------------------------
    def lu_solve_dense(self, A, b):
        """Solves A x = b with the device LU (A: (n, n) numpy).  Used by the LU parity tests."""
        n, ld = self.n, self.ld
        Ap = np.eye(ld)
        Ap[:n, :n] = A
        bp = np.zeros(ld)
        bp[:n] = b
        dA = torch.from_numpy(Ap).to(self.device)
        db = torch.from_numpy(bp).to(self.device)
        with torch.cuda.device(self.device):
            _native.check(self.lib.lls_lu_factor(self.handle, _ptr(dA), self._stream()))
            _native.check(self.lib.lls_lu_solve(self.handle, _ptr(dA), _ptr(db), self._stream()))
            torch.cuda.synchronize(self.device)
        return db[:n].cpu().numpy(), dA.cpu().numpy()


class BatchDeviceScene:
    """``ncases`` flight states of one geometry solved by the same launches (lls_bind_batch): the device side of alpha /
    beta / rate sweeps and of the finite-difference stencils of the derivative and trim drivers.  Geometry tables are
    uploaded once; per case only the 12-double aircraft state block differs."""

    def __init__(self, tables, ncases, device=None):
        if not torch.cuda.is_available():
            raise _native.LlsError("machupx_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = _native.load()
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.n, self.ld, self.ncases = int(tables["n"]), int(tables["ld"]), int(ncases)
        self.n_aircraft, self.n_segments, self.flags = int(tables["n_aircraft"]), int(tables["n_segments"]), int(tables["flags"])
        af = np.ascontiguousarray(tables["airfoils"], dtype=np.float64).reshape(-1, L.AF_STRIDE)
        self.n_airfoils = af.shape[0]
        sizes = _native.Sizes()
        _native.check(self.lib.lls_query_sizes(self.n, self.n_segments, ctypes.byref(sizes)))
        self.work_stride = (int(sizes.work_bytes) + 255) // 256 * 256
        nc, dev = self.ncases, self.device
        with torch.cuda.device(dev):
            self.handle = ctypes.c_void_p()
            _native.check(self.lib.lls_create(ctypes.byref(self.handle), self.n, self.n_aircraft, self.n_segments, self.n_airfoils, self.flags))
            self.d_tab = torch.from_numpy(np.ascontiguousarray(tables["tab"], dtype=np.float64)).to(dev)
            self.d_itab = torch.from_numpy(np.ascontiguousarray(tables["itab"], dtype=np.int32)).to(dev)
            self.d_af = torch.from_numpy(af).to(dev)
            self.d_out = torch.zeros((nc, L.NO, self.ld), dtype=torch.float64, device=dev)
            self.d_V = torch.empty((nc, self.n, 3, self.ld), dtype=torch.float64, device=dev)
            self.d_M = torch.empty((nc, self.ld, self.ld), dtype=torch.float64, device=dev)
            self.d_work = torch.zeros((nc, self.work_stride), dtype=torch.uint8, device=dev)
            self.d_ctl = torch.zeros(2 * nc + 4, dtype=torch.int32, device=dev)
            self.d_ac = torch.zeros((nc, self.n_aircraft, L.AC_STRIDE), dtype=torch.float64, device=dev)
            self.d_seg_fm = torch.zeros((nc, self.n_segments, 12), dtype=torch.float64, device=dev)
            self.h_ac = torch.zeros((nc, self.n_aircraft, L.AC_STRIDE), dtype=torch.float64).pin_memory()
            self.h_seg_fm = torch.zeros((nc, self.n_segments, 12), dtype=torch.float64).pin_memory()
            _native.check(self.lib.lls_bind_batch(self.handle, nc, _ptr(self.d_tab), _ptr(self.d_itab), _ptr(self.d_af), _ptr(self.d_out),
                                                  _ptr(self.d_V), _ptr(self.d_M), _ptr(self.d_work), ctypes.c_size_t(self.work_stride),
                                                  _ptr(self.d_ctl)))
            _native.check(self.lib.lls_set_state(self.handle, _ptr(self.d_ac)))

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.lls_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def upload_tables(self, tables):
        """Shared geometry / control / atmosphere tables of every case of the batch."""
        af = np.ascontiguousarray(tables["airfoils"], dtype=np.float64).reshape(-1, L.AF_STRIDE)
        assert af.shape[0] == self.n_airfoils
        self.d_tab.copy_(torch.from_numpy(np.ascontiguousarray(tables["tab"], dtype=np.float64)))
        self.d_itab.copy_(torch.from_numpy(np.ascontiguousarray(tables["itab"], dtype=np.int32)))
        self.d_af.copy_(torch.from_numpy(af))

    def upload_rows(self, tab, fields):
        for f in fields:
            self.d_tab[f].copy_(torch.from_numpy(np.ascontiguousarray(tab[f], dtype=np.float64)))

    def upload_case_tables(self, case_tables):
        """(ncases, NF, ld) per-case copies of the double table (cases that differ in control deflections or pose)."""
        t = np.ascontiguousarray(case_tables, dtype=np.float64)
        assert t.shape == (self.ncases, L.NF, self.ld), t.shape
        if getattr(self, "d_tab_cases", None) is None:
            self.d_tab_cases = torch.empty((self.ncases, L.NF, self.ld), dtype=torch.float64, device=self.device)
        self.d_tab_cases.copy_(torch.from_numpy(t))
        with torch.cuda.device(self.device):
            _native.check(self.lib.lls_set_case_tables(self.handle, _ptr(self.d_tab_cases), ctypes.c_size_t(L.NF * self.ld)))
        self._case_tables_on = True

    def use_shared_tables(self):
        if getattr(self, "_case_tables_on", False):
            with torch.cuda.device(self.device):
                _native.check(self.lib.lls_set_case_tables(self.handle, _ptr(self.d_tab), ctypes.c_size_t(0)))
            self._case_tables_on = False

    def upload_states(self, ac_states):
        self.h_ac.copy_(torch.from_numpy(np.ascontiguousarray(L.pad_ac_state(np.asarray(ac_states, dtype=np.float64).reshape(self.ncases, self.n_aircraft, -1)))))
        self.d_ac.copy_(self.h_ac, non_blocking=True)

    def solve(self, solver_type, convergence, relaxation, max_iterations, impingement_threshold=1e-10, linear_start=True):
        """flow state + influence build + linear solve (+ Newton) + integration for every case.
        Returns (seg_fm (ncases, n_segments, 12) numpy, iterations, final_error, status)."""
        nc = self.ncases
        its = (ctypes.c_int * nc)()
        err = (ctypes.c_double * nc)()
        status = (ctypes.c_int * nc)()
        with torch.cuda.device(self.device):
            st = self._stream()
            _native.check(self.lib.lls_flow_state(self.handle, st))
            _native.check(self.lib.lls_build_influence(self.handle, float(impingement_threshold), st))
            if solver_type == "linear" or linear_start:
                _native.check(self.lib.lls_solve_linear(self.handle, st))
            if solver_type == "nonlinear":
                _native.check(self.lib.lls_solve_nonlinear_batch(self.handle, float(convergence), float(relaxation), int(max_iterations), st,
                                                                 its, err, status))
            else:
                _native.check(self.lib.lls_get_status_batch(self.handle, st, status))
            _native.check(self.lib.lls_integrate(self.handle, _ptr(self.d_seg_fm), st))
            self.h_seg_fm.copy_(self.d_seg_fm, non_blocking=True)
            late = (ctypes.c_int * nc)()       # bits raised by the integration (section bounds of CD / Cm); synchronises the stream
            _native.check(self.lib.lls_get_status_batch(self.handle, st, late))
        return (self.h_seg_fm.numpy().copy(), np.frombuffer(its, dtype=np.int32).copy(), np.frombuffer(err, dtype=np.float64).copy(),
                np.frombuffer(status, dtype=np.int32) | np.frombuffer(late, dtype=np.int32))

    def gamma(self):
        return self.d_out[:, L.O_GAMMA, :self.n].cpu().numpy()
