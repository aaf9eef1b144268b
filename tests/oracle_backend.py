"""TEST INFRASTRUCTURE: stand-ins for machupx_b200.device.DeviceScene / BatchDeviceScene that answer with the CPU oracle
(oracle/lls_oracle.py), so the HOST logic of machupx_b200.Scene -- table revisions, per-case tables of the batched
finite-difference stencils, driver arithmetic, error policy -- can be exercised in the CPU test-suite against the reference's
golden answers.  Nothing outside tests/ may import this module; the product has no CPU path (tests/test_abi.py checks)."""
import numpy as np

from machupx_b200 import layout as L
from oracle import lls_oracle as O


def _solve(tables, ac_state, solver_type, convergence, relaxation, max_iterations):
    T = O.Tables(tables, ac_state)
    res = O.solve(T, solver_type=solver_type, convergence=convergence, relaxation=relaxation, max_iterations=max_iterations)
    status = 0 if res["converged"] else L.STATUS_NOT_CONVERGED
    if res.get("section_bounds"):
        status |= L.STATUS_SECTION_BOUNDS
    return res, status


class OracleDeviceScene:
    def __init__(self, tables, device=None):
        self.tables = {k: (np.array(v, copy=True) if isinstance(v, np.ndarray) else v) for k, v in tables.items()}
        self.n, self.ld = int(tables["n"]), int(tables["ld"])
        self.n_aircraft, self.n_segments, self.flags = int(tables["n_aircraft"]), int(tables["n_segments"]), int(tables["flags"])
        self.n_airfoils = len(tables["airfoils"])
        self.bytes_h2d = self.bytes_d2h = 0
        self.last_status = 0
        self.uploads = ["tables"]
        self._solver = ("linear",)

    def upload_tables(self, tables):
        self.tables.update(tab=np.array(tables["tab"], copy=True), itab=np.array(tables["itab"], copy=True),
                           airfoils=np.array(tables["airfoils"], copy=True))
        self.uploads.append("tables")

    def upload_rows(self, tab, fields):
        for f in fields:
            self.tables["tab"][f] = tab[f]
        self.uploads.append(("rows", tuple(fields)))

    def upload_state(self, ac_state):
        self.ac = np.array(L.pad_ac_state(ac_state), copy=True)

    def flow_state(self):
        self._solver = ("none",)
        self._linear_start = False

    def build_influence(self, impingement_threshold=1e-10):
        pass

    def solve_linear(self):
        self._solver = ("linear",)
        self._linear_start = True

    def solve_nonlinear(self, convergence, relaxation, max_iterations):
        self._solver = ("nonlinear", convergence, relaxation, max_iterations)
        if self._linear_start or getattr(self, "_res", None) is None:
            self._res, status = _solve(self.tables, self.ac, "nonlinear", convergence, relaxation, max_iterations)
        else:
            # initial_guess="previous" (scene.py:1462-1478): Newton starts from the circulation the last solve left in the result
            # table -- after its integration, i.e. rescaled if match_machup_pro is set (scene.py:938-939)
            T = O.Tables(self.tables, self.ac)
            F = O.flow_state(T)
            V = O.influence(T, F)
            gamma, it, err, ok, hist = O.solve_nonlinear(T, F, V, np.array(self._res["dist"]["gamma"], copy=True), convergence, relaxation,
                                                         max_iterations)
            seg, D = O.integrate(T, F, V, gamma)
            self._res = {"flow": F, "V": V, "gamma": gamma, "iterations": it, "error": err, "converged": ok, "history": hist,
                         "seg_fm": seg, "dist": D, "section_bounds": T.section_bounds}
            status = (0 if ok else L.STATUS_NOT_CONVERGED) | (L.STATUS_SECTION_BOUNDS if T.section_bounds else 0)
        return self._res["iterations"], self._res["error"], status, list(self._res["history"])

    def integrate(self):
        if self._solver[0] != "nonlinear":
            self._res, self.last_status = _solve(self.tables, self.ac, "linear", 1e-10, 1.0, 1)
        else:
            self.last_status = L.STATUS_SECTION_BOUNDS if self._res.get("section_bounds") else 0
        return np.array(self._res["seg_fm"], copy=True)

    def status(self):
        if self._solver[0] != "nonlinear":       # the flow state and the linear solve have evaluated CL at the freestream angle
            T = O.Tables(self.tables, self.ac)
            O.flow_state(T)
            return L.STATUS_SECTION_BOUNDS if T.section_bounds else 0
        return 0

    def out(self):
        """The per-control-point result table the kernels leave behind (enum lls_out_field), from the oracle's intermediate
        results: what flow_state_kernel, the last residual and integrate_kernel write."""
        res, D, F = self._res, self._res["dist"], self._res["flow"]
        out = np.zeros((L.NO, self.n))
        out[L.O_GAMMA] = D["gamma"]             # with match_machup_pro the integration rescales it in place (scene.py:938-939)
        out[L.O_VI:L.O_VI + 3] = np.asarray(D["v_i"]).T
        out[L.O_ALPHA], out[L.O_CD], out[L.O_CM], out[L.O_RE], out[L.O_MACH] = D["alpha"], D["CD"], D["Cm"], D["Re"], D["M"]
        for f, key in ((L.O_DFINV, "dF_inv"), (L.O_DMINV, "dM_inv"), (L.O_DFVISC, "dF_visc"), (L.O_DMVISC, "dM_visc")):
            out[f:f + 3] = np.asarray(D[key]).T
        out[L.O_AL0] = F.aL0
        out[L.O_REDIM_FULL] = D["redim_full"]
        if D["redim_in_plane"] is not None:
            out[L.O_REDIM_IP] = D["redim_in_plane"]
        out[L.O_UTRAIL0:L.O_UTRAIL0 + 3], out[L.O_UTRAIL1:L.O_UTRAIL1 + 3] = F.u_trailing_0.T, F.u_trailing_1.T
        out[L.O_VI2] = D["V_i_2"]
        if self._solver[0] == "nonlinear":       # section lift of the last residual evaluation (scene.py:911-913)
            T = O.Tables(self.tables, self.ac)
            out[L.O_CL] = O.residual(T, F, res["V"], res["gamma"]).CL
        else:                                    # a linear solve leaves the flow-state estimate (scene.py:659-666)
            out[L.O_CL] = F.CL
        return out


class OracleBatchDeviceScene:
    def __init__(self, tables, ncases, device=None):
        self.base = OracleDeviceScene(tables)
        self.ncases = int(ncases)
        self.n, self.ld, self.flags = self.base.n, self.base.ld, self.base.flags
        self.n_aircraft, self.n_segments = self.base.n_aircraft, self.base.n_segments
        self.case_tables = None
        self.log = []

    def upload_tables(self, tables):
        self.base.upload_tables(tables)

    def upload_rows(self, tab, fields):
        self.base.upload_rows(tab, fields)

    def upload_case_tables(self, case_tables):
        assert np.shape(case_tables) == (self.ncases, L.NF, self.ld)
        self.case_tables = np.array(case_tables, copy=True)
        self.log.append("case_tables")

    def use_shared_tables(self):
        self.case_tables = None
        self.log.append("shared")

    def upload_states(self, ac_states):
        self.ac = np.array(L.pad_ac_state(np.asarray(ac_states).reshape(self.ncases, self.n_aircraft, -1)), copy=True)

    def solve(self, solver_type, convergence, relaxation, max_iterations, impingement_threshold=1e-10, linear_start=True):
        nc = self.ncases
        seg = np.zeros((nc, self.n_segments, 12))
        its, err, status = np.zeros(nc, dtype=np.int32), np.zeros(nc), np.zeros(nc, dtype=np.int32)
        for c in range(nc):
            tables = dict(self.base.tables)
            if self.case_tables is not None:
                tables["tab"] = self.case_tables[c]
            res, status[c] = _solve(tables, self.ac[c], solver_type, convergence, relaxation, max_iterations)
            seg[c], its[c], err[c] = res["seg_fm"], res["iterations"], res["error"]
        return seg, its, err, status


def install(monkeypatch):
    """Routes machupx_b200.Scene to the oracle stand-ins for the duration of one test."""
    import machupx_b200.device as device
    monkeypatch.setattr(device, "DeviceScene", OracleDeviceScene)
    monkeypatch.setattr(device, "BatchDeviceScene", OracleBatchDeviceScene)
