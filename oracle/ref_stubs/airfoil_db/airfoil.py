"""Linear section model of airfoil_db.Airfoil, restated from its documented behaviour.

Sources for the formulas (file:line in /root/reference):
  * machupX/obsolete/airfoil.py:141-279  (CL, CD, Cm, aL0 of the linear model)
  * machupX/wing_segment.py:664-676      (Phillips flap effectiveness, hinge efficiency, Cm_delta)
  * docs/source/creating_input_files.md:188-218 (input keys and defaults)
The deflection-efficiency roll-off above 11 deg is NOT pinned by any reference test.

The "database" type (tables of CL, CD, Cm over alpha, Rey, ...) restates airfoil_db's look-up as
scipy.interpolate.griddata(method="linear") on the variables scaled to [0, 1], central differences for CLa / CLRe / CLM
(steps 0.001 / 1000 / 0.05) and the secant method for aL0.  That restatement IS pinned: oracle/recover_database_pin.py shows it
reproduces the golden of the reference's database test to the 16th digit once Qhull's arbitrary diagonals in exactly rectangular
cells are those of the golden's generation run.  Used by oracle/fuzz_host_mirror.py with scattered (tie-free) tables.
"""
import json
import numpy as np

_DEFL_KNEE = 0.19198621771937624  # 11 deg in rad
_DB_DEFAULTS = {"alpha": 0.0, "Rey": 1e6, "Mach": 0.0, "trailing_flap_deflection": 0.0, "trailing_flap_fraction": 0.0}
_DB_STEP = {"alpha": 0.001, "Rey": 1000.0, "Mach": 0.05}


class _Database:
    def __init__(self, name, filename):
        from .exceptions import DatabaseBoundsError
        self._error = DatabaseBoundsError
        self.name = name
        with open(filename) as fh:
            header = fh.readline().strip().lstrip("#").split()
        data = np.loadtxt(filename, skiprows=1)
        self.dofs = [d for d in _DB_DEFAULTS if d in header]
        self.X = np.stack([data[:, header.index(d)] for d in self.dofs], axis=1)
        self.vals = {k: data[:, header.index(k)] for k in ("CL", "CD", "Cm")}
        self.lo, self.hi = self.X.min(axis=0), self.X.max(axis=0)
        self.Xn = (self.X - self.lo) / (self.hi - self.lo)

    def look(self, coef, kw, shift=None):
        import scipy.interpolate as interp
        xs = [np.asarray(kw.get(d, _DB_DEFAULTS[d]), dtype=float) for d in self.dofs]
        shape = np.broadcast(*xs).shape
        P = np.stack([np.broadcast_to(x, shape).ravel() for x in xs], axis=1)
        if shift is not None:
            P[:, self.dofs.index(shift[0])] += shift[1]
        if len(self.dofs) == 1:
            order = np.argsort(self.X[:, 0])
            r = np.interp(P[:, 0], self.X[order, 0], self.vals[coef][order], left=np.nan, right=np.nan)
        else:
            r = interp.griddata(self.Xn, self.vals[coef], (P - self.lo) / (self.hi - self.lo), method="linear").flatten()
        if np.isnan(r).any():
            raise self._error(self.name, np.argwhere(np.isnan(r)).flatten(), kw)
        r = r.reshape(shape)
        return r.item() if r.ndim == 0 else r

    def slope(self, dof, kw):
        if dof not in self.dofs:
            shape = np.broadcast(*[np.asarray(kw.get(d, _DB_DEFAULTS[d]), dtype=float) for d in self.dofs]).shape
            return np.zeros(shape) if shape else 0.0
        dx = _DB_STEP[dof]
        return (self.look("CL", kw, (dof, dx)) - self.look("CL", kw, (dof, -dx))) / (2.0 * dx)

    def aL0(self, kw):
        kw = {k: v for k, v in kw.items() if k != "alpha"}
        shape = np.broadcast(*([np.asarray(kw.get(d, _DB_DEFAULTS[d]), dtype=float) for d in self.dofs if d != "alpha"] or [0.0])).shape
        a0, a1 = np.zeros(shape), np.full(shape, 0.01)
        f0 = np.asarray(self.look("CL", dict(kw, alpha=a0)))
        f1 = np.asarray(self.look("CL", dict(kw, alpha=a1)))
        for _ in range(50):
            act = np.abs(a1 - a0) > 1e-10
            if not act.any():
                break
            with np.errstate(divide="ignore", invalid="ignore"):
                a2 = np.where(act, a1 - f1 * (a0 - a1) / (f0 - f1), a1)
            a0, f0, a1, f1 = a1, f1, a2, np.asarray(self.look("CL", dict(kw, alpha=a2)))
        return a1.item() if a1.ndim == 0 else a1


def _flap_terms(delta, c_f):
    delta = np.asarray(delta, dtype=float)
    c_f = np.asarray(c_f, dtype=float)
    theta_f = np.arccos(2.0 * c_f - 1.0)
    eps_ideal = 1.0 - (theta_f - np.sin(theta_f)) / np.pi
    eta_hinge = 3.9598 * np.arctan((c_f + 0.006527) * 89.2574 + 4.898015) - 5.18786
    ad = np.abs(delta)
    eta_defl = np.where(ad <= _DEFL_KNEE, 1.0, -0.4995016675499485 * ad + 1.09589743589744)
    eps = eps_ideal * eta_hinge * eta_defl
    dCm = (np.sin(2.0 * theta_f) - 2.0 * np.sin(theta_f)) / 4.0
    return eps, dCm


class Airfoil:
    def __init__(self, name, airfoil_input, **kwargs):
        self.name = name
        if isinstance(airfoil_input, str):
            with open(airfoil_input) as fh:
                airfoil_input = json.load(fh)
        self._input = airfoil_input
        self._type = airfoil_input.get("type", "linear")
        if self._type == "database":
            db = self._db = _Database(name, airfoil_input["input_file"])
            self.get_CL = lambda **kw: db.look("CL", kw)
            self.get_CD = lambda **kw: db.look("CD", kw)
            self.get_Cm = lambda **kw: db.look("Cm", kw)
            self.get_CLa = lambda **kw: db.slope("alpha", kw)
            self.get_CLRe = lambda **kw: db.slope("Rey", kw)
            self.get_CLM = lambda **kw: db.slope("Mach", kw)
            self.get_aL0 = lambda **kw: db.aL0(kw)
            self._max_camber = self._max_thickness = 0.0
            return
        if self._type != "linear":
            raise NotImplementedError("oracle airfoil_db stub supports only 'linear' and 'database' airfoils, got %r" % self._type)
        g = airfoil_input.get
        self._aL0 = float(g("aL0", 0.0))
        self._CLa = float(g("CLa", 2.0 * np.pi))
        self._CmL0 = float(g("CmL0", 0.0))
        self._Cma = float(g("Cma", 0.0))
        self._CD0 = float(g("CD0", 0.0))
        self._CD1 = float(g("CD1", 0.0))
        self._CD2 = float(g("CD2", 0.0))
        self._CL_max = float(g("CL_max", np.inf))
        naca = airfoil_input.get("geometry", {}).get("NACA", None)
        if naca is not None and len(naca) == 4:
            self._max_camber = int(naca[0]) / 100.0
            self._max_thickness = int(naca[2:]) / 100.0
        else:
            self._max_camber = 0.0
            self._max_thickness = 0.0

    # geometry (stored by the reference, unused by the solve)
    def get_max_camber(self):
        return self._max_camber

    def get_max_thickness(self):
        return self._max_thickness

    def set_err_state(self, **kwargs):
        pass

    @staticmethod
    def _args(kw):
        alpha = np.asarray(kw.get("alpha", 0.0), dtype=float)
        delta = np.asarray(kw.get("trailing_flap_deflection", 0.0), dtype=float)
        c_f = np.asarray(kw.get("trailing_flap_fraction", 0.0), dtype=float)
        return alpha, delta, c_f

    def get_CL(self, **kw):
        alpha, delta, c_f = self._args(kw)
        eps, _ = _flap_terms(delta, c_f)
        CL = self._CLa * (alpha - self._aL0 + delta * eps)
        return np.clip(CL, -self._CL_max, self._CL_max)

    def get_CLa(self, **kw):
        alpha, delta, c_f = self._args(kw)
        return self._CLa * np.ones(np.broadcast(alpha, delta, c_f).shape) if np.broadcast(alpha, delta, c_f).shape else self._CLa

    def get_aL0(self, **kw):
        alpha, delta, c_f = self._args(kw)
        eps, _ = _flap_terms(delta, c_f)
        return self._aL0 - delta * eps + 0.0 * alpha

    def get_CLRe(self, **kw):
        alpha, delta, c_f = self._args(kw)
        return np.zeros(np.broadcast(alpha, delta, c_f).shape)

    def get_CLM(self, **kw):
        alpha, delta, c_f = self._args(kw)
        return np.zeros(np.broadcast(alpha, delta, c_f).shape)

    def get_CD(self, **kw):
        alpha, delta, c_f = self._args(kw)
        CL0 = np.clip(self._CLa * (alpha - self._aL0), -self._CL_max, self._CL_max)
        return self._CD0 + self._CD1 * CL0 + self._CD2 * CL0 * CL0 + 0.002 * np.abs(delta) * 180.0 / np.pi + 0.0 * c_f

    def get_Cm(self, **kw):
        alpha, delta, c_f = self._args(kw)
        _, dCm = _flap_terms(delta, c_f)
        return self._Cma * (alpha - self._aL0) + self._CmL0 + delta * dCm
