"""Section aerodynamics (the part of the hot path the reference delegates to ``airfoil_db``).

The reference evaluates CL, CLa, aL0, CLRe, CLM, CD, Cm per control point through the
third-party ``airfoil_db.Airfoil`` (>=1.4.3, reference setup.py:13; call sites
wing_segment.py:1081-1085, 1111-1120).  That package is not part of the reference tree, so
this module restates its *linear* airfoil model.  The formulas are anchored on the
reference's own material:

* machupX/obsolete/airfoil.py:141-279 — CL = CLa (alpha - aL0 + eps*delta) clipped at
  +-CL_max; CD = CD0 + CD1 CL + CD2 CL^2 (CL without flap) + 0.002 |delta| 180/pi.
* machupX/wing_segment.py:664-676 — Phillips' ideal flap effectiveness, hinge efficiency
  curve fit (Mechanics of Flight fig. 1.7.4) and flap moment increment.
* the golden values in test/wing_segment_tests/test_airfoil_interpolation.py and
  test/scene_tests/test_solve_forces.py, which select Cm = Cma (alpha - aL0) + CmL0 + ...

The same arithmetic runs on the device inside the Newton loop (csrc/section.cuh); this host
copy serves ``WingSegment.get_cp_*`` and exports the parameter block the kernels read.

NOT pinned by any reference test (see DESIGN.md "parity unpinned"): the deflection
efficiency roll-off beyond 11 degrees and the CL_max clip interplay with CLa.

``"database"`` airfoils (tables of CL, CD, Cm over any of alpha, Rey, Mach, flap deflection, flap fraction): airfoil_db
interpolates them with ``scipy.interpolate.griddata(..., method="linear")`` on the variables scaled to [0, 1], i.e.
barycentric interpolation on the Delaunay triangulation of the data points; CLa / CLRe / CLM are central differences of
that interpolant (steps 0.001 rad / 1000 / 0.05) and aL0 its root in alpha.  PINNED: with this model the reference's one
"database" test (test/scene_tests/test_solve_forces.py:453-494) is reproduced to the 16th digit once the triangulation's
arbitrary diagonals in the table's exactly rectangular cells are the ones the reference's golden was generated with
(tests/test_database_airfoils.py; Qhull breaks those ties differently from build to build, and each flipped diagonal moves
the forces by 1e-4 relative).  The triangulation is built HERE, on the host, with the user's own scipy -- the same call
airfoil_db makes -- and travels to the device as a table of simplices; evaluation is on the device.
"""
import json

import numpy as np

from . import layout as L

_DEFLECTION_KNEE = 0.19198621771937624  # 11 deg


def flap_effectiveness(delta, c_f):
    """Returns (eps, dCm_ddelta) for trailing-edge flap deflection delta [rad], chord fraction c_f."""
    delta = np.asarray(delta, dtype=float)
    c_f = np.asarray(c_f, dtype=float)
    theta_f = np.arccos(2.0 * c_f - 1.0)
    eps_ideal = 1.0 - (theta_f - np.sin(theta_f)) / np.pi
    eta_hinge = 3.9598 * np.arctan((c_f + 0.006527) * 89.2574 + 4.898015) - 5.18786
    mag = np.abs(delta)
    eta_defl = np.where(mag <= _DEFLECTION_KNEE, 1.0, -0.4995016675499485 * mag + 1.09589743589744)
    dcm = (np.sin(2.0 * theta_f) - 2.0 * np.sin(theta_f)) / 4.0
    return eps_ideal * eta_hinge * eta_defl, dcm


class DatabaseBoundsError(Exception):
    """machupX/__init__.py:11.  Raised by the host getters of a "database" airfoil asked for a point outside its table; the
    device reports the same condition through LLS_STATUS_SECTION_BOUNDS (Scene error policy "database_bounds")."""

    _TEXT = "The inputs to the airfoil database fell outside the bounds of available data."

    def __init__(self, airfoil="", exception_indices=(), inputs_dict=None):
        """airfoil_db's signature (airfoil name, indices of the offending points, the inputs).  Scene raises it with one
        argument, a ready-made text, which then is the message."""
        text_only = bool(airfoil) and inputs_dict is None and len(exception_indices) == 0
        self.airfoil = "" if text_only else airfoil
        self.exception_indices = exception_indices
        self.inputs_dict = inputs_dict or {}
        super().__init__(airfoil if text_only else self._TEXT)


def triangulate(points):
    """Simplices (S, V+1) of the Delaunay triangulation of `points` (P, V >= 2): the call scipy's griddata makes
    (LinearNDInterpolator -> scipy.spatial.Delaunay with its default Qhull options).  A module-level seam so that a test can
    pin the arbitrary diagonals of exactly co-circular cells."""
    from scipy.spatial import Delaunay
    return np.asarray(Delaunay(points).simplices, dtype=np.int64)


DB_EPS = 1e-12          # barycentric tolerance of the point location (normalised variables are O(1))
DB_STEP = {"alpha": 0.001, "Rey": 1000.0, "Mach": 0.05}      # central-difference steps of CLa / CLRe / CLM


class Airfoil:
    """Airfoil section with the subset of the airfoil_db API MachUpX uses."""

    def __init__(self, name, airfoil_input=None, **kwargs):
        self.name = name
        if airfoil_input is None:
            airfoil_input = {}
        if isinstance(airfoil_input, str):
            with open(airfoil_input) as handle:
                airfoil_input = json.load(handle)
        self._input_dict = airfoil_input
        self._type = airfoil_input.get("type", "linear")
        if self._type == "poly_fit":
            self._init_poly_fit(airfoil_input)
            return
        if self._type == "database":
            self._init_database(airfoil_input)
            return
        if self._type != "linear":
            raise NotImplementedError(
                "Airfoil '{0}': section type '{1}' cannot be evaluated on the device; only 'linear', 'poly_fit' and 'database' "
                "airfoils are supported (callables would need a CPU path).".format(name, self._type))
        get = airfoil_input.get
        self.aL0 = float(get("aL0", 0.0))
        self.CLa = float(get("CLa", 2.0 * np.pi))
        self.CmL0 = float(get("CmL0", 0.0))
        self.Cma = float(get("Cma", 0.0))
        self.CD0 = float(get("CD0", 0.0))
        self.CD1 = float(get("CD1", 0.0))
        self.CD2 = float(get("CD2", 0.0))
        self.CL_max = float(get("CL_max", np.inf))
        naca = airfoil_input.get("geometry", {}).get("NACA", None)
        if isinstance(naca, str) and len(naca) == 4:
            self._max_camber = int(naca[0]) / 100.0
            self._max_thickness = int(naca[2:]) / 100.0
        else:
            self._max_camber = 0.0
            self._max_thickness = 0.0

    # -- poly_fit airfoils ---------------------------------------------------------------------
    # Multivariate polynomial fits of CL, CD, Cm in any of (alpha, Rey, Mach, trailing_flap_deflection,
    # trailing_flap_fraction), in the file format MachUpX's tooling writes (reference dev/NACA_0012_fits.json; exponent
    # tuple row-major with the last variable fastest, machupX/poly_fits.py:398-431).  airfoil_db is not in the reference tree,
    # so the semantics beyond the polynomial itself are DEFINED HERE (parity unpinned): CLa / CLRe / CLM are
    # the analytic partial derivatives of the CL fit, aL0 its root in alpha, variables the fit does not declare are ignored
    # (in particular flaps act only through declared flap variables), evaluation outside `limits` sets
    # LLS_STATUS_SECTION_BOUNDS on the device (DatabaseBoundsError policy on the host) and extrapolates the polynomial.
    _DEFAULTS = {"alpha": 0.0, "Rey": 1e6, "Mach": 0.0, "trailing_flap_deflection": 0.0, "trailing_flap_fraction": 0.0}

    def _init_poly_fit(self, airfoil_input):
        spec = airfoil_input
        if "input_file" in airfoil_input:
            with open(airfoil_input["input_file"]) as handle:
                spec = json.load(handle)
        self.dofs = list(spec["degrees_of_freedom"])
        for d in self.dofs:
            if d not in L.POLY_DOF_KINDS:
                raise IOError("Airfoil '{0}': unknown poly_fit variable '{1}'.".format(self.name, d))
        self.limits = np.asarray(spec["limits"], dtype=float).reshape(len(self.dofs), 2)
        self.fit_degrees = {k: [int(x) for x in spec["fit_degrees"][k]] for k in ("CL", "CD", "Cm")}
        self.fit_coefs = {k: np.asarray(spec["fit_coefs"][k], dtype=float) for k in ("CL", "CD", "Cm")}
        for k in ("CL", "CD", "Cm"):
            if len(self.fit_coefs[k]) != int(np.prod([d + 1 for d in self.fit_degrees[k]])):
                raise IOError("Airfoil '{0}': fit_coefs['{1}'] does not match fit_degrees.".format(self.name, k))
        self._max_camber = 0.0
        self._max_thickness = 0.0

    def _poly(self, which, kwargs, d_wrt=None):
        xs = [np.asarray(kwargs.get(d, self._DEFAULTS[d]), dtype=float) for d in self.dofs]
        shape = np.broadcast(*xs).shape if xs else ()
        xs = [np.broadcast_to(x, shape) for x in xs]
        coefs = self.fit_coefs[which].reshape([d + 1 for d in self.fit_degrees[which]])
        if d_wrt is not None:
            if d_wrt not in self.dofs:
                return np.zeros(shape)
            ax = self.dofs.index(d_wrt)
            n = np.arange(coefs.shape[ax]).reshape([-1 if a == ax else 1 for a in range(coefs.ndim)])
            coefs = np.delete(coefs * n, 0, axis=ax)          # d/dx sum a_n x^n = sum n a_n x^(n-1)
            if coefs.shape[ax] == 0:
                return np.zeros(shape)
        # straightforward evaluation: sum over the exponent grid
        total = np.zeros(shape)
        for idx in np.ndindex(*coefs.shape):
            term = coefs[idx]
            if term == 0.0:
                continue
            t = np.full(shape, term)
            for x, e in zip(xs, idx):
                if e:
                    t = t * x ** e
            total = total + t
        return total

    def _poly_aL0(self, kwargs):
        kw = dict(kwargs)
        a = np.zeros(np.broadcast(*[np.asarray(kw.get(d, self._DEFAULTS[d]), dtype=float) for d in self.dofs if d != "alpha"] or [0.0]).shape)
        for _ in range(30):
            kw["alpha"] = a
            step = self._poly("CL", kw) / self._poly("CL", kw, "alpha")
            a = a - step
            if not (np.abs(step) > 1e-15).any():
                break
        return a

    def out_of_bounds(self, **kwargs):
        bad = False
        for d, (lo, hi) in zip(self.dofs, self.limits):
            x = np.asarray(kwargs.get(d, self._DEFAULTS[d]), dtype=float)
            bad = bad | ~((x >= lo) & (x <= hi))
        return bad

    def lls_pool(self):
        """Pool record of a poly_fit / database airfoil (layout in include/lls.h)."""
        if self._type == "database":
            return self.lls_db_pool()
        rec = [self.limits.ravel()]
        for k in ("CL", "CD", "Cm"):
            rec.append(np.concatenate(([float(len(self.fit_coefs[k]))], np.asarray(self.fit_degrees[k], dtype=float), self.fit_coefs[k])))
        return np.concatenate(rec)

    # -- database airfoils -----------------------------------------------------------------------
    # File format (airfoil_db's export; reference test/NACA 2414.txt): one header line `# alpha Rey ... CL CD Cm` naming the
    # columns, then one data point per row.  See the module docstring for the model and how it is pinned.
    def _init_database(self, airfoil_input):
        path = airfoil_input.get("input_file")
        if path is None:
            raise IOError("Airfoil '{0}': a 'database' airfoil needs an 'input_file'.".format(self.name))
        with open(path) as handle:
            header = handle.readline().strip().lstrip("#").split()
        data = np.atleast_2d(np.loadtxt(path, skiprows=1))
        if data.shape[1] != len(header):
            raise IOError("Airfoil '{0}': header and data of {1} disagree.".format(self.name, path))
        self.dofs = [h for h in L.POLY_DOF_KINDS if h in header]
        for k in ("CL", "CD", "Cm"):
            if k not in header:
                raise IOError("Airfoil '{0}': column {1} missing in {2}.".format(self.name, k, path))
        if not self.dofs:
            raise IOError("Airfoil '{0}': {1} has no independent variable.".format(self.name, path))
        X = np.stack([data[:, header.index(d)] for d in self.dofs], axis=1)
        self._db_values = np.stack([data[:, header.index(k)] for k in ("CL", "CD", "Cm")], axis=1)
        self.limits = np.stack([X.min(axis=0), X.max(axis=0)], axis=1)
        if (self.limits[:, 1] <= self.limits[:, 0]).any():
            raise IOError("Airfoil '{0}': a variable of {1} never changes; leave its column out.".format(self.name, path))
        self._db_points = X
        self._db_normed = self._db_normalise(X)
        V = len(self.dofs)
        if V == 1:
            order = np.argsort(X[:, 0], kind="stable")
            simplices = np.stack([order[:-1], order[1:]], axis=1)
        else:
            simplices = triangulate(self._db_normed)
        self._set_simplices(simplices)
        self._max_camber = 0.0
        self._max_thickness = 0.0

    def _db_normalise(self, X):
        return (X - self.limits[:, 0]) / (self.limits[:, 1] - self.limits[:, 0])

    def _set_simplices(self, simplices):
        """Stores the simplices and, per simplex, the affine map to barycentric coordinates: c[:V] = Tinv (x - r), r the last
        vertex (the layout of scipy's Delaunay.transform).  Degenerate (flat) simplices get a NaN map and are never selected."""
        simplices = np.asarray(simplices, dtype=np.int64)
        V = len(self.dofs)
        P = self._db_normed[simplices]                                  # (S, V+1, V)
        r = P[:, V, :]
        T = np.transpose(P[:, :V, :] - r[:, None, :], (0, 2, 1))        # columns = edge vectors
        Tinv = np.full_like(T, np.nan)
        ok = np.abs(np.linalg.det(T)) > 1e-300
        Tinv[ok] = np.linalg.inv(T[ok])
        self._db_simplices, self._db_Tinv, self._db_r = simplices, Tinv, r

    def _db_lookup(self, which, kwargs, shift=None):
        """Values of coefficient `which` (0 CL, 1 CD, 2 Cm) at the points the keyword arguments describe; `shift` = (dof, dx)
        moves one variable.  Returns (values, found); points outside the table extrapolate from the closest simplex."""
        xs = [np.asarray(kwargs.get(d, self._DEFAULTS[d]), dtype=float) for d in self.dofs]
        shape = np.broadcast(*xs).shape
        X = np.stack([np.broadcast_to(x, shape).ravel() for x in xs], axis=1)
        if shift is not None:
            X[:, self.dofs.index(shift[0])] += shift[1]
        xn = self._db_normalise(X)
        V = len(self.dofs)
        out = np.zeros(len(xn))
        found = np.zeros(len(xn), dtype=bool)
        for q in range(len(xn)):
            c = np.einsum("sij,sj->si", self._db_Tinv, xn[q] - self._db_r)
            c = np.concatenate((c, 1.0 - c.sum(axis=1, keepdims=True)), axis=1)
            cmin = np.where(np.isnan(c).any(axis=1), -np.inf, c.min(axis=1))
            inside = np.nonzero(cmin >= -DB_EPS)[0]
            s = inside[0] if inside.size else int(np.argmax(cmin))
            found[q] = inside.size > 0
            out[q] = np.dot(c[s], self._db_values[self._db_simplices[s], which])
        return out.reshape(shape), found.reshape(shape)

    def _db_value(self, which, kwargs):
        out, found = self._db_lookup(which, kwargs)
        if not found.all():
            raise DatabaseBoundsError(self.name, np.argwhere(~np.atleast_1d(found)).flatten(), kwargs)
        return out.item() if out.ndim == 0 else out

    def _db_slope(self, dof, kwargs):
        """Central difference of the CL table in `dof`; the two sample points are kept inside the table's limits (one-sided at
        its edges).  Zero if the table does not depend on `dof`."""
        shape = np.broadcast(*[np.asarray(kwargs.get(d, self._DEFAULTS[d]), dtype=float) for d in self.dofs]).shape
        if dof not in self.dofs:
            return np.zeros(shape) if shape else 0.0
        x = np.broadcast_to(np.asarray(kwargs.get(dof, self._DEFAULTS[dof]), dtype=float), shape)
        lo, hi = self.limits[self.dofs.index(dof)]
        up = np.minimum(x + DB_STEP[dof], np.maximum(hi, x)) - x
        dn = np.maximum(x - DB_STEP[dof], np.minimum(lo, x)) - x
        f1, ok1 = self._db_lookup(0, kwargs, (dof, up))
        f0, ok0 = self._db_lookup(0, kwargs, (dof, dn))
        if not (ok1 & ok0).all():
            raise DatabaseBoundsError(self.name, np.argwhere(~np.atleast_1d(ok1 & ok0)).flatten(), kwargs)
        out = (f1 - f0) / (up - dn)
        return out.item() if out.ndim == 0 else out

    def _db_aL0(self, kwargs):
        """Root of the CL table in alpha by the secant method from (0, 0.01) until two iterates agree to 1e-10 (exact on a
        piecewise-linear function once both lie on one piece); 50 iterations at most."""
        kw = {k: v for k, v in kwargs.items() if k != "alpha"}
        shape = np.broadcast(*([np.asarray(kw.get(d, self._DEFAULTS[d]), dtype=float) for d in self.dofs if d != "alpha"] or [0.0])).shape
        if "alpha" not in self.dofs:
            return np.zeros(shape) if shape else 0.0
        a0, a1 = np.zeros(shape), np.full(shape, 0.01)
        f0 = self._db_lookup(0, dict(kw, alpha=a0))[0]
        f1 = self._db_lookup(0, dict(kw, alpha=a1))[0]
        for _ in range(50):
            active = np.abs(a1 - a0) > 1e-10
            if not active.any():
                break
            with np.errstate(divide="ignore", invalid="ignore"):
                a2 = np.where(active, a1 - f1 * (a0 - a1) / (f0 - f1), a1)
            f2 = self._db_lookup(0, dict(kw, alpha=a2))[0]
            a0, f0, a1, f1 = a1, f1, a2, f2
        return a1.item() if a1.ndim == 0 else a1

    def lls_db_pool(self):
        """Pool record of a database airfoil (layout in include/lls.h)."""
        V = len(self.dofs)
        S = len(self._db_simplices)
        head = np.array([float(len(self._db_points)), float(S)])
        lim = np.stack([self.limits[:, 0], self.limits[:, 1] - self.limits[:, 0]], axis=1).ravel()
        simp = np.concatenate([self._db_simplices.astype(float), self._db_Tinv.reshape(S, V * V), self._db_r], axis=1)
        return np.concatenate([head, lim, self._db_values.ravel(), simp.ravel()])

    # -- device export ---------------------------------------------------------------------
    def lls_params(self):
        if self._type in ("poly_fit", "database"):
            blk = np.zeros(L.AF_STRIDE)
            blk[L.AF_TYPE] = 1.0 if self._type == "poly_fit" else 2.0
            blk[L.AF_POLY_NDOF] = float(len(self.dofs))
            for v, d in enumerate(self.dofs):
                blk[L.AF_POLY_DOF0 + v] = float(L.POLY_DOF_KINDS.index(d))
            return blk            # AF_POLY_OFF is filled in by tables.build_tables
        blk = np.zeros(L.AF_STRIDE)
        blk[L.AF_TYPE] = 0.0
        blk[L.AF_AL0] = self.aL0
        blk[L.AF_CLA] = self.CLa
        blk[L.AF_CML0] = self.CmL0
        blk[L.AF_CMA] = self.Cma
        blk[L.AF_CD0] = self.CD0
        blk[L.AF_CD1] = self.CD1
        blk[L.AF_CD2] = self.CD2
        blk[L.AF_CLMAX] = self.CL_max
        return blk

    # -- geometry (stored per control point by the reference, unused by the solve) -----------
    def get_max_camber(self):
        return self._max_camber

    def get_max_thickness(self):
        return self._max_thickness

    def set_err_state(self, **kwargs):
        pass

    # -- coefficients -----------------------------------------------------------------------
    @staticmethod
    def _unpack(kwargs):
        alpha = np.asarray(kwargs.get("alpha", 0.0), dtype=float)
        delta = np.asarray(kwargs.get("trailing_flap_deflection", 0.0), dtype=float)
        c_f = np.asarray(kwargs.get("trailing_flap_fraction", 0.0), dtype=float)
        shape = np.broadcast(alpha, delta, c_f).shape
        return alpha, delta, c_f, shape

    def get_CL(self, **kwargs):
        if self._type == "database":
            return self._db_value(0, kwargs)
        if self._type == "poly_fit":
            return self._poly("CL", kwargs)
        alpha, delta, c_f, _ = self._unpack(kwargs)
        eps, _ = flap_effectiveness(delta, c_f)
        return np.clip(self.CLa * (alpha - self.aL0 + delta * eps), -self.CL_max, self.CL_max)

    def get_CLa(self, **kwargs):
        if self._type == "database":
            return self._db_slope("alpha", kwargs)
        if self._type == "poly_fit":
            return self._poly("CL", kwargs, "alpha")
        _, _, _, shape = self._unpack(kwargs)
        return np.full(shape, self.CLa) if shape else self.CLa

    def get_aL0(self, **kwargs):
        if self._type == "database":
            return self._db_aL0(kwargs)
        if self._type == "poly_fit":
            return self._poly_aL0(kwargs)
        alpha, delta, c_f, shape = self._unpack(kwargs)
        eps, _ = flap_effectiveness(delta, c_f)
        return np.broadcast_to(self.aL0 - delta * eps, shape).copy() if shape else self.aL0 - delta * eps

    def get_CLRe(self, **kwargs):
        if self._type == "database":
            return self._db_slope("Rey", kwargs)
        if self._type == "poly_fit":
            return self._poly("CL", kwargs, "Rey")
        _, _, _, shape = self._unpack(kwargs)
        return np.zeros(shape)

    def get_CLM(self, **kwargs):
        if self._type == "database":
            return self._db_slope("Mach", kwargs)
        if self._type == "poly_fit":
            return self._poly("CL", kwargs, "Mach")
        _, _, _, shape = self._unpack(kwargs)
        return np.zeros(shape)

    def get_CD(self, **kwargs):
        if self._type == "database":
            return self._db_value(1, kwargs)
        if self._type == "poly_fit":
            return self._poly("CD", kwargs)
        alpha, delta, c_f, shape = self._unpack(kwargs)
        cl_clean = np.clip(self.CLa * (alpha - self.aL0), -self.CL_max, self.CL_max)
        cd = self.CD0 + self.CD1 * cl_clean + self.CD2 * cl_clean * cl_clean + 0.002 * np.abs(delta) * 180.0 / np.pi
        return np.broadcast_to(cd, shape).copy() if shape else cd

    def get_Cm(self, **kwargs):
        if self._type == "database":
            return self._db_value(2, kwargs)
        if self._type == "poly_fit":
            return self._poly("Cm", kwargs)
        alpha, delta, c_f, _ = self._unpack(kwargs)
        _, dcm = flap_effectiveness(delta, c_f)
        return self.Cma * (alpha - self.aL0) + self.CmL0 + delta * dcm
