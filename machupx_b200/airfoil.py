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
    """Kept for API compatibility (machupX/__init__.py:11); the analytic models never raise it."""


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
        if self._type != "linear":
            raise NotImplementedError(
                "Airfoil '{0}': section type '{1}' cannot be evaluated on the device; only 'linear' and 'poly_fit' airfoils "
                "are supported (there is no CPU fallback for table look-ups).".format(name, self._type))
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
        """Pool record of a poly_fit airfoil (layout in include/lls.h)."""
        rec = [self.limits.ravel()]
        for k in ("CL", "CD", "Cm"):
            rec.append(np.concatenate(([float(len(self.fit_coefs[k]))], np.asarray(self.fit_degrees[k], dtype=float), self.fit_coefs[k])))
        return np.concatenate(rec)

    # -- device export ---------------------------------------------------------------------
    def lls_params(self):
        if self._type == "poly_fit":
            blk = np.zeros(L.AF_STRIDE)
            blk[L.AF_TYPE] = 1.0
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
        if self._type == "poly_fit":
            return self._poly("CL", kwargs)
        alpha, delta, c_f, _ = self._unpack(kwargs)
        eps, _ = flap_effectiveness(delta, c_f)
        return np.clip(self.CLa * (alpha - self.aL0 + delta * eps), -self.CL_max, self.CL_max)

    def get_CLa(self, **kwargs):
        if self._type == "poly_fit":
            return self._poly("CL", kwargs, "alpha")
        _, _, _, shape = self._unpack(kwargs)
        return np.full(shape, self.CLa) if shape else self.CLa

    def get_aL0(self, **kwargs):
        if self._type == "poly_fit":
            return self._poly_aL0(kwargs)
        alpha, delta, c_f, shape = self._unpack(kwargs)
        eps, _ = flap_effectiveness(delta, c_f)
        return np.broadcast_to(self.aL0 - delta * eps, shape).copy() if shape else self.aL0 - delta * eps

    def get_CLRe(self, **kwargs):
        if self._type == "poly_fit":
            return self._poly("CL", kwargs, "Rey")
        _, _, _, shape = self._unpack(kwargs)
        return np.zeros(shape)

    def get_CLM(self, **kwargs):
        if self._type == "poly_fit":
            return self._poly("CL", kwargs, "Mach")
        _, _, _, shape = self._unpack(kwargs)
        return np.zeros(shape)

    def get_CD(self, **kwargs):
        if self._type == "poly_fit":
            return self._poly("CD", kwargs)
        alpha, delta, c_f, shape = self._unpack(kwargs)
        cl_clean = np.clip(self.CLa * (alpha - self.aL0), -self.CL_max, self.CL_max)
        cd = self.CD0 + self.CD1 * cl_clean + self.CD2 * cl_clean * cl_clean + 0.002 * np.abs(delta) * 180.0 / np.pi
        return np.broadcast_to(cd, shape).copy() if shape else cd

    def get_Cm(self, **kwargs):
        if self._type == "poly_fit":
            return self._poly("Cm", kwargs)
        alpha, delta, c_f, _ = self._unpack(kwargs)
        _, dcm = flap_effectiveness(delta, c_f)
        return self.Cma * (alpha - self.aL0) + self.CmL0 + delta * dcm
