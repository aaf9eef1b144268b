"""Linear section model of airfoil_db.Airfoil, restated from its documented behaviour.

Sources for the formulas (file:line in /root/reference):
  * machupX/obsolete/airfoil.py:141-279  (CL, CD, Cm, aL0 of the linear model)
  * machupX/wing_segment.py:664-676      (Phillips flap effectiveness, hinge efficiency, Cm_delta)
  * docs/source/creating_input_files.md:188-218 (input keys and defaults)
The deflection-efficiency roll-off above 11 deg is NOT pinned by any reference test.
"""
import json
import numpy as np

_DEFL_KNEE = 0.19198621771937624  # 11 deg in rad


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
        if self._type != "linear":
            raise NotImplementedError("oracle airfoil_db stub supports only 'linear' airfoils, got %r" % self._type)
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
