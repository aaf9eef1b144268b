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
        if self._type != "linear":
            raise NotImplementedError(
                "Airfoil '{0}': section type '{1}' cannot be evaluated on the device; only 'linear' airfoils are "
                "supported (there is no CPU fallback for table look-ups).".format(name, self._type))
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

    # -- device export ---------------------------------------------------------------------
    def lls_params(self):
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
        alpha, delta, c_f, _ = self._unpack(kwargs)
        eps, _ = flap_effectiveness(delta, c_f)
        return np.clip(self.CLa * (alpha - self.aL0 + delta * eps), -self.CL_max, self.CL_max)

    def get_CLa(self, **kwargs):
        _, _, _, shape = self._unpack(kwargs)
        return np.full(shape, self.CLa) if shape else self.CLa

    def get_aL0(self, **kwargs):
        alpha, delta, c_f, shape = self._unpack(kwargs)
        eps, _ = flap_effectiveness(delta, c_f)
        return np.broadcast_to(self.aL0 - delta * eps, shape).copy() if shape else self.aL0 - delta * eps

    def get_CLRe(self, **kwargs):
        _, _, _, shape = self._unpack(kwargs)
        return np.zeros(shape)

    def get_CLM(self, **kwargs):
        _, _, _, shape = self._unpack(kwargs)
        return np.zeros(shape)

    def get_CD(self, **kwargs):
        alpha, delta, c_f, shape = self._unpack(kwargs)
        cl_clean = np.clip(self.CLa * (alpha - self.aL0), -self.CL_max, self.CL_max)
        cd = self.CD0 + self.CD1 * cl_clean + self.CD2 * cl_clean * cl_clean + 0.002 * np.abs(delta) * 180.0 / np.pi
        return np.broadcast_to(cd, shape).copy() if shape else cd

    def get_Cm(self, **kwargs):
        alpha, delta, c_f, _ = self._unpack(kwargs)
        _, dcm = flap_effectiveness(delta, c_f)
        return self.Cma * (alpha - self.aL0) + self.CmL0 + delta * dcm
