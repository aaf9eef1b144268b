"""1976 US Standard Atmosphere (geometric height in, English or SI out).

Same layer table, constants and unit factors as the reference's machupX/standard_atmosphere.py:23-33, 45-52, 85-88,
105-107, 122-124, 138-140 — the default density/viscosity/speed of sound of every scene come from here
(scene.py:117, 219, 245), so the rounded conversion constants are part of the parity contract.
"""
import numpy as np

_H_BASE = np.array([0.0, 11000.0, 20000.0, 32000.0, 47000.0, 51000.0, 71000.0, 84852.0])    # geopotential m
_LAPSE = np.array([-0.0065, 0.0, 0.001, 0.0028, 0.0, -0.0028, -0.002])                       # K/m
_T_OFFSET = np.array([0.0, -71.5, -71.5, -59.5, -17.5, -17.5, -73.5, -101.204])              # K relative to sea level
_R_EARTH = 6356766.0
_G0 = 9.80665
_M0 = 28.9644
_R_STAR = 8.31432e3
_P0 = 1.01325e5
_SUTHERLAND_S = 110.4
_SUTHERLAND_BETA = 1.458e-6
_GAMMA = 1.40
_PA_TO_PSF = 0.02088543815038
_KGM3_TO_SLUGFT3 = 0.0019403203
_PAS_TO_PSFS = 0.020885434273039


class StandardAtmosphere:
    def __init__(self, unit_sys, atmos_type="standard"):
        self._english = unit_sys == "English"
        self._unit_sys = unit_sys
        self._T_sea_level_C = 15

    # -- helpers ------------------------------------------------------------------------------
    def _geopotential(self, h):
        h_arr = np.asarray(h, dtype=float)
        if not self._english and (h_arr > 86000.0).any():
            raise IOError("Standard atmosphere only goes up to 86 km.")
        Z = h_arr * 0.3048 if self._english else h_arr
        H = (_R_EARTH * Z) / (_R_EARTH + Z)
        return H.item() if H.ndim == 0 else H

    def _kelvin(self, h):
        return np.interp(self._geopotential(h), _H_BASE, _T_OFFSET) + self._T_sea_level_C + 273.15

    def _pascal(self, h):
        H = self._geopotential(h)
        scalar = np.ndim(H) == 0
        H_arr = np.atleast_1d(np.asarray(H, dtype=float))
        P = np.ones(H_arr.shape) * _P0
        for idx, Hi in enumerate(H_arr):
            layer = 0
            while layer < 6 and Hi > _H_BASE[layer]:
                T_base = self._T_sea_level_C + _T_OFFSET[layer] + 273.15
                top = min(Hi, _H_BASE[layer + 1])
                if abs(_LAPSE[layer]) < 1e-6:
                    P[idx] *= np.exp(-_G0 * _M0 * (top - _H_BASE[layer]) / (_R_STAR * T_base))
                else:
                    expo = (_G0 * _M0) / (_R_STAR * _LAPSE[layer])
                    P[idx] *= (T_base / (T_base + _LAPSE[layer] * (top - _H_BASE[layer]))) ** expo
                layer += 1
        return P.item() if scalar else P

    # -- public API (same method names as the reference) ------------------------------------------
    def T(self, h):
        T = self._kelvin(h)
        return 9.0 / 5.0 * T if self._english else T

    def P(self, h):
        P = self._pascal(h)
        return P * _PA_TO_PSF if self._english else P

    def rho(self, h):
        # the reference converts to English and back (standard_atmosphere.py:96-100): keep the same rounding path
        P, T = self.P(h), self.T(h)
        if self._english:
            P = P / _PA_TO_PSF
            T = T * 5.0 / 9.0
        rho = P * _M0 / (_R_STAR * T)
        return rho * _KGM3_TO_SLUGFT3 if self._english else rho

    def mu(self, h):
        T = self.T(h)
        if self._english:
            T = T * 5.0 / 9.0
        mu = _SUTHERLAND_BETA * T ** 1.5 / (T + _SUTHERLAND_S)
        return mu * _PAS_TO_PSFS if self._english else mu

    def a(self, h):
        T = self.T(h)
        if self._english:
            T = T * 5.0 / 9.0
        a = np.sqrt(_GAMMA * _R_STAR * T / _M0)
        return a / 0.3048 if self._english else a

    def nu(self, h):
        return self.mu(h) / self.rho(h)
