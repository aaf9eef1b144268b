"""Input-value parsing and unit conversion for the JSON input format.

Behavioural contract: reference machupX/helpers.py:17-148 (``convert_units`` / ``import_value``): a value may be a
bare number, ``[number, "unit"]``, a 3/4-vector with an optional trailing unit string, a 2-D list (distribution
table, optionally ending in a row of unit strings), a path to a .csv holding such a table, ``["elliptic", c_root(,
unit)]``, a plain string, a callable or an ndarray.  The conversion factors below are the reference's own rounded
constants (helpers.py:22-63) — parity depends on using exactly these, not more precise ones.
"""
import os

import numpy as np

_TO_ENGLISH = {
    "ft": 1.0, "in": 0.083333333, "m": 3.28084, "cm": 0.0328084,
    "ft/s": 1.0, "m/s": 3.28084, "mph": 1.466666666, "kph": 0.9113444, "kn": 1.687811,
    "ft^2": 1.0, "m^2": 10.76391,
    "slug/ft^3": 1.0, "kg/m^3": 0.0019403203,
    "lbf": 1.0, "N": 0.22480894244319,
    "deg": 1.0, "rad": 57.29578, "deg/s": 0.01745329, "rad/s": 1.0,
}
_TO_SI = {
    "ft": 0.3048, "in": 0.0254, "m": 1.0, "cm": 0.01,
    "ft/s": 0.3048, "m/s": 1.0, "mph": 0.44704, "kph": 0.277777777, "kn": 0.514444444,
    "ft^2": 0.09290304, "m^2": 1.0,
    "slug/ft^3": 515.378819, "kg/m^3": 1.0,
    "lbf": 4.4482216, "N": 1.0,
    "deg": 1.0, "rad": 57.29578, "deg/s": 0.01745329, "rad/s": 1.0,
}


def require_file(path, extension):
    if extension not in path:
        raise IOError("File {0} has the wrong extension. Expected a {1} file.".format(path, extension))
    if not os.path.exists(path):
        raise IOError("Cannot find file {0}.".format(path))


def unit_factor(unit, system):
    if unit == "-":
        return 1.0
    table = _TO_ENGLISH if system == "English" else _TO_SI
    try:
        return table[unit.strip(" \t\r\n")]
    except KeyError:
        raise IOError("Improper units specified; {0} is not an allowable unit definition.".format(unit))


def _convert_table(rows, system):
    """2-D list / object array -> float array, applying a trailing row of unit strings column by column."""
    last = rows[-1]
    if isinstance(last[0], str):
        arr = np.asarray(rows)
        data = arr[:-1, :].astype(float)
        factors = np.array([unit_factor(str(u), system) for u in arr[-1, :]])
        return data * factors[np.newaxis, :]
    return np.asarray(rows)


def read_value(key, source, system, default):
    """Fetches ``source[key]`` (or ``default``) and normalises it as described in the module docstring.
    ``default=None`` marks a mandatory key."""
    raw = source.get(key, default)
    if raw is None:
        raise IOError("Key '{0}' is not optional. Please specify.".format(key))
    if isinstance(raw, bool):
        return float(raw)
    if isinstance(raw, float):
        return raw
    if isinstance(raw, int):
        return float(raw)
    if isinstance(raw, str):
        if ".csv" in raw:
            require_file(raw, ".csv")
            with open(raw, "r") as handle:
                table = np.genfromtxt(handle, delimiter=",", dtype=None, encoding="utf-8")
            return _convert_table(table, system)
        return raw
    if isinstance(raw, list):
        if any(isinstance(row, list) for row in raw):
            return _convert_table(raw, system)
        if raw[0] == "elliptic":
            root_chord = raw[1] * unit_factor(raw[2], system) if len(raw) == 3 else raw[1]
            return ("elliptic", root_chord)
        if isinstance(raw[-1], str):
            scaled = np.asarray(raw[:-1], dtype=float) * unit_factor(raw[-1], system)
            return scaled.item() if scaled.size == 1 else scaled
        if len(raw) in (3, 4):
            return np.asarray(raw)
        raise ValueError("Did not recognize value format {0}.".format(raw))
    if callable(raw) or isinstance(raw, np.ndarray):
        return raw
    raise ValueError("Did not recognize value format {0}.".format(raw))
