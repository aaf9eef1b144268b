"""`machupX.helpers` under the reference's own names, so that user scripts and the reference's test-suite
(`MX.helpers.parse_input(...)` in 17 of its tests) run unchanged against this package.

Every function here is a thin adapter onto `units.py` / `rotations.py`; the behavioural contract is the reference's
machupX/helpers.py (file:line per function below).  Nothing on the solve path calls this module.
"""
import copy
import json

import numpy as np

from . import rotations as _rot
from . import units as _units


def check_filepath(input_filename, correct_ext):
    """helpers.py:9-14: IOError for a wrong extension or a missing file."""
    _units.require_file(input_filename, correct_ext)


def convert_units(in_value, units, system):
    """helpers.py:17-71: value in `units` -> the default unit of `system` ("English" or anything else = SI)."""
    return in_value * _units.unit_factor(units, system) if units != "-" else in_value


vectorized_convert_units = np.vectorize(convert_units)


def import_value(key, dict_of_vals, system, default_value):
    """helpers.py:74-145: fetch + normalise one input value (`default_value=None` marks a mandatory key)."""
    return _units.read_value(key, dict_of_vals, system, default_value)


def quat_trans(q, v):
    """helpers.py:151-166: earth-frame vector(s) v, last axis of length 3, expressed in the frame of orientation q."""
    return _rot.earth_to_body(q, v)


def quat_inv_trans(q, v):
    """helpers.py:169-184: inverse of `quat_trans`."""
    return _rot.body_to_earth(q, v)


def quat_mult(quat0, quat1):
    """helpers.py:187-209: Hamilton product; 3-vectors count as pure quaternions."""
    return _rot.quat_product(quat0, quat1)


def euler_to_quat(E):
    """helpers.py:212-231."""
    return _rot.euler_to_quat(E)


def quat_to_euler(q):
    """helpers.py:234-258."""
    return _rot.quat_to_euler(q)


def quat_conj(q):
    """helpers.py:261-262: a plain list, as in the reference."""
    return [q[0], -q[1], -q[2], -q[3]]


def parse_input(mux_input):
    """helpers.py:265-315: splits a scene input (dict or path of a JSON file) into the scene dictionary without its
    aircraft, and the names / airplane dictionaries / states / control states of the aircraft -- scalars instead of
    lists when there is exactly one aircraft."""
    if isinstance(mux_input, str):
        with open(mux_input, "r") as handle:
            scene_dict = json.load(handle)
    else:
        scene_dict = copy.deepcopy(mux_input)
    entries = scene_dict["scene"].pop("aircraft", {})
    names, airplanes, states, controls = [], [], [], []
    for name, entry in entries.items():
        names.append(name)
        with open(entry["file"], "r") as handle:
            airplanes.append(json.load(handle))
        states.append(entry.get("state", {}))
        controls.append(entry.get("control_state", {}))
    if len(names) == 1:
        return scene_dict, names[0], airplanes[0], states[0], controls[0]
    return scene_dict, names, airplanes, states, controls
