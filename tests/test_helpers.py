"""`machupx_b200.helpers`: the reference's `machupX.helpers` names (helpers.py:9-315) as adapters onto units.py / rotations.py.
Expected numbers are the reference's own conversion constants and textbook quaternion identities."""
import json

import numpy as np
import pytest

from machupx_b200 import helpers as H


def test_unit_conversion_and_value_import():
    assert H.convert_units(2.0, "m", "English") == pytest.approx(2.0 * 3.28084, rel=0, abs=1e-15)      # helpers.py:25
    assert H.convert_units(3.0, "kn", "SI") == pytest.approx(3.0 * 0.514444444, rel=0, abs=1e-15)      # helpers.py:52
    assert H.convert_units(5.0, "-", "SI") == 5.0
    with pytest.raises(IOError):
        H.convert_units(1.0, "parsec", "SI")
    assert np.allclose(H.vectorized_convert_units(np.array([1.0, 2.0]), "in", "English"), [0.083333333, 0.166666666])
    assert H.import_value("x", {"x": [10.0, "m/s"]}, "English", None) == pytest.approx(32.8084)
    assert H.import_value("x", {}, "SI", 4) == 4.0
    assert H.import_value("c", {"c": ["elliptic", 1.0, "m"]}, "English", None) == ("elliptic", pytest.approx(3.28084))
    table = H.import_value("t", {"t": [[0.0, 1.0], [1.0, 2.0], ["-", "m"]]}, "English", None)
    assert np.allclose(table, [[0.0, 3.28084], [1.0, 6.56168]])
    with pytest.raises(IOError):
        H.import_value("missing", {}, "SI", None)
    with pytest.raises(IOError):
        H.check_filepath("table.txt", ".csv")
    with pytest.raises(IOError):
        H.check_filepath("no_such_file.csv", ".csv")


def test_quaternion_helpers():
    E = np.radians([10.0, -20.0, 35.0])
    q = H.euler_to_quat(E)
    assert abs(np.linalg.norm(q) - 1.0) < 1e-15 and np.allclose(H.quat_to_euler(q), E, atol=1e-14)
    v = np.array([[1.0, 2.0, 3.0], [-0.5, 0.25, 4.0]])
    body = H.quat_trans(q, v)
    assert body.shape == v.shape and np.allclose(H.quat_inv_trans(q, body), v, atol=1e-14)
    assert np.allclose(np.linalg.norm(body, axis=1), np.linalg.norm(v, axis=1))
    # q^-1 (0, v) q written out with the Hamilton product (helpers.py:151-166 does the same in two steps)
    manual = H.quat_mult(H.quat_mult(H.quat_conj(q), v[0]), q)
    assert abs(manual[0]) < 1e-14 and np.allclose(manual[1:], body[0], atol=1e-14)
    # a pure yaw of 90 degrees: the earth x axis is the body's -y axis
    yaw = H.euler_to_quat([0.0, 0.0, np.pi / 2])
    assert np.allclose(H.quat_trans(yaw, [1.0, 0.0, 0.0]), [0.0, -1.0, 0.0], atol=1e-15)
    assert H.quat_conj([1.0, 2.0, 3.0, 4.0]) == [1.0, -2.0, -3.0, -4.0]


def test_parse_input(tmp_path):
    plane = {"CG": [0, 0, 0], "weight": 10.0, "wings": {}}
    (tmp_path / "plane.json").write_text(json.dumps(plane))
    scene = {"tag": "t", "scene": {"atmosphere": {}, "aircraft": {"a": {"file": str(tmp_path / "plane.json"), "state": {"velocity": 10.0}}}}}
    (tmp_path / "scene.json").write_text(json.dumps(scene))
    stripped, name, airplane, state, controls = H.parse_input(str(tmp_path / "scene.json"))
    assert "aircraft" not in stripped["scene"] and name == "a" and airplane == plane and state == {"velocity": 10.0} and controls == {}
    scene["scene"]["aircraft"]["b"] = {"file": str(tmp_path / "plane.json"), "control_state": {"elevator": 1.0}}
    stripped, names, airplanes, states, controls = H.parse_input(scene)
    assert names == ["a", "b"] and len(airplanes) == 2 and states[1] == {} and controls[1] == {"elevator": 1.0}
    assert "aircraft" in scene["scene"]                      # the caller's dictionary is left alone
