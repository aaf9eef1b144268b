"""Host-side API behaviour that needs no GPU: atmosphere getters, units, flight-state handling, wing-segment geometry and
section-coefficient interpolation.  Expected numbers are the reference's own test constants (file:line cited)."""
import copy
import json
import os

import numpy as np
import pytest

from machupx_b200 import Scene
from machupx_b200.atmosphere import StandardAtmosphere
from machupx_b200.units import read_value

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "input_for_testing_embedded.json")) as fh:
    BASE = json.load(fh)


def _scene(units=None, atmosphere=None, state=None):
    d = copy.deepcopy(BASE)
    if units:
        d["units"] = units
    if atmosphere:
        d["scene"]["atmosphere"].update(atmosphere)
    if state is not None:
        d["scene"]["aircraft"]["test_plane"]["state"] = state
    return Scene(d)


# ---- test/scene_tests/test_get_density.py ------------------------------------------------------------------------
@pytest.mark.parametrize("units,atmos,expected", [
    ("SI", {}, 1.224999155887),                                            # :9-20
    ("English", {}, 0.00237689072965),                                     # :23-34
    ("SI", {"rho": [0.0023769, "slug/ft^3"]}, 1.225003914881),             # :37-49
    ("English", {"rho": [1.225, "kg/m^3"]}, 0.0023768923675),              # :52-64
])
def test_constant_density(units, atmos, expected):
    scene = _scene(units, atmos)
    rng = np.random.default_rng(0)
    for _ in range(5):
        assert np.allclose(scene._get_density(rng.random(3) * 10000), expected, rtol=0.0, atol=1e-10)


def test_standard_atmosphere_profiles():
    scene = _scene("English", {"rho": "standard"})                         # :67-84
    for alt, rho in zip([0, 5000, 10000, 15000, 20000],
                        [0.0023768907296517925, 0.0020481711609077187, 0.0017555489303339121, 0.0014961556118776782,
                         0.001267258247483606]):
        assert np.allclose(scene._get_density(np.array([3.0, 4.0, -alt])), rho, rtol=0.0, atol=1e-10)
    scene = _scene("SI", {"rho": "standard"})                              # :87-104
    for alt, rho in zip([0, 2000, 4000, 6000, 8000],
                        [1.2249991558877122, 1.0065532169786469, 0.8193463086549038, 0.6601112106182213, 0.5257860233001964]):
        assert np.allclose(scene._get_density(np.array([3.0, 4.0, -alt])), rho, rtol=0.0, atol=1e-10)
    atm = StandardAtmosphere("SI")
    assert atm.a(0.0) == pytest.approx(340.294, abs=1e-3) and atm.T(12000.0) == pytest.approx(216.65, abs=1e-9)


def test_density_profile_array():
    prof = [[0.0, 1.0], [1000.0, 0.9], [2000.0, 0.7]]
    scene = _scene("SI", {"rho": prof})
    assert np.allclose(scene._get_density(np.array([0.0, 0.0, -500.0])), 0.95)
    assert np.allclose(scene._get_density(np.array([[0.0, 0.0, -1500.0], [1.0, 1.0, -2500.0]])), [0.8, 0.7])


# ---- test/scene_tests/test_get_wind.py ----------------------------------------------------------------------------
def test_wind_getters():
    assert np.allclose(_scene("SI")._get_wind(np.array([1.0, 2.0, 3.0])), [0, 0, 0], atol=1e-10)                     # :9-20
    assert np.allclose(_scene("SI", {"V_wind": [100, 100, 100, "kn"]})._get_wind(np.array([1.0, 2.0, 3.0])),
                       [51.4444444] * 3, rtol=0.0, atol=1e-10)                                                       # :23-35
    assert np.allclose(_scene("English", {"V_wind": [100, 100, 100, "kph"]})._get_wind(np.array([1.0, 2.0, 3.0])),
                       [91.13444] * 3, rtol=0.0, atol=1e-10)                                                         # :38-50
    prof = [[0.0, 100.0, 100.0, 100.0], [1000.0, 80.0, 120.0, 150.0], [2000.0, 60.0, 140.0, 150.0],
            [3000.0, 40.0, 160.0, 150.0], [5000.0, 20.0, 180.0, 150.0]]
    scene = _scene("SI", {"V_wind": prof})                                                                           # :53-80
    for alt, wind in zip([500, 1500, 2500, 4000, 4500, 6000],
                         [[90, 110, 125], [70, 130, 150], [50, 150, 150], [30, 170, 150], [25, 175, 150], [20, 180, 150]]):
        assert np.allclose(scene._get_wind(np.array([5.0, 6.0, -alt])), wind, rtol=0.0, atol=1e-10)


# ---- test/airplane_tests/test_set_state.py --------------------------------------------------------------------------
def test_velocity_with_wind():
    s = _scene(atmosphere={"V_wind": [100, 0, 0]}, state={"velocity": 100, "alpha": 0.0, "beta": 0.0})               # :9-27
    assert np.allclose(s._airplanes["test_plane"].v, [200.0, 0.0, 0.0], rtol=0.0, atol=1e-10)
    s = _scene(atmosphere={"V_wind": [100, 100, 0]}, state={"velocity": 100, "alpha": 5.0, "beta": 5.0})             # :30-48
    assert np.allclose(s._airplanes["test_plane"].v, [199.24038765, 108.71557427, 8.68240888], rtol=0.0, atol=1e-8)
    ap = s._airplanes["test_plane"]
    a, b, V = ap.get_aerodynamic_state(v_wind=s._get_wind(ap.p_bar))
    assert (a, b, V) == pytest.approx((5.0, 5.0, 100.0), abs=1e-10)


def test_state_forms_and_errors():
    s = _scene(state={"velocity": [100.0, 0.0, 10.0], "orientation": [0.0, 10.0, 0.0], "angular_rates": [0.1, 0.0, 0.0]})
    ap = s._airplanes["test_plane"]
    assert abs(np.linalg.norm(ap.q) - 1.0) < 1e-15 and np.allclose(ap.w, [0.1, 0.0, 0.0])
    with pytest.raises(IOError):
        _scene(state={"velocity": [100.0, 0.0, 10.0], "alpha": 2.0})
    with pytest.raises(IOError):
        _scene(state={"velocity": 100.0, "angular_rate_frame": "nonsense"})
    with pytest.raises(IOError):
        read_value("x", {"x": [1.0, "furlong"]}, "English", None)
    assert read_value("x", {"x": [1.0, "m"]}, "English", None) == pytest.approx(3.28084)
    assert read_value("x", {}, "SI", 2) == 2.0 and isinstance(read_value("x", {}, "SI", 2), float)


# ---- test/wing_segment_tests/test_wing_segment_geometry.py ---------------------------------------------------------------
def test_root_and_tip_locations():
    d = copy.deepcopy(BASE)
    wings = d["scene"]["aircraft"]["test_plane"]["file"]["wings"]
    for key in wings:
        wings[key].pop("sweep", None)
        wings[key].pop("twist", None)
    segs = Scene(d)._airplanes["test_plane"].wing_segments
    roots = {"main_wing_left": [0, 0, 0], "main_wing_right": [0, 0, 0], "h_stab_left": [-3.0, 0, 0],
             "h_stab_right": [-3.0, 0, 0], "v_stab_right": [-3.0, 0, -0.1]}                                          # :31-42
    for key, loc in roots.items():
        assert np.allclose(np.asarray(segs[key].get_root_loc()).flatten(), loc, rtol=0, atol=1e-10)
    tips = {"main_wing_left": [0, -4.0, 0], "main_wing_right": [0, 4.0, 0], "h_stab_left": [-3.0, -2.0, 0],
            "h_stab_right": [-3.0, 2.0, 0], "v_stab_right": [-3.0, 0, -2.1]}                                         # :64-75
    for key, loc in tips.items():
        assert np.allclose(np.asarray(segs[key].get_tip_loc()).flatten(), loc, rtol=0, atol=1e-10)


def test_swept_dihedral_tip_and_wing_grouping():
    d = copy.deepcopy(BASE)
    wings = d["scene"]["aircraft"]["test_plane"]["file"]["wings"]
    wings["main_wing"]["sweep"] = 30.0
    wings["main_wing"]["dihedral"] = 10.0
    ap = Scene(d)._airplanes["test_plane"]
    b = 4.0
    tip = ap.wing_segments["main_wing_right"].get_tip_loc()
    assert np.allclose(tip, [-b * np.tan(np.radians(30.0)), b * np.cos(np.radians(10.0)), -b * np.sin(np.radians(10.0))], atol=1e-12)
    # main wing halves share one lifting line, the horizontal tail halves another, the fin is alone (airplane.py:709-869)
    assert [[s.name for s in w] for w in ap._segments_in_wings] == [["main_wing_left", "main_wing_right"],
                                                                     ["h_stab_left", "h_stab_right"], ["v_stab_right"]]
    assert [(sl.start, sl.stop) for sl in ap.wing_slices] == [(0, 80), (80, 160), (160, 200)]
    assert ap.N == 200 and ap.S_w == 8.0


# ---- test/wing_segment_tests/test_airfoil_interpolation.py --------------------------------------------------------------
def test_section_coefficients_blend_linearly_in_span():
    d = copy.deepcopy(BASE)
    wing = d["scene"]["aircraft"]["test_plane"]["file"]["wings"]["main_wing"]
    wing["airfoil"] = [[0.0, "NACA_0010"], [1.0, "NACA_4410"]]
    seg = Scene(d)._airplanes["test_plane"].wing_segments["main_wing_right"]
    n = seg.N
    alpha, Re, M = np.full(n, 0.05), np.full(n, 1e6), np.full(n, 0.1)
    a0, a1 = seg._airfoils
    s = seg.cp_span_locs
    for getter, name in ((seg.get_cp_CL, "get_CL"), (seg.get_cp_CD, "get_CD"), (seg.get_cp_Cm, "get_Cm"), (seg.get_cp_CLa, "get_CLa")):
        kw = dict(alpha=alpha, Rey=Re, Mach=M, trailing_flap_deflection=seg._delta_flap, trailing_flap_fraction=seg._cp_c_f)
        expect = (1 - s) * getattr(a0, name)(**kw) + s * getattr(a1, name)(**kw)
        assert np.allclose(getter(alpha, Re, M), expect, rtol=0, atol=1e-14)
    assert np.allclose(seg.get_cp_aL0(Re, M), (1 - s) * 0.0 + s * (-0.0746), atol=1e-14)


def test_unsupported_airfoil_types_fail_loudly():
    d = copy.deepcopy(BASE)
    d["scene"]["aircraft"]["test_plane"]["file"]["airfoils"]["NACA_0010"] = {"type": "functional", "CL": lambda alpha: 6.0 * alpha}
    with pytest.raises(NotImplementedError):            # callables cannot run on the device and there is no CPU path
        Scene(d)
    d["scene"]["aircraft"]["test_plane"]["file"]["airfoils"]["NACA_0010"] = {"type": "database", "input_file": "no_such_table.txt"}
    with pytest.raises(IOError):                        # database airfoils are tables on the device; a missing table is an input error
        Scene(d)


# ---- batches: state blocks of a sweep (BASELINE.json configs[2]) ----------------------------------------------------------
def test_sweep_state_blocks_vectorised_equal_per_state():
    """A plain alpha / beta / airspeed sweep is converted to per-case state blocks in one vectorised pass; the blocks must equal
    what Airplane.set_state produces one state at a time (the path every other kind of state takes)."""
    scene = _scene(state={"position": [0.0, 0.0, 0.0], "velocity": 100.0, "alpha": 2.0})
    name = "test_plane"
    rng = np.random.default_rng(3)
    plain = [{"position": [0.0, 0.0, 0.0], "velocity": float(v), "alpha": float(a), "beta": float(b)}
             for v, a, b in zip(rng.uniform(50, 200, 64), rng.uniform(-10, 15, 64), rng.uniform(-12, 12, 64))]
    fast = scene._batch_rows(name, plain)
    # the same states in a form the vectorised pass does not take (an explicit zero rate vector)
    slow = scene._batch_rows(name, [dict(st, angular_rates=[0.0, 0.0, 0.0]) for st in plain])
    assert fast.shape == slow.shape == (64, 1, slow.shape[2])
    assert np.abs(fast - slow).max() <= 1e-13 * np.abs(slow).max()
    # the airplane is left as it was
    assert abs(scene._airplanes[name].get_aerodynamic_state()[0] - 2.0) < 1e-12
    with pytest.raises(ValueError):
        scene._batch_rows(name, [{"position": [1.0, 0.0, 0.0], "velocity": 100.0, "alpha": 1.0}])


# ---- error behaviour of the drop-in API that needs no device (scene.py:1449-1450, 3864-3884) -----------------------------------
def test_empty_scene_and_bad_requests_fail_like_the_reference():
    empty = Scene({"scene": {"atmosphere": {}}})
    with pytest.raises(RuntimeError, match="no aircraft"):
        empty.solve_forces()                                   # scene.py:1449-1450
    with pytest.raises(RuntimeError, match="no aircraft"):
        empty.solve_forces_batch([{"velocity": 10.0}])
    scene = _scene()
    with pytest.raises(ValueError):                            # a batch keeps the geometry: the position may not move
        scene._batch_rows("test_plane", [{"position": [0.0, 0.0, -10.0], "velocity": 100.0}])
    # batched stencils build their totals with the same assembly as solve_forces: every frame is available
    assert scene._batched_ok({"batched": True, "stab_frame": True}) is True
    assert scene._batched_ok({}) is False and scene._batched_ok({"batched": True}) is True


def test_table_revisions_drive_every_device_upload():
    """table freshness is a revision counter, not object identity.  A control change bumps the revision and
    re-uploads five flap rows; a pose change bumps the full revision and re-uploads everything (no GPU: a recording device)."""
    from machupx_b200 import layout as L

    class Recorder:
        def __init__(self):
            self.calls = []

        def upload_tables(self, tables):
            self.calls.append("tables")

        def upload_rows(self, tab, fields):
            self.calls.append(("rows", tuple(fields), float(tab[L.F_FLAP_DEFL].max())))

    scene = _scene()
    dev = Recorder()
    have = scene._upload_if_stale(dev, None)
    assert dev.calls == ["tables"]
    assert scene._upload_if_stale(dev, have) == have and dev.calls == ["tables"]          # nothing changed: nothing uploaded
    scene.set_aircraft_control_state({"elevator": 4.0})
    scene._refresh_tables()
    assert scene._tables["tab"][L.F_FLAP_DEFL].max() == pytest.approx(np.radians(4.0))
    have2 = scene._upload_if_stale(dev, have)
    assert have2 != have and dev.calls[-1][0] == "rows" and dev.calls[-1][1] == L.FLAP_FIELDS
    assert dev.calls[-1][2] == pytest.approx(np.radians(4.0))
    other = Recorder()                                                                      # a second consumer that is two revisions behind
    scene.set_aircraft_control_state({"elevator": -2.0})
    scene._refresh_tables()
    assert scene._upload_if_stale(other, have)[0] == scene._tables_rev and other.calls[-1][0] == "rows"
    scene.set_aircraft_state({"position": [0.0, 0.0, -500.0], "velocity": 100.0, "alpha": 1.0})
    scene._refresh_tables()
    have3 = scene._upload_if_stale(dev, have2)
    assert dev.calls[-1] == "tables" and have3[1] == have2[1] + 1
