"""`"database"` airfoils (airfoil_db's table look-up type) against the reference's own golden.

The reference has ONE test of a database airfoil, test/scene_tests/test_solve_forces.py:453-494 (a swept wing blending a
NACA 2414 table into a NACA 0012 table, nonlinear solve); its nine golden numbers are restated below.  airfoil_db interpolates
on the Delaunay triangulation of the data points, and these tables are exact rectangular grids in (alpha, Rey): every cell is
co-circular and Qhull picks its diagonal arbitrarily, differently from build to build.  Each diagonal in the range the solve
visits moves the forces by 1e-4 ... 7e-3 (relative), so the golden fixes the diagonals of its generation run; they were
recovered by flipping cells until all nine numbers matched to the 16th digit (three cells of the tip table differ from what
the scipy of this image produces) and are enforced here through the `_set_simplices` seam.  With them, the model of
machupx_b200/airfoil.py -- barycentric interpolation, CLa as the central difference with step 0.001 rad, aL0 as the root of the
table -- reproduces the reference to 1e-10 both through the oracle (CPU) and through liblls (GPU): database airfoils are
PINNED.  A second test checks the device against the oracle with scipy's own triangulation left alone.

The two data tables come from the staged copy of the reference's test directory (oracle/stage_ref.py -> oracle/_ref/test;
git-ignored, travels to the GPU box); without it the tests skip."""
import os

import numpy as np
import pytest

import oracle_backend

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(ROOT, "oracle", "_ref")
pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(STAGED, "test", "NACA 2414.txt")),
                                reason="reference test tables not staged (oracle/stage_ref.py needs /root/reference)")

# reference test/scene_tests/test_solve_forces.py:486-494
GOLDEN = {"FL": -1.921431422738798, "FD": 0.6730156268500228, "FS": 0.0, "Fx": -0.622487790173485, "Fy": 0.0,
          "Fz": 1.9383904914534897, "Mx": 0.0, "My": 0.41401938251627074, "Mz": 0.0}
# diagonals of the golden's triangulation between Rey = 5e5 and 1e6; key = lower edge of the cell in half degrees of alpha,
# "/" joins (alpha_lo, Rey_lo) with (alpha_hi, Rey_hi), "\\" the other two corners
GOLDEN_DIAGONALS = {"root": {-5: "/", -4: "\\", -3: "/", -2: "\\", -1: "\\", 0: "/"},
                    "tip": {-5: "/", -4: "\\", -3: "/", -2: "/", -1: "\\", 0: "/"}}


def _enforce_diagonals(af, want, re_lo=5e5, re_hi=1e6):
    X = af._db_points
    half = np.round(np.degrees(X[:, 0]) * 2).astype(int)
    index = {(int(h), float(r)): i for i, (h, r) in enumerate(zip(half, X[:, 1]))}
    simplices = [tuple(int(v) for v in t) for t in af._db_simplices]
    flipped = 0
    for h, diag in want.items():
        c00, c10, c01, c11 = index[(h, re_lo)], index[(h + 1, re_lo)], index[(h, re_hi)], index[(h + 1, re_hi)]
        quad = {c00, c10, c01, c11}
        inside = [t for t in simplices if set(t) <= quad]
        assert len(inside) == 2, (af.name, h, inside)
        new = [(c00, c10, c11), (c00, c11, c01)] if diag == "/" else [(c00, c10, c01), (c10, c11, c01)]
        if {frozenset(t) for t in inside} != {frozenset(t) for t in new}:
            flipped += 1
            simplices = [t for t in simplices if t not in inside] + new
    af._set_simplices(np.asarray(simplices, dtype=np.int64))
    return flipped


def _scene(monkeypatch, pin):
    """The reference's test, statement by statement (test_solve_forces.py:456-484)."""
    import machupx_b200 as MX
    monkeypatch.chdir(STAGED)
    input_dict, name, aircraft, state, control_state = MX.helpers.parse_input("test/input_for_testing.json")
    input_dict["solver"]["type"] = "nonlinear"
    aircraft["airfoils"] = {"root": {"type": "database", "input_file": "test/NACA 2414.txt"},
                            "tip": {"type": "database", "input_file": "test/NACA 0012.txt"}}
    aircraft["wings"]["main_wing"]["sweep"] = 20.0
    aircraft["wings"]["main_wing"]["airfoil"] = [[0.0, "root"], [1.0, "tip"]]
    aircraft["wings"]["main_wing"]["grid"]["N"] = 100
    aircraft["wings"].pop("h_stab")
    aircraft["wings"].pop("v_stab")
    state["alpha"] = -1.5
    if pin:
        from machupx_b200 import airfoil as A
        plain = A.Airfoil._init_database

        def pinned(self, airfoil_input):
            plain(self, airfoil_input)
            _enforce_diagonals(self, GOLDEN_DIAGONALS[self.name])
        monkeypatch.setattr(A.Airfoil, "_init_database", pinned)
    scene = MX.Scene(input_dict)
    scene.add_aircraft(name, aircraft, state=state, control_state=control_state)
    return scene


def _check_golden(scene):
    total = scene.solve_forces(non_dimensional=True)["test_plane"]["total"]
    for key, ref in GOLDEN.items():
        assert abs(total[key] - ref) < 1e-10, (key, total[key], ref)


def test_host_getters_of_a_database_airfoil(monkeypatch):
    from machupx_b200 import DatabaseBoundsError
    from machupx_b200.airfoil import Airfoil
    monkeypatch.chdir(STAGED)
    af = Airfoil("root", {"type": "database", "input_file": "test/NACA 2414.txt"})
    assert af.dofs == ["alpha", "Rey"] and af._db_values.shape[1] == 3
    # data points are reproduced exactly, whatever the triangulation
    k = np.arange(0, len(af._db_points), 37)
    X = af._db_points[k]
    for which, getter in enumerate((af.get_CL, af.get_CD, af.get_Cm)):
        assert np.abs(getter(alpha=X[:, 0], Rey=X[:, 1]) - af._db_values[k, which]).max() < 1e-12
    # along a grid line of constant Rey the interpolant is the 1-D linear interpolation of that line
    line = af._db_points[:, 1] == 1e6
    a = np.linspace(-0.1, 0.15, 41)
    expect = np.interp(a, af._db_points[line, 0], af._db_values[line, 0])
    assert np.abs(af.get_CL(alpha=a, Rey=1e6) - expect).max() < 1e-12
    # central difference, root, scalar calls with airfoil_db's default Rey = 1e6 (the Kuchemann offset asks for CLa(alpha=0))
    assert abs(af.get_CLa(alpha=0.0) - (af.get_CL(alpha=0.001) - af.get_CL(alpha=-0.001)) / 0.002) < 1e-12
    assert isinstance(af.get_CLa(alpha=0.0), float) and abs(af.get_CL(alpha=af.get_aL0())) < 1e-12
    assert af.get_CLM(alpha=a, Rey=1e6, Mach=0.1).tolist() == [0.0] * 41          # the table does not depend on Mach
    assert np.abs(af.get_CLRe(alpha=0.05, Rey=7e5) - (af.get_CL(alpha=0.05, Rey=7.01e5) - af.get_CL(alpha=0.05, Rey=6.99e5)) / 2e3) < 1e-15
    with pytest.raises(DatabaseBoundsError):
        af.get_CL(alpha=0.0, Rey=1e5)
    with pytest.raises(DatabaseBoundsError):
        af.get_CD(alpha=np.array([0.0, 1.0]), Rey=1e6)


def test_reference_golden_through_the_oracle(monkeypatch):
    oracle_backend.install(monkeypatch)
    scene = _scene(monkeypatch, pin=True)
    _check_golden(scene)
    # what the pin is worth: with this image's own triangulation the same solve is 3.5e-4 off in lift
    monkeypatch.undo()
    oracle_backend.install(monkeypatch)
    total = _scene(monkeypatch, pin=False).solve_forces(non_dimensional=True)["test_plane"]["total"]
    assert abs(total["FL"] - GOLDEN["FL"]) < 5e-3 * abs(GOLDEN["FL"])


@pytest.mark.gpu
def test_reference_golden_on_the_device(monkeypatch):
    scene = _scene(monkeypatch, pin=True)
    _check_golden(scene)
    assert scene._last_iterations == 12                       # the reference with the same tables: 12 Newton iterations


@pytest.mark.gpu
def test_device_against_the_oracle_with_scipys_own_triangulation(monkeypatch):
    from machupx_b200.tables import build_aircraft_state
    from oracle import lls_oracle as O
    scene = _scene(monkeypatch, pin=False)
    scene.set_aircraft_state({"alpha": 1.3, "beta": -2.0, "velocity": 170.0})     # Rey ~ 9.5e5, between two levels of the tables
    scene.solve_forces(report_by_segment=True)
    T = O.Tables(scene._tables, build_aircraft_state(scene))
    ref = O.solve(T, solver_type="nonlinear", convergence=scene._solver_convergence, relaxation=scene._solver_relaxation,
                  max_iterations=scene._max_solver_iterations)
    assert scene._last_iterations == ref["iterations"] == 13          # errors 4.1e-10 -> 7.4e-11 around the threshold of 1e-10
    assert np.abs(np.asarray(scene._gamma) - ref["gamma"]).max() <= 1e-10 * np.abs(ref["gamma"]).max()
    assert np.abs(scene._seg_fm - ref["seg_fm"]).max() <= 1e-10 * np.abs(ref["seg_fm"]).max()


@pytest.mark.gpu
def test_device_flags_a_point_outside_the_table(monkeypatch):
    from machupx_b200 import DatabaseBoundsError
    scene = _scene(monkeypatch, pin=False)
    scene.set_aircraft_state({"alpha": 2.0, "velocity": 12.0})          # Rey ~ 7e4, below the tables' 5e5
    with pytest.raises(DatabaseBoundsError):
        scene.solve_forces()
