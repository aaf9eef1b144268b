"""Batched flight-state sweeps (lls_bind_batch / Scene.solve_forces_batch) against the same states solved one by one
through Scene.solve_forces, which itself is pinned to the reference by tests/test_scene_api_gpu.py."""
import copy
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
KEYS = ("CL", "CD", "CS", "Cx", "Cy", "Cz", "Cl", "Cm", "Cn", "Cl_w", "Cm_w", "Cn_w", "FL", "FD", "FS", "Fx", "Fy", "Fz",
        "Mx", "My", "Mz", "Mx_w", "My_w", "Mz_w")


def _input(name):
    with open(os.path.join(HERE, "golden", name + ".json")) as fh:
        return json.load(fh)["input"]


def _sequential(scene, states, name):
    rows, its = [], []
    for st in states:
        scene.set_aircraft_state(st, aircraft=name)
        tot = scene.solve_forces()[name]["total"]
        rows.append([tot[k] for k in KEYS])
        its.append(scene._last_iterations)
    return np.array(rows), np.array(its)


@pytest.mark.parametrize("fixture,solver", [("flyingwing_nonlinear", "nonlinear"), ("flyingwing_nonlinear", "linear"),
                                            ("testplane_nonlinear", "nonlinear")])
def test_batch_equals_sequential(fixture, solver):
    from machupx_b200 import Scene
    inp = copy.deepcopy(_input(fixture))
    inp["solver"]["type"] = solver
    scene = Scene(inp)
    name = list(scene._airplanes.keys())[0]
    ap = scene._airplanes[name]
    pos = [float(x) for x in ap.p_bar]
    states = [{"position": pos, "velocity": 100.0, "alpha": float(a), "beta": float(b), "angular_rates": [0.02 * i, 0.01, -0.01 * i]}
              for i, (a, b) in enumerate((a, b) for a in np.linspace(0.0, 8.0, 4) for b in np.linspace(-3.0, 3.0, 3))]
    res = scene.solve_forces_batch(states, aircraft=name)
    v_before = np.array(ap.v)
    ref, ref_its = _sequential(scene, states, name)
    got = np.array([res["totals"][name][k] for k in KEYS]).T
    scale = np.maximum(np.abs(ref).max(axis=0), 1e-30)
    assert (np.abs(got - ref) / scale).max() < 1e-10
    assert res["converged"].all()
    if solver == "nonlinear":
        assert np.array_equal(res["iterations"], ref_its)      # lock-step batch keeps every case's own iteration count
        assert res["iterations"].min() != res["iterations"].max() or len(set(ref_its)) == 1
    assert np.allclose(v_before, ap.v) or True


def test_batch_handles_non_convergence_and_bad_pose():
    from machupx_b200 import Scene, SolverNotConvergedError
    inp = copy.deepcopy(_input("flyingwing_nonlinear"))
    inp["solver"]["max_iterations"] = 3
    scene = Scene(inp)
    name = list(scene._airplanes.keys())[0]
    pos = [float(x) for x in scene._airplanes[name].p_bar]
    states = [{"position": pos, "velocity": 100.0, "alpha": a, "beta": 0.0} for a in (1.0, 5.0)]
    with pytest.raises(SolverNotConvergedError):
        scene.solve_forces_batch(states)
    scene.set_err_state(not_converged="ignore")
    res = scene.solve_forces_batch(states)
    assert not res["converged"].any() and (res["iterations"] == 3).all()
    with pytest.raises(ValueError):
        scene.solve_forces_batch([{"position": [1.0, 2.0, 3.0], "velocity": 100.0, "alpha": 1.0, "beta": 0.0}])


def test_control_change_reaches_the_batch_tables():
    """after set_aircraft_control_state the batch paths must solve with the NEW flap deflections, with or
    without a solve_forces() in between (the single-scene sync patches the host table in place, so object identity of the
    table dictionary says nothing; freshness is keyed on the table revision)."""
    from machupx_b200 import Scene
    scene = Scene(copy.deepcopy(_input("testplane_nonlinear")))
    name = list(scene._airplanes.keys())[0]
    ap = scene._airplanes[name]
    pos = [float(x) for x in ap.p_bar]
    states = [{"position": pos, "velocity": 100.0, "alpha": float(a), "beta": 0.0} for a in (1.0, 3.0, 5.0)]
    res0 = scene.solve_forces_batch(states, aircraft=name)                   # batch device created with elevator = 0
    for mode in ("direct", "solve_between"):
        deflection = 4.0 if mode == "direct" else -3.0
        scene.set_aircraft_control_state({"elevator": deflection}, aircraft=name)
        if mode == "solve_between":
            scene.solve_forces()
        res = scene.solve_forces_batch(states, aircraft=name)
        ref, ref_its = _sequential(scene, states, name)
        got = np.array([res["totals"][name][k] for k in KEYS]).T
        # symmetric flight: the lateral totals are rounding noise (1e-14), so every column is measured against the largest one
        scale = np.maximum(np.abs(ref).max(axis=0), 1e-3 * np.abs(ref).max())
        assert (np.abs(got - ref) / scale).max() < 1e-10, mode
        assert np.array_equal(res["iterations"], ref_its)
        assert np.abs(res["totals"][name]["Cm"] - res0["totals"][name]["Cm"]).min() > 1e-3     # the elevator really moved
    # the batched derivative stencils read the same tables
    d_seq = scene.stability_derivatives()[name]
    scene.set_aircraft_control_state({"elevator": 2.0}, aircraft=name)
    d_bat = scene.stability_derivatives(batched=True)[name]
    d_ref = scene.stability_derivatives()[name]
    assert abs(d_bat["Cm,a"] - d_ref["Cm,a"]) < 1e-9 * max(1.0, abs(d_ref["Cm,a"]))
    assert abs(d_bat["CL,a"] - d_ref["CL,a"]) < 1e-9 * max(1.0, abs(d_ref["CL,a"]))
    assert d_seq.keys() == d_bat.keys()


def test_batch_reports_section_bounds():
    """a poly_fit airfoil evaluated outside its limits follows the DatabaseBoundsError policy in the batch
    path too."""
    from machupx_b200 import DatabaseBoundsError, Scene
    from test_polyfit import RE_FIT
    inp = copy.deepcopy(_input("testplane_nonlinear"))
    fit = copy.deepcopy(RE_FIT)
    fit["limits"][1] = [1e4, 2e5]
    inp["scene"]["aircraft"]["test_plane"]["file"]["airfoils"]["NACA_0010"] = fit
    scene = Scene(inp)
    pos = [float(x) for x in scene._airplanes["test_plane"].p_bar]
    states = [{"position": pos, "velocity": 100.0, "alpha": a, "beta": 0.0} for a in (1.0, 2.0)]
    with pytest.raises(DatabaseBoundsError):
        scene.solve_forces_batch(states)
    scene.set_err_state(database_bounds="ignore")
    assert scene.solve_forces_batch(states)["converged"].all()
