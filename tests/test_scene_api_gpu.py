"""The reference's own Scene-level tests, replayed through machupx_b200.Scene on the GPU.

tests/api_scenarios.py restates test/scene_tests/test_solve_forces.py, test_derivatives.py, test_other_methods.py,
test_multiple_aircraft.py and test_distributions.py of the reference; oracle/gen_api_golden.py ran the same scenario
functions through the unmodified reference and stored its answers in tests/golden/api_scenarios.json (asserting a few of
the reference's hard-coded constants on the way).  Tolerances are the reference tests' own (1e-10 absolute on forces
and coefficients of O(1-30); 1e-9 on a few sideslip derivatives)."""
import json
import os

import numpy as np
import pytest

import api_scenarios as S

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "api_scenarios.json")) as fh:
    GOLD = json.load(fh)
with open(os.path.join(HERE, "golden", "input_for_testing_embedded.json")) as fh:
    BASE = json.load(fh)


def _scene_cls():
    from machupx_b200 import Scene
    return Scene


def _compare(got, ref, tol, path=""):
    if isinstance(ref, dict):
        assert set(got.keys()) == set(ref.keys()), "%s: keys %s vs %s" % (path, sorted(got.keys()), sorted(ref.keys()))
        for k in ref:
            _compare(got[k], ref[k], tol, path + "/" + str(k))
    elif isinstance(ref, list):
        g, r = np.asarray(got, dtype=float), np.asarray(ref, dtype=float)
        assert g.shape == r.shape, path
        assert np.abs(g - r).max() <= tol * max(1.0, np.abs(r).max()), "%s: %.3e" % (path, np.abs(g - r).max())
    else:
        assert abs(got - ref) <= tol * max(1.0, abs(ref)), "%s: %r vs %r" % (path, got, ref)


@pytest.mark.parametrize("name", sorted(S.SOLVE_CASES))
def test_solve_forces(name):
    res = S.solve_case(_scene_cls(), BASE, name)
    ref = GOLD["solve"][name]
    _compare(res["FM"], ref["FM"], 1e-10)
    g, r = np.asarray(res["gamma"]), np.asarray(ref["gamma"])
    assert np.abs(g - r).max() <= 1e-10 * np.abs(r).max()


@pytest.mark.parametrize("which", ["stability", "damping", "control", "all", "state"])
def test_derivatives(which):
    res = S.derivatives_case(_scene_cls(), BASE, which)
    # finite differences of 1e-10-accurate solves divided by steps of 0.005-0.5: the reference's own tests use 1e-9..1e-10
    _compare(res, GOLD["derivatives"][which], 2e-8 if which in ("damping", "all", "state") else 1e-9)


@pytest.mark.parametrize("which", ["stability", "damping", "control", "all", "state"])
def test_derivatives_batched(which):
    """The same stencils as ONE batched device pass per aircraft (batched=True; control and state stencils carry per-case
    tables): against the reference's answers at the same tolerances."""
    res = S.derivatives_case(_scene_cls(), BASE, which, batched=True)
    _compare(res, GOLD["derivatives"][which], 2e-8 if which in ("damping", "all", "state") else 1e-9)


@pytest.mark.parametrize("which", ["target_CL", "pitch_trim", "pitch_trim_finite_moment", "pitch_trim_orientation", "aero_center"])
def test_other_methods_batched(which):
    res = S.other_case(_scene_cls(), BASE, which, batched=True)
    _compare(res, GOLD["other"][which], 1e-9)


@pytest.mark.parametrize("batched", [False, True])
def test_export_pylot_model(tmp_path, batched):
    """scene.py:3521-3747 (59 solves of the hot path): the file the reference writes, key by key."""
    res = S.pylot_case(_scene_cls(), BASE, str(tmp_path / "pylot.json"), batched=batched)
    ref = GOLD["pylot"]
    assert set(res.keys()) == set(ref.keys())
    for key, val in ref.items():
        if key == "coefficients":
            for k, v in val.items():
                if isinstance(v, dict):
                    for kk, vv in v.items():
                        assert abs(res[key][k][kk] - vv) <= 2e-8 * max(1.0, abs(vv)), (k, kk)
                elif isinstance(v, str):
                    assert res[key][k] == v
                else:
                    assert abs(res[key][k] - v) <= 2e-8 * max(1.0, abs(v)), k
        else:
            assert res[key] == val, key


@pytest.mark.parametrize("solver", ["linear", "nonlinear"])
def test_c3_sweep_sub_grid_against_the_reference(solver):
    """BASELINE.json configs[2]: the 8 x 8 sub-grid of the 64 x 64 alpha / beta sweep of examples/FlyingWing that
    oracle/gen_api_golden.py ran through the reference's serial loop; here ONE batched device pass (Scene.solve_forces_batch).
    All 24 totals to 1e-10 relative, Newton iteration counts equal."""
    import copy
    gold = GOLD["c3"]
    inp = copy.deepcopy(gold["input"])
    inp["solver"]["type"] = solver
    inp["solver"].pop("relaxation", None)
    scene = _scene_cls()(inp)
    name = gold[solver]["aircraft"]
    res = scene.solve_forces_batch(S.c3_states(gold["alpha_idx"], gold["beta_idx"]), aircraft=name)
    assert res["converged"].all()
    ref = gold[solver]["totals"]
    for key in ref[0]:
        r = np.array([t[key] for t in ref])
        g = np.asarray(res["totals"][name][key])
        assert np.abs(g - r).max() <= 1e-10 * max(1.0, np.abs(r).max()), key
    if solver == "nonlinear":
        assert [int(x) for x in res["iterations"]] == gold[solver]["iterations"]


@pytest.mark.parametrize("which", ["target_CL", "pitch_trim", "pitch_trim_finite_moment", "pitch_trim_orientation", "aero_center", "MAC"])
def test_other_methods(which):
    res = S.other_case(_scene_cls(), BASE, which)
    _compare(res, GOLD["other"][which], 1e-9)


@pytest.mark.parametrize("separation", [25, 10000])
def test_two_aircraft(separation):
    res = S.two_aircraft_case(_scene_cls(), BASE, separation)
    _compare(res, GOLD["two_aircraft"][str(separation)], 1e-10)
    a, b = res["test_plane_0"], res["test_plane_1"]
    for key, sign in (("FL", 1), ("FD", 1), ("FS", -1), ("Fx", 1), ("Fy", -1), ("Fz", 1), ("Mx", -1), ("My", 1), ("Mz", -1)):
        if separation == 25:    # mirror images (test_multiple_aircraft.py:30-38)
            assert abs(a[key] - sign * b[key]) < 1e-6


def test_distributions(tmp_path):
    Scene = _scene_cls()
    res = S.distributions_case(Scene, BASE)
    _compare(res, GOLD["distributions"], 1e-10)
    # CSV export: header and %20.12e columns (scene.py:3244-3250)
    scene = Scene(json.loads(json.dumps(BASE)))
    out = tmp_path / "dist.csv"
    scene.distributions(filename=str(out))
    lines = out.read_text().splitlines()
    assert lines[0].startswith("aircraft,segment,span_fraction,control_x") and lines[0].endswith("Circ,CD_i")
    assert len(lines) == 1 + 200 and len(lines[1].split(",")) == 33


def test_error_policy_and_unsupported_inputs():
    Scene = _scene_cls()
    from machupx_b200 import SolverNotConvergedError
    d = json.loads(json.dumps(BASE))
    d["solver"]["max_iterations"] = 2
    scene = Scene(d)
    with pytest.raises(SolverNotConvergedError):
        scene.solve_forces()
    scene.set_err_state(not_converged="ignore")
    FM = scene.solve_forces()          # integrates with the last circulation, like the reference (scene.py:1494-1503)
    assert "test_plane" in FM
    with pytest.warns(UserWarning):
        scene.set_err_state(not_converged="warn")
        scene.solve_forces()
    d = json.loads(json.dumps(BASE))
    d["solver"]["type"] = "scipy_fsolve"
    with pytest.raises(NotImplementedError):
        Scene(d)
    with pytest.raises(RuntimeError):
        Scene({}).solve_forces()
