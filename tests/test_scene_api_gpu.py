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


@pytest.mark.parametrize("which", ["stability", "damping", "all"])
def test_derivatives_batched(which):
    """The same stencils as ONE batched device pass per aircraft (batched=True): against the reference's answers at the same
    tolerances, and against the one-solve-at-a-time path."""
    import copy
    scene = _scene_cls()(copy.deepcopy(BASE))
    call = {"stability": scene.stability_derivatives, "damping": scene.damping_derivatives, "all": scene.derivatives}[which]
    res = call(batched=True)
    seq = call()
    _compare(res, seq, 1e-9)
    if which != "all":      # the golden "all" case also carries the stability frame, which a batch does not report
        _compare(res, GOLD["derivatives"][which], 2e-8 if which == "damping" else 1e-9)
    with pytest.raises(NotImplementedError):
        scene.stability_derivatives(batched=True, stab_frame=True)


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
