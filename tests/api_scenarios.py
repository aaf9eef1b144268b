"""Scene-level scenarios shared by the golden generator (oracle/gen_api_golden.py, which runs them through the
UNMODIFIED reference) and the GPU tests (tests/test_scene_api_gpu.py, which run them through machupx_b200.Scene).

Each scenario mirrors one of the reference's own tests (file:line cited) or one of its public drivers, takes the
``Scene`` class to use and the embedded base input of the reference's unit-test aeroplane
(tests/golden/input_for_testing_embedded.json = test/input_for_testing.json with test/airplane_for_testing.json
inlined), and returns plain JSON-able results."""
import copy

import numpy as np

BASE_STATE = {"position": [0, 0, 1000], "angular_rates": [0.0, 0.0, 0.0], "velocity": 100, "alpha": 2.0, "beta": 0.0}
NO_CONTROLS = {"elevator": 0.0, "rudder": 0.0, "aileron": 0.0}


def _with(base, solver=None, state=None, controls=None, wings=None, solver_opts=None):
    d = copy.deepcopy(base)
    ac = d["scene"]["aircraft"]["test_plane"]
    if solver is not None:
        d["solver"]["type"] = solver
    if solver_opts:
        d["solver"].update(solver_opts)
    if state is not None:
        ac["state"] = copy.deepcopy(state)
    if controls is not None:
        ac["control_state"] = copy.deepcopy(controls)
    for wing, mods in (wings or {}).items():
        for key, val in mods.items():
            if key == "grid":
                ac["file"]["wings"][wing].setdefault("grid", {}).update(val)
            else:
                ac["file"]["wings"][wing][key] = val
    return d


def _total(FM, name="test_plane"):
    return dict(FM[name]["total"])


# ---- test/scene_tests/test_solve_forces.py -----------------------------------------------------------------
SOLVE_CASES = {
    "linear_solver": dict(solver="linear", state=BASE_STATE, controls=NO_CONTROLS),                                      # :11-46
    "linear_solver_with_rotation": dict(solver="linear", state=dict(BASE_STATE, angular_rates=[0.2, 0.2, 0.2]), controls=NO_CONTROLS),  # :49-84
    "nonlinear_solver": dict(solver="nonlinear", state=BASE_STATE, controls=NO_CONTROLS),                                # :87-122
    "sideslip": dict(solver="linear", state=dict(BASE_STATE, beta=3.0), controls=NO_CONTROLS),                           # :125-160
    "angular_rotation": dict(solver="nonlinear", state=dict(BASE_STATE, angular_rates=[0.0, 0.1, 0.1], beta=3.0), controls=NO_CONTROLS),  # :163-198
    "flap_deflection": dict(solver="linear", state=BASE_STATE, controls={"elevator": -2.0, "rudder": 0.0, "aileron": 3.0}),  # :201-236
    "swept_wing": dict(solver="nonlinear", wings={"main_wing": {"sweep": 30.0}}),                                        # :277-301
    "swept_wing_with_controls": dict(solver="nonlinear", wings={"main_wing": {"sweep": 30.0}, "h_stab": {"sweep": 30.0}},
                                     controls={"elevator": 5.0, "aileron": 5.0, "rudder": 5.0}),                        # :304-335
    "tapered_wing": dict(solver="nonlinear", wings={"main_wing": {"chord": [[0.0, 1.0], [1.0, 0.5]], "dihedral": 10.0}}),  # :338-363
    "step_change_in_twist_wing": dict(solver="nonlinear", wings={"main_wing": {
        "chord": [[0.0, 1.0], [1.0, 0.5]], "dihedral": 10.0, "twist": [[0.0, 5.0], [0.5, 5.0], [0.5, 0.0], [1.0, 0.0]],
        "sweep": 30.0, "grid": {"N": 100}}}),                                                                            # :366-397
    # solver-flag branches no reference test toggles (SURVEY.md section 8c "parity unpinned" item 6): pinned here by the reference itself
    "unswept_sections": dict(solver="nonlinear", wings={"main_wing": {"sweep": 25.0}}, solver_opts={"use_swept_sections": False}),
    "not_in_plane": dict(solver="nonlinear", wings={"main_wing": {"sweep": 25.0}}, solver_opts={"use_in_plane": False}),
    "match_machup_pro": dict(solver="nonlinear", state=dict(BASE_STATE, angular_rates=[0.1, 0.05, 0.02]),
                             solver_opts={"match_machup_pro": True}),
    "kuchemann_offset": dict(solver="nonlinear", wings={"main_wing": {"sweep": 30.0, "ll_offset": "kuchemann"}}),
    "previous_guess": dict(solver="nonlinear", state=BASE_STATE, controls=NO_CONTROLS),
    # scene.py:595-611: trailing directions projected into the aircraft's x-y plane (banked, pitched, yawed and rotating aircraft so
    # that the body z axis is nowhere near an earth axis and the joints move with different velocities; tail raised out of the
    # wing plane, otherwise the confined sheet grazes the tail's control points and the case is near-impingement)
    "constrain_vortex_sheet": dict(solver="nonlinear", state=dict(BASE_STATE, alpha=4.0, beta=3.0, angular_rates=[0.1, 0.05, 0.02],
                                                                  orientation=[20.0, 8.0, 30.0]),
                                   wings={"h_stab": {"connect_to": {"ID": 1, "location": "root", "dx": -3.0, "dz": -0.5}},
                                          "v_stab": {"connect_to": {"ID": 1, "location": "root", "dx": -3.0, "dz": -0.6}}},
                                   solver_opts={"constrain_vortex_sheet": True}),
}


def solve_case(Scene, base, name):
    spec = SOLVE_CASES[name]
    scene = Scene(_with(base, **spec))
    if name == "previous_guess":
        scene.solve_forces()
        scene.set_aircraft_state(dict(BASE_STATE, alpha=3.0))
        FM = scene.solve_forces(initial_guess="previous", report_by_segment=True, stab_frame=True)
    else:
        FM = scene.solve_forces(report_by_segment=True, stab_frame=True)
    return {"FM": FM, "gamma": [float(x) for x in np.asarray(scene._gamma)]}


# ---- test/scene_tests/test_derivatives.py ---------------------------------------------------------------------
def derivatives_case(Scene, base, which, **kw):
    """``kw``: extra keyword arguments of this package's drivers (batched=True); the reference is called without any."""
    scene = Scene(copy.deepcopy(base))
    if which == "stability":
        return scene.stability_derivatives(**kw)             # :9-33
    if which == "damping":
        return scene.damping_derivatives(**kw)               # :36-69
    if which == "control":
        return scene.control_derivatives(**kw)               # :72-106
    if which == "all":
        return scene.derivatives(stab_frame=True, **kw)      # :109-160 (plus the stability frame)
    if which == "state":
        return scene.state_derivatives(**kw)
    raise KeyError(which)


# ---- test/scene_tests/test_other_methods.py ----------------------------------------------------------------------
def other_case(Scene, base, which, **kw):
    scene = Scene(_with(base, state=BASE_STATE, controls=NO_CONTROLS))
    if which == "target_CL":
        return {"alpha": scene.target_CL(CL=0.5, **kw)}      # :11-35
    if which == "pitch_trim":
        return scene.pitch_trim(**kw)                        # :38-63
    if which == "pitch_trim_finite_moment":
        return scene.pitch_trim(CL=0.1, Cm=0.2, **kw)        # :66-91
    if which == "pitch_trim_orientation":
        state, controls = scene.pitch_trim_using_orientation(CL=0.1, Cm=0.2, **kw)   # :94-122
        return {"state": {k: [float(x) for x in v] for k, v in state.items()}, "controls": controls}
    if which == "aero_center":
        return scene.aero_center(**kw)
    if which == "MAC":
        return scene.MAC()
    raise KeyError(which)


# ---- test/scene_tests/test_multiple_aircraft.py --------------------------------------------------------------------
def two_aircraft_case(Scene, base, separation):
    """Two copies of the test aeroplane flying side by side, yawed 90 degrees (:10-38 with separation 25, :41-87 with 10000)."""
    d = copy.deepcopy(base)
    spec = d["scene"]["aircraft"].pop("test_plane")
    d["solver"]["type"] = "nonlinear"
    scene = Scene(d)
    state = copy.deepcopy(spec["state"])
    state["orientation"] = [np.sqrt(2) / 2, 0, 0, np.sqrt(2) / 2]
    scene.add_aircraft("test_plane_0", copy.deepcopy(spec["file"]), state=state, control_state=spec["control_state"])
    state["position"] = [separation, 0, 0]
    scene.add_aircraft("test_plane_1", copy.deepcopy(spec["file"]), state=state, control_state=spec["control_state"])
    FM = scene.solve_forces()
    return {k: dict(v["total"]) for k, v in FM.items()}


# ---- test/scene_tests/test_distributions.py ---------------------------------------------------------------------------
def distributions_case(Scene, base):
    scene = Scene(copy.deepcopy(base))
    dist = scene.distributions()
    return {ac: {seg: {k: [float(x) for x in v] for k, v in cols.items()} for seg, cols in segs.items()} for ac, segs in dist.items()}


# ---- Scene.export_pylot_model (scene.py:3521-3747): 59 solves, no reference test; pinned by running the reference ------------------
def pylot_case(Scene, base, filename, **kw):
    """The linearised model as the file the call writes (the reference returns nothing)."""
    import json
    scene = Scene(_with(base, state=BASE_STATE, controls=NO_CONTROLS))
    scene.export_pylot_model(filename=filename, set_accel_derivs=True, stall_angle_of_attack=15.0, **kw)
    with open(filename) as fh:
        return json.load(fh)


# ---- BASELINE.json configs[2]: alpha / beta sweep of examples/FlyingWing (SURVEY.md section 8d "C3") ---------------------------------
C3_ALPHAS = np.linspace(-2.0, 10.0, 64)
C3_BETAS = np.linspace(-6.0, 6.0, 64)
C3_SUB = (0, 9, 18, 27, 36, 45, 54, 63)          # the 8 x 8 sub-grid the reference is run on for the golden


def c3_states(alpha_idx=None, beta_idx=None):
    ai = range(64) if alpha_idx is None else alpha_idx
    bi = range(64) if beta_idx is None else beta_idx
    return [{"velocity": 100.0, "alpha": float(C3_ALPHAS[a]), "beta": float(C3_BETAS[b])} for a in ai for b in bi]
