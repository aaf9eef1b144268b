#!/usr/bin/env python
"""Runs tests/api_scenarios.py through the UNMODIFIED reference (machupX from /root/reference with the stand-ins in
oracle/ref_stubs for airfoil_db / matplotlib / numpy-stl) and writes tests/golden/api_scenarios.json plus the embedded
base input tests/golden/input_for_testing_embedded.json.  TEST INFRASTRUCTURE; run in the build container only
(/root/reference does not exist on the GPU box):

    python oracle/gen_api_golden.py

A few of the reference's own hard-coded test constants are asserted here, so the scenarios provably restate its tests.
"""
import json
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
sys.path.insert(0, os.path.join(HERE, "ref_stubs"))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import machupX as MX  # noqa: E402  (the reference)
import api_scenarios as S  # noqa: E402
from oracle.gen_golden import _embed, _load_json  # noqa: E402


def jsonable(x):
    import numpy as np
    if isinstance(x, dict):
        return {k: jsonable(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [jsonable(v) for v in x]
    if isinstance(x, np.ndarray):
        return [jsonable(v) for v in x.tolist()]
    if isinstance(x, (np.floating, np.integer)):
        return x.item()
    return x


def c3_golden():
    """BASELINE.json configs[2] on the 8 x 8 sub-grid S.C3_SUB of the 64 x 64 alpha / beta grid: the reference's serial
    set_aircraft_state -> solve_forces loop (scene.py:1535-1598) on examples/FlyingWing, linear (as shipped) and nonlinear
    (default relaxation), with the Newton iteration count of every case read from the verbose printout."""
    import contextlib
    import io
    import numpy as np
    d = os.path.join(REF, "examples/FlyingWing")
    inp = _embed(_load_json(os.path.join(d, "flying_wing_input.json")), d)
    inp.pop("run", None)
    res = {"input": inp, "alpha_idx": list(S.C3_SUB), "beta_idx": list(S.C3_SUB)}
    states = S.c3_states(S.C3_SUB, S.C3_SUB)
    for solver in ("linear", "nonlinear"):
        cfg = json.loads(json.dumps(inp))
        cfg["solver"]["type"] = solver
        cfg["solver"].pop("relaxation", None)
        scene = MX.Scene(cfg)
        name = list(scene._airplanes.keys())[0]
        totals, its = [], []
        for st in states:
            scene.set_aircraft_state(state=dict(st), aircraft=name)
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf):
                FM = scene.solve_forces(verbose=True)
            its.append(sum(1 for line in buf.getvalue().splitlines() if len(line.split()) == 2 and line.split()[0].isdigit()))
            totals.append(FM[name]["total"])
        res[solver] = {"totals": jsonable(totals), "iterations": its, "aircraft": name}
        print("c3", solver, "iterations", min(its), "-", max(its), "CL", totals[0]["CL"], "..", totals[-1]["CL"])
    return res


def main():
    warnings.simplefilter("ignore")
    os.chdir(REF)
    base = _embed(_load_json(os.path.join(REF, "test/input_for_testing.json")), REF)
    base.pop("run", None)
    out = {"solve": {}, "derivatives": {}, "other": {}, "two_aircraft": {}}
    for name in S.SOLVE_CASES:
        out["solve"][name] = jsonable(S.solve_case(MX.Scene, base, name))
        print("solve", name, out["solve"][name]["FM"]["test_plane"]["total"].get("FL"))
    # the reference's own constants (test_solve_forces.py:38, 114, 293; test_other_methods.py:35)
    assert abs(out["solve"]["linear_solver"]["FM"]["test_plane"]["total"]["FL"] - 20.961790953642968) < 1e-10
    assert abs(out["solve"]["nonlinear_solver"]["FM"]["test_plane"]["total"]["FL"] - 20.959074770349314) < 1e-10
    assert abs(out["solve"]["swept_wing"]["FM"]["test_plane"]["total"]["My"] + 30.13434354434114) < 1e-10
    for which in ("stability", "damping", "control", "all", "state"):
        out["derivatives"][which] = jsonable(S.derivatives_case(MX.Scene, base, which))
        print("derivatives", which)
    assert abs(out["derivatives"]["stability"]["test_plane"]["CL,a"] - 6.322012906729068) < 1e-10   # test_derivatives.py:16
    for which in ("target_CL", "pitch_trim", "pitch_trim_finite_moment", "pitch_trim_orientation", "aero_center", "MAC"):
        out["other"][which] = jsonable(S.other_case(MX.Scene, base, which))
        print("other", which, out["other"][which])
    assert abs(out["other"]["target_CL"]["alpha"] - 4.531085683208453) < 1e-10                      # test_other_methods.py:35
    assert abs(out["other"]["pitch_trim"]["test_plane"]["alpha"] - 12.27375383887546) < 1e-10       # test_other_methods.py:62
    for sep in (25, 10000):
        out["two_aircraft"][str(sep)] = jsonable(S.two_aircraft_case(MX.Scene, base, sep))
    out["distributions"] = jsonable(S.distributions_case(MX.Scene, base))
    out["pylot"] = jsonable(S.pylot_case(MX.Scene, base, "/tmp/_pylot_ref.json"))
    print("pylot", out["pylot"]["coefficients"]["CD0"])
    out["c3"] = c3_golden()
    os.chdir(ROOT)
    with open(os.path.join(ROOT, "tests/golden/input_for_testing_embedded.json"), "w") as fh:
        json.dump(base, fh, indent=1)
    with open(os.path.join(ROOT, "tests/golden/api_scenarios.json"), "w") as fh:
        json.dump(out, fh)
    print("wrote tests/golden/api_scenarios.json")


if __name__ == "__main__":
    main()
