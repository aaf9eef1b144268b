"""The CPU oracle against the golden vectors produced by the unmodified reference (oracle/gen_golden.py),
and against known answers hard-coded in the reference's own test-suite."""
import numpy as np
import pytest

from golden_utils import CASES, load, segment_sums
from oracle import lls_oracle as O


def _solve(name):
    tables, ac_state, g, meta = load(name)
    T = O.Tables(tables, ac_state)
    sp = meta["solver"]
    res = O.solve(T, solver_type=sp["type"], convergence=sp["convergence"], relaxation=sp["relaxation"],
                  max_iterations=sp["max_iterations"])
    return tables, g, meta, res


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference(name):
    tables, g, meta, res = _solve(name)
    n = tables["n"]
    scale = np.abs(g["V_rows"]).max()
    assert np.abs(res["V"][g["V_rows_idx"]] - g["V_rows"]).max() / scale < 1e-12
    assert np.abs(res["V"].sum(axis=1) - g["V_rowsum"]).max() / np.abs(g["V_rowsum"]).max() < 1e-11
    assert np.abs(res["V"].sum(axis=0) - g["V_colsum"]).max() / np.abs(g["V_colsum"]).max() < 1e-11
    assert np.abs(res["V"][np.arange(n), np.arange(n)] - g["V_diag"]).max() / scale < 1e-12
    assert np.abs(res["gamma_linear"] - g["gamma_linear"]).max() / np.abs(g["gamma_linear"]).max() < 1e-11
    assert np.abs(res["gamma"] - g["gamma"]).max() / np.abs(g["gamma"]).max() < 1e-11
    assert res["iterations"] == meta["iterations"]
    np.testing.assert_allclose(res["history"], g["error_history"], rtol=1e-6, atol=5e-12)  # late residuals are rounding noise
    ref = segment_sums(g, tables)
    assert np.abs(res["seg_fm"] - ref).max() / np.abs(ref).max() < 1e-11
    for key in ("alpha", "CD", "Cm", "Re", "M"):
        np.testing.assert_allclose(res["dist"][key], g[key], rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(res["flow"].u_trailing_0, g["u_trailing_0"], rtol=0, atol=1e-14)


def _fm_from(seg_fm, meta, **kw):
    from machupx_b200.forces import AircraftFrame, assemble_fm
    frames = [AircraftFrame(a["name"], a["q"], a["v"], a["wind"], a["rho"], a["S_w"], a["l_ref_lat"], a["l_ref_lon"], a["segments"])
              for a in meta["aircraft"]]
    return assemble_fm(seg_fm, frames, **kw)


def _compare_fm(mine, ref, tol=1e-10):
    assert set(mine) == set(ref)
    for ac in ref:
        for part in ("inviscid", "viscous"):
            assert set(mine[ac][part]) == set(ref[ac][part]), (part, set(mine[ac][part]) ^ set(ref[ac][part]))
            for key, d in ref[ac][part].items():
                assert set(mine[ac][part][key]) == set(d)
                for where, val in d.items():
                    assert abs(mine[ac][part][key][where] - val) <= tol * max(1.0, abs(val)), (ac, part, key, where)
        assert set(mine[ac]["total"]) == set(ref[ac]["total"])
        for key, val in ref[ac]["total"].items():
            assert abs(mine[ac]["total"][key] - val) <= tol * max(1.0, abs(val)), (ac, key)


@pytest.mark.parametrize("name", ["testplane_nonlinear", "testplane_rot_sideslip_flaps", "flyingwing_nonlinear", "swarm_nonlinear"])
def test_force_dictionary_layout_and_values(name):
    """Host FM assembly (machupx_b200/forces.py) fed with the oracle's segment sums == the reference's FM dict."""
    tables, g, meta, res = _solve(name)
    mine = _fm_from(res["seg_fm"], meta, report_by_segment=True, stab_frame=True)
    _compare_fm(mine, meta["FM"])


def test_reference_known_answers_nonlinear():
    """Hard-coded goldens of the reference's test_nonlinear_solver (test/scene_tests/test_solve_forces.py:114-122)."""
    tables, g, meta, res = _solve("testplane_nonlinear")
    total = _fm_from(res["seg_fm"], meta)["test_plane"]["total"]
    # the constants the reference asserts (to 1e-10) in test/scene_tests/test_solve_forces.py:114-122
    gold = {"FL": 20.959074770349314, "FD": 1.3934114313254868, "FS": 0.0, "Fx": -0.661101441894963,
            "Fy": 0.0, "Fz": -20.994936425947238, "Mx": 0.0, "My": -12.873864820066961, "Mz": 0.0}
    ref_total = meta["FM"]["test_plane"]["total"]
    for k, v in gold.items():
        assert abs(total[k] - v) < 1e-10, (k, total[k], v)           # oracle vs the reference's hard-coded answer
        assert abs(total[k] - ref_total[k]) < 1e-10, k                # oracle vs the fixture generated from the reference


def test_circulation_golden_file_of_the_reference():
    """Gamma of the unit-test aeroplane against the Circ column of the reference's test/input_for_testing.txt
    (a distributions() dump shipped with the reference; 12 printed digits), stored in the golden fixture."""
    tables, g, meta, res = _solve("testplane_nonlinear")
    if "circ_file" not in g:
        pytest.skip("fixture without the Circ column")
    assert np.abs(res["gamma"] - g["circ_file"]).max() < 1e-11 * np.abs(g["circ_file"]).max()   # 12 printed digits
    assert np.abs(np.degrees(res["dist"]["alpha"]) - g["alpha_file"]).max() < 1e-10 or \
        np.abs(res["dist"]["alpha"] - g["alpha_file"]).max() < 1e-11


def test_c3_sub_grid_cases_of_the_reference():
    """The oracle on three corners of the reference's C3 sub-grid (tests/golden/api_scenarios.json "c3"): CL / CD / Cm and the
    Newton iteration count of the reference's serial sweep loop."""
    import copy
    import json
    import os
    import api_scenarios as S
    from machupx_b200 import Scene
    from machupx_b200.forces import assemble_fm
    from machupx_b200.tables import build_aircraft_state
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "api_scenarios.json")) as fh:
        gold = json.load(fh)["c3"]
    inp = copy.deepcopy(gold["input"])
    inp["solver"]["type"] = "nonlinear"
    inp["solver"].pop("relaxation", None)
    scene = Scene(inp)
    name = gold["nonlinear"]["aircraft"]
    states = S.c3_states(gold["alpha_idx"], gold["beta_idx"])
    for case in (0, 27, 63):
        scene.set_aircraft_state(states[case], aircraft=name)
        scene._refresh_tables()
        T = O.Tables(scene._tables, build_aircraft_state(scene))
        res = O.solve(T, solver_type="nonlinear", convergence=scene._solver_convergence, relaxation=scene._solver_relaxation,
                      max_iterations=scene._max_solver_iterations)
        assert res["iterations"] == gold["nonlinear"]["iterations"][case]
        tot = assemble_fm(res["seg_fm"], scene._frames())[name]["total"]
        for key, val in gold["nonlinear"]["totals"][case].items():
            assert abs(tot[key] - val) <= 1e-10 * max(1.0, abs(val)), (case, key)
