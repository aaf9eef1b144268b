"""Generates tests/golden/* from the UNMODIFIED reference (run in the build container only).

    python oracle/gen_golden.py            # writes tests/golden/<case>.npz + <case>.json

The reference (/root/reference, read-only) is imported with oracle/ref_stubs standing in for
airfoil_db / matplotlib / numpy-stl (61 of the reference's 62 non-CLI
tests pass unchanged on those stubs).  For each case this script
  1. builds the reference Scene from the reference's own JSON inputs (paths resolved, aircraft
     and airfoil files embedded so the fixture is self-contained),
  2. flattens the reference objects with machupx_b200.tables.build_tables (duck-typed),
  3. runs the reference solve and stores its V_ji (selected rows + row/column checksums),
     linear and nonlinear circulation, residual history, per-control-point arrays and the
     force/moment dictionary,
  4. checks oracle/lls_oracle.py against all of it (the "pin").
Nothing under tests/ or the product reads /root/reference at run time.
"""
import copy
import io
import contextlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = "/root/reference"
sys.path.insert(0, REPO)
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(HERE, "ref_stubs"))

import machupX as MX  # noqa: E402  (the reference)
from machupx_b200.tables import build_tables, build_aircraft_state  # noqa: E402
from oracle import lls_oracle as O  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")


def _load_json(path):
    with open(path) as fh:
        return json.load(fh)


def _embed(scene_input, base_dir):
    """Resolve "file"/"airfoils" paths into embedded dicts."""
    d = copy.deepcopy(scene_input)
    for name, ac in d["scene"]["aircraft"].items():
        f = ac["file"]
        if isinstance(f, str):
            ac["file"] = _load_json(os.path.join(base_dir, f))
        af = ac["file"].get("airfoils", None)
        if isinstance(af, str):
            ac["file"]["airfoils"] = _load_json(os.path.join(base_dir, af))
    return d


def _cases():
    cases = {}
    # ---- the reference's unit-test aeroplane (test/input_for_testing.json), N = 200 --------------
    base = _embed(_load_json(os.path.join(REF, "test/input_for_testing.json")), REF)
    c = copy.deepcopy(base)
    c["scene"]["aircraft"]["test_plane"]["state"] = {"position": [0, 0, 1000], "angular_rates": [0.0, 0.0, 0.0],
                                                     "velocity": 100, "alpha": 2.0, "beta": 0.0}
    cases["testplane_nonlinear"] = c                        # test_solve_forces.py:87-122
    c = copy.deepcopy(base)
    c["scene"]["aircraft"]["test_plane"]["state"] = {"position": [0, 0, 1000], "angular_rates": [0.1, 0.05, 0.01],
                                                     "velocity": 100, "alpha": 2.0, "beta": 3.0}
    c["scene"]["aircraft"]["test_plane"]["control_state"] = {"elevator": 3.0, "rudder": -2.0, "aileron": 5.0}
    cases["testplane_rot_sideslip_flaps"] = c               # test_solve_forces.py:163-236 combined
    # swept, tapered, dihedral main wing (test_solve_forces.py:277-363 style modifications)
    c = copy.deepcopy(base)
    wing = c["scene"]["aircraft"]["test_plane"]["file"]["wings"]["main_wing"]
    wing["sweep"] = 30.0
    wing["dihedral"] = [[0.0, 0.0], [1.0, 8.0]]
    wing["chord"] = [[0.0, 1.0], [1.0, 0.5]]
    wing["twist"] = [[0.0, 2.0], [1.0, -2.0]]
    c["scene"]["aircraft"]["test_plane"]["control_state"] = {"elevator": -2.0, "rudder": 0.0, "aileron": 4.0}
    cases["testplane_swept_tapered"] = c
    # spanwise airfoil distribution (test_airfoil_interpolation.py style)
    c = copy.deepcopy(base)
    wing = c["scene"]["aircraft"]["test_plane"]["file"]["wings"]["main_wing"]
    wing["airfoil"] = [[0.0, "NACA_0010"], [0.5, "NACA_2410"], [1.0, "NACA_4410"]]
    cases["testplane_airfoil_dist"] = c
    # trailing vortex sheet constrained to the aircraft's x-y plane (scene.py:595-611) on a banked, pitched, yawed, rotating aircraft
    c = copy.deepcopy(base)
    c["scene"]["aircraft"]["test_plane"]["state"] = {"position": [0, 0, 1000], "angular_rates": [0.1, 0.05, 0.02], "velocity": 100,
                                                     "alpha": 4.0, "beta": 3.0, "orientation": [20.0, 8.0, 30.0]}
    c["solver"]["constrain_vortex_sheet"] = True
    # tail raised out of the wing's x-y plane: with the sheet confined to that plane the main wing's trailing legs would otherwise
    # graze the tail's control points, and the reference's own V_ji is then only reproducible to ~1e-7 (near-impingement)
    wings = c["scene"]["aircraft"]["test_plane"]["file"]["wings"]
    wings["h_stab"]["connect_to"]["dz"] = -0.5
    wings["v_stab"]["connect_to"]["dz"] = -0.6
    cases["testplane_constrained_sheet"] = c
    # ---- examples/Traditional (C1: linear, N = 200) ------------------------------------------------
    d = os.path.join(REF, "examples/Traditional")
    c = _embed(_load_json(os.path.join(d, "traditional_input.json")), d)
    c.pop("run", None)
    c["solver"]["type"] = "linear"
    cases["traditional_linear"] = c
    c = copy.deepcopy(c)
    c["solver"]["type"] = "nonlinear"
    cases["traditional_nonlinear"] = c
    # ---- examples/FlyingWing (C3 geometry, N = 180), one swept state --------------------------------
    d = os.path.join(REF, "examples/FlyingWing")
    c = _embed(_load_json(os.path.join(d, "flying_wing_input.json")), d)
    c.pop("run", None)
    c["solver"]["type"] = "nonlinear"
    ac = next(iter(c["scene"]["aircraft"].values()))
    ac["state"] = {"velocity": 100.0, "alpha": 5.0, "beta": 3.0}
    cases["flyingwing_nonlinear"] = c
    # ---- examples/Swarm (C4 geometry at its shipped grid, N = 800, 4 aircraft), alpha ~ 5 deg --------
    d = os.path.join(REF, "examples/Swarm")
    c = _embed(_load_json(os.path.join(d, "swarm_input.json")), d)
    c.pop("run", None)
    c["solver"]["type"] = "nonlinear"
    for ac in c["scene"]["aircraft"].values():
        ac["state"]["velocity"] = [39.85, 0.0, 3.486, "kn"]
    cases["swarm_nonlinear"] = c
    return cases


def _big_cases():
    """Full-size configs of BASELINE.json (tables + reference results only; no V_ji rows).  Slow: minutes."""
    cases = {}
    d = os.path.join(REF, "examples/Traditional")
    c = _embed(_load_json(os.path.join(d, "traditional_input.json")), d)
    c.pop("run", None)
    c["solver"]["type"] = "nonlinear"
    for w in next(iter(c["scene"]["aircraft"].values()))["file"]["wings"].values():
        w.setdefault("grid", {})["N"] = 800
    cases["c2_traditional_n4000"] = c     # BASELINE.json configs[1]
    return cases


def _solver_params(inp):
    s = inp.get("solver", {})
    return {"type": s.get("type", "nonlinear"), "convergence": s.get("convergence", 1e-10),
            "relaxation": s.get("relaxation", 1.0), "max_iterations": s.get("max_iterations", 100)}


def _run_reference(inp):
    scene = MX.Scene(copy.deepcopy(inp))
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        FM = scene.solve_forces(verbose=True, report_by_segment=True, stab_frame=True)
    hist = []
    for line in buf.getvalue().splitlines():
        parts = line.split()
        if len(parts) == 2 and parts[0].isdigit():
            hist.append(float(parts[1]))
    return scene, FM, hist


def generate(name, inp, check=True, lean=False):
    scene, FM, hist = _run_reference(inp)
    tables = build_tables(scene)
    ac_state = build_aircraft_state(scene)
    n = tables["n"]
    sp = _solver_params(inp)

    # snapshot what the nonlinear solve left behind before anything else touches the scene
    V_ref = scene._V_ji
    gamma = scene._gamma.copy()
    snap = {k: np.array(getattr(scene, "_" + k)) for k in
            ("v_i", "alpha", "CL", "CD", "Cm", "Re", "M", "dF_inv", "dM_inv", "dF_visc", "dM_visc", "u_trailing_0", "u_trailing_1")}
    n_pre = tables["n"]
    rows_pre = sorted(set([0, 1, n_pre // 4, n_pre // 2 - 1, n_pre // 2, n_pre - 2, n_pre - 1]))
    V_rows = V_ref[rows_pre].copy()
    V_stats = {"V_rowsum": V_ref.sum(axis=1), "V_colsum": V_ref.sum(axis=0), "V_abs_sum": np.abs(V_ref).sum(),
               "V_diag": V_ref[np.arange(n_pre), np.arange(n_pre)].copy()}
    # the reference's linear solution
    if lean:
        scene._solver_type = "linear"
        scene.solve_forces()
        gamma_lin = scene._gamma.copy()
    else:
        scene_lin = MX.Scene(copy.deepcopy(inp))
        scene_lin._solver_type = "linear"
        scene_lin.solve_forces()
        gamma_lin = scene_lin._gamma.copy()

    rows = sorted(set([0, 1, n // 4, n // 2 - 1, n // 2, n - 2, n - 1]))
    gold = {
        "tab": tables["tab"], "itab": tables["itab"], "airfoils": tables["airfoils"], "ac_state": ac_state,
        "n": n, "ld": tables["ld"], "n_aircraft": tables["n_aircraft"], "n_segments": tables["n_segments"],
        "flags": tables["flags"],
        "V_rows_idx": np.array(rows), "V_rows": V_rows,
        "gamma_linear": gamma_lin, "gamma": gamma, "error_history": np.array(hist),
    }
    gold.update(V_stats)
    gold.update(snap)
    if name == "testplane_nonlinear":
        # the reference ships a distributions() dump of this aeroplane (test/input_for_testing.txt): keep its
        # circulation, lift-coefficient and angle-of-attack columns as an independent golden vector
        d = np.genfromtxt(os.path.join(REF, "test/input_for_testing.txt"), skip_header=1, dtype=None, encoding="utf-8")
        gold["circ_file"] = np.array([row[31] for row in d], dtype=float)
        gold["cl_file"] = np.array([row[21] for row in d], dtype=float)
        gold["alpha_file"] = np.array([row[13] for row in d], dtype=float)
    if lean:
        for k in ("V_colsum", "dM_inv", "dM_visc", "dF_visc", "u_trailing_0", "u_trailing_1"):
            gold.pop(k)
        gold["V_rows_idx"] = np.array(rows[:3])
        gold["V_rows"] = V_rows[:3]
    meta = {"input": inp, "solver": sp, "FM": FM, "iterations": len(hist),
            "segment_names": [list(x) for x in tables["segment_names"]],
            "reference": "usuaero/MachUpX v2.7.2 + oracle/ref_stubs", "n": n,
            "aircraft": [{"name": ap.name, "q": list(map(float, ap.q)), "v": list(map(float, ap.v)),
                          "p_bar": list(map(float, ap.p_bar)), "S_w": float(ap.S_w), "l_ref_lat": float(ap.l_ref_lat),
                          "l_ref_lon": float(ap.l_ref_lon), "wind": list(map(float, scene._get_wind(ap.p_bar))),
                          "rho": float(scene._get_density(ap.p_bar)), "segments": [sg.name for sg in ap.segments]}
                         for ap in scene._airplane_objects]}

    if check and not lean:
        T = O.Tables(tables, ac_state)
        res = O.solve(T, solver_type=sp["type"], convergence=sp["convergence"], relaxation=sp["relaxation"],
                      max_iterations=sp["max_iterations"])
        scale = np.abs(V_ref).max()
        eV = np.abs(res["V"] - V_ref).max() / scale
        eG = np.abs(res["gamma"] - gamma).max() / np.abs(gamma).max()
        eGl = np.abs(res["gamma_linear"] - gamma_lin).max() / np.abs(gamma_lin).max()
        print("  [pin] %-30s N=%4d  V_ji %.2e  gamma_lin %.2e  gamma %.2e  iters ref %d oracle %d"
              % (name, n, eV, eGl, eG, len(hist), res["iterations"]))
        assert eV < 1e-12 and eG < 1e-11 and eGl < 1e-11, "oracle does not reproduce the reference"
        assert res["iterations"] == len(hist)

    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **gold)
    with open(os.path.join(OUT, name + ".json"), "w") as fh:
        json.dump(meta, fh, indent=1)
    return scene


if __name__ == "__main__":
    os.chdir(REF)  # the reference resolves relative "file" paths against the CWD
    only = sys.argv[1:]
    for cname, cinp in _big_cases().items():
        if cname in only:
            print("generating", cname)
            generate(cname, cinp, lean=True)
    for cname, cinp in _cases().items():
        if only and cname not in only:
            continue
        print("generating", cname)
        generate(cname, cinp)
