"""Synthetic large scenes of BASELINE.json configs[3]-[4], shared by bench.py, the GPU tests and the
scripts.  Built from the committed fixture of examples/Swarm (tests/golden/swarm_nonlinear.json holds the reference's own JSON
input with the aircraft file embedded); nothing here reads /root/reference."""
import copy
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def swarm_base():
    with open(os.path.join(HERE, "golden", "swarm_nonlinear.json")) as fh:
        return json.load(fh)["input"]


def swarm_input(n_aircraft, grid_n):
    """n_aircraft = 4: examples/Swarm as shipped (4 drones) with ``grid.N = grid_n`` on every wing (5 segments per drone ->
    4 x 5 x grid_n control points; grid_n = 1000 is configs[3]).  Otherwise: the lattice of SURVEY.md section 8d "C5" with 8 columns
    and n_aircraft / 8 rows, position = [-10 r, 10 (c - 3.5), -50 - 5 ((r + c) mod 2) + eps], eps ~ U(-0.5, 0.5) from
    numpy.random.default_rng(0) (keeps trailing legs off downstream control points); 64 aircraft at grid_n = 200 is configs[4].
    Flight state of every drone: velocity = [39.85, 0, 3.486] kn (alpha ~ 5 deg), nonlinear solver."""
    base = swarm_base()
    inp = copy.deepcopy(base)
    drone = copy.deepcopy(next(iter(base["scene"]["aircraft"].values())))
    for w in drone["file"]["wings"].values():
        w.setdefault("grid", {})["N"] = int(grid_n)
    if n_aircraft == 4:
        for ac in inp["scene"]["aircraft"].values():
            ac["file"] = copy.deepcopy(drone["file"])
        return inp
    assert n_aircraft % 8 == 0, "lattice scenes have 8 columns"
    rng = np.random.default_rng(0)
    inp["scene"]["aircraft"] = {}
    for r in range(n_aircraft // 8):
        for c in range(8):
            ac = copy.deepcopy(drone)
            eps = float(rng.uniform(-0.5, 0.5))
            ac["state"]["position"] = [-10.0 * r, 10.0 * (c - 3.5), -50.0 - 5.0 * ((r + c) % 2) + eps]
            inp["scene"]["aircraft"]["drone_%d_%d" % (r, c)] = ac
    return inp
