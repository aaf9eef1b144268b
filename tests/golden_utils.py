"""Loading of the golden vectors generated from the unmodified reference by oracle/gen_golden.py."""
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CASES = ["testplane_nonlinear", "testplane_rot_sideslip_flaps", "testplane_swept_tapered", "testplane_airfoil_dist",
         "testplane_constrained_sheet",
         "traditional_linear", "traditional_nonlinear", "flyingwing_nonlinear", "swarm_nonlinear"]


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    with open(os.path.join(GOLDEN_DIR, name + ".json")) as fh:
        meta = json.load(fh)
    g = {k: z[k] for k in z.files}
    tables = {"tab": g["tab"], "itab": g["itab"], "airfoils": g["airfoils"], "n": int(g["n"]), "ld": int(g["ld"]),
              "n_aircraft": int(g["n_aircraft"]), "n_segments": int(g["n_segments"]), "flags": int(g["flags"])}
    from machupx_b200.layout import pad_ac_state
    return tables, pad_ac_state(g["ac_state"]), g, meta


def segment_sums(g, tables):
    """Per-segment sums of the reference's per-control-point force/moment elements (scene.py:1108-1117)."""
    from machupx_b200 import layout as L
    n = tables["n"]
    seg = tables["itab"][L.I_SEGMENT, :n]
    out = np.zeros((tables["n_segments"], 12))
    for s in range(tables["n_segments"]):
        m = seg == s
        out[s, 0:3] = g["dF_inv"][m].sum(axis=0)
        out[s, 3:6] = g["dM_inv"][m].sum(axis=0)
        out[s, 6:9] = g["dF_visc"][m].sum(axis=0)
        out[s, 9:12] = g["dM_visc"][m].sum(axis=0)
    return out
