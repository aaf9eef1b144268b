"""How far the device V_ji is from the oracle's (= the reference's, to 1e-12) on every golden scene, for a given library build.
Run once with the shipped library and once with one built with -DINF_FAST=0 (IEEE divisions instead of the 3-instruction
reciprocal) to see which part of the distance is the reciprocal.  usage: python scripts/influence_accuracy.py [lib.so]"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from machupx_b200 import _native
if len(sys.argv) > 1:
    _native.LIB_PATH = os.path.abspath(sys.argv[1])
import torch
from golden_utils import CASES, load
from machupx_b200.device import DeviceScene
from oracle import lls_oracle as O      # checker only
rows = []
for name in CASES:
    tables, ac_state, g, meta = load(name)
    dev = DeviceScene(tables)
    dev.upload_state(ac_state)
    dev.flow_state()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dev.build_influence(1e-10)
    e0.record(); dev.build_influence(1e-10); e1.record(); torch.cuda.synchronize()
    V = dev.V_ji()
    T = O.Tables(tables, ac_state)
    Vo, amp = O.influence_with_conditioning(T, O.flow_state(T))
    scale = np.abs(Vo).max()
    err = np.abs(V - Vo).max(axis=-1)
    rows.append({"scene": name, "n": int(tables["n"]), "max_err_over_scale": float(err.max() / scale),
                 "median_err_over_scale": float(np.median(err) / scale),
                 "max_err_over_conditioning_bound": float((err / (1e-11 * scale + 4e-13 * amp)).max()),
                 "frac_entries_above_1e-13_scale": float((err > 1e-13 * scale).mean()), "ms": e0.elapsed_time(e1)})
print(json.dumps({"lib": os.path.basename(_native.LIB_PATH), "scenes": rows}))
