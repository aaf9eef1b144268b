"""Times lls_build_influence on the C2 tables for a given library build.  usage: python scripts/influence_bench.py <lib.so>"""
import ctypes, os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from machupx_b200 import _native
if len(sys.argv) > 1:
    _native.LIB_PATH = os.path.abspath(sys.argv[1])
from machupx_b200.device import DeviceScene
from golden_utils import load
tables, ac_state, g, meta = load("c2_traditional_n4000")
dev = DeviceScene(tables)
dev.upload_state(ac_state)
dev.flow_state()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for r in range(8):
    e0.record(); dev.build_influence(1e-10); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
dev.solve_linear(); it, err, status, hist = dev.solve_nonlinear(1e-10, 0.9, 100)
gam = dev.gamma()
print(json.dumps({"lib": os.path.basename(_native.LIB_PATH), "influence_ms": best, "Gpair_s": 16e6 / best * 1e-6,
                  "gamma_rel_err": float(np.abs(gam - g["gamma"]).max() / np.abs(g["gamma"]).max()), "iters": it}))
