"""Smallest / largest pivot magnitude the LU's guard saw in the last factorisation of a solve (scalars[1], scalars[2])."""
import ctypes, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from machupx_b200 import Scene, _native, layout as L
from workloads import swarm_input
from golden_utils import load
import copy, warnings

def scalars_of(dev):
    off = ctypes.c_size_t(0)
    _native.check(dev.lib.lls_status_offset(dev.handle, ctypes.byref(off)))
    sc = dev.d_work[off.value - 256: off.value - 256 + 32].view(torch.float64).cpu().numpy()
    st = int(dev.d_work[off.value: off.value + 4].view(torch.int32).cpu().numpy()[0])
    return {"min_pivot": float(sc[1]), "max_pivot": float(sc[2]), "status": st}

cases = {"c2": lambda: copy.deepcopy(load("c2_traditional_n4000")[3]["input"]), "swarm800": lambda: copy.deepcopy(load("swarm_nonlinear")[3]["input"]),
         "lattice8": lambda: swarm_input(8, 200)}
for name, make in cases.items():
    scene = Scene(make())
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        scene.solve_forces()
        a = scalars_of(scene._device)
        gam = np.array(scene._gamma)
        dev = scene._device
        dM = dev.d_M[:scene._N, :scene._N].diagonal().abs()
        a["diag_U_min_max"] = [float(dM.min()), float(dM.max())]
        scene.solve_forces(partitioned=True)
        b = scalars_of(scene._part_device)
    print(json.dumps({"case": name, "n": scene._N, "fused": a, "partitioned": b, "warnings": [str(x.message)[:60] for x in w]}))
