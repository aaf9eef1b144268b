"""BASELINE.json configs[2]: FlyingWing alpha/beta sweep, 64 x 64 = 4,096 flight states, batched on one GPU.
usage: python scripts/c3_sweep_bench.py [n_alpha n_beta] [linear|nonlinear]"""
import copy, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from machupx_b200 import Scene
na = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 64
solver = sys.argv[3] if len(sys.argv) > 3 else "nonlinear"
inp = json.load(open(os.path.join(ROOT, "tests/golden/flyingwing_nonlinear.json")))["input"]
inp = copy.deepcopy(inp); inp["solver"]["type"] = solver
scene = Scene(inp)
name = list(scene._airplanes.keys())[0]
pos = [float(x) for x in scene._airplanes[name].p_bar]
states = [{"position": pos, "velocity": 100.0, "alpha": float(a), "beta": float(b)}
          for a in np.linspace(-2.0, 10.0, na) for b in np.linspace(-6.0, 6.0, nb)]
scene.set_err_state(not_converged="ignore")
res = scene.solve_forces_batch(states)       # warm-up (allocations)
torch.cuda.synchronize()
t0 = time.perf_counter(); res = scene.solve_forces_batch(states); torch.cuda.synchronize(); t_e2e = time.perf_counter() - t0
dev = scene._batch_device
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
dev.solve(scene._solver_type, scene._solver_convergence, scene._solver_relaxation, scene._max_solver_iterations)
e1.record(); torch.cuda.synchronize()
t_dev = e0.elapsed_time(e1) * 1e-3
# sequential single-case path on a sample
k = min(32, len(states)); idx = np.linspace(0, len(states) - 1, k).astype(int)
t0 = time.perf_counter()
seq = []
for i in idx:
    scene.set_aircraft_state(states[i]); seq.append(scene.solve_forces()[name]["total"]["CL"])
torch.cuda.synchronize(); t_seq = (time.perf_counter() - t0) / k
err = max(abs(res["totals"][name]["CL"][i] - s) for i, s in zip(idx, seq))
print(json.dumps({"workload": "FlyingWing N=%d, %d x %d alpha/beta grid, %s" % (scene._N, na, nb, solver), "cases": len(states),
                  "batched_cases_per_s_device": len(states) / t_dev, "batched_cases_per_s_e2e": len(states) / t_e2e,
                  "sequential_cases_per_s": 1.0 / t_seq, "iterations_min_max": [int(res["iterations"].min()), int(res["iterations"].max())],
                  "converged": int(res["converged"].sum()), "max_abs_CL_diff_batch_vs_sequential": err,
                  "device_seconds": t_dev, "e2e_seconds": t_e2e}))
