"""Builds profiles/r2_scaling_swarm.json from the RAW bench lines of the multi-GPU runs (profiles/r2_raw/*_bench_n*.json, copied
unchanged from gpurun_out/): solves/s, speed-up and efficiency against the one-GPU line of the same scene, LU TFLOP/s per GPU,
exposed exchange wait, and the force sums that must agree between the world sizes.  Nothing here is typed in by hand."""
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RAW = os.path.join(ROOT, "profiles", "r2_raw")
pick = sys.argv[1:] or ["scale2_bench_n1.json", "scale2b_bench_n2.json", "scale8c_bench_n4.json", "scale8c_bench_n8.json"]
rows = []
for name in pick:
    path = os.path.join(RAW, name)
    if not os.path.exists(path):
        continue
    d = json.loads(open(path).read().strip().splitlines()[-1])
    waits = d["exchange"]["per_rank_wait_ms_per_step"]["panel_wait, slot_wait"]
    rows.append({"file": "profiles/r2_raw/" + name, "n_gpus": d["n_gpus"], "control_points": d["config"]["control_points"],
                 "solves_per_s": d["value"], "ms_per_solve": d["ms_per_step"], "e2e_solves_per_s": d["e2e"]["value"],
                 "lu_ms_per_solve": d["phase_ms"]["lu"], "lu_tflops_per_gpu": d["roofline"]["achieved"], "dmma_peak_tflops": d["roofline"]["peak"],
                 "newton_iterations": d["per_rank_newton_iterations"], "max_panel_wait_ms_per_solve": max(w[0] for w in waits) if waits else None,
                 "exchange": d["exchange"]["path"], "multicast": d["exchange"].get("multicast"), "CL_sum": d["parity"]["CL_sum"], "final_residual": d["parity"]["final_residual"],
                 "sm_mhz": (d.get("clocks") or {}).get("sm_mhz"), "throttle_reasons": (d.get("clocks") or {}).get("reasons")})
base = next((r for r in rows if r["n_gpus"] == 1), None)
for r in rows:
    if base:
        r["speedup_vs_1_gpu"] = r["solves_per_s"] / base["solves_per_s"]
        r["efficiency"] = r["speedup_vs_1_gpu"] / r["n_gpus"]
out = {"workload": "synthetic 32-aircraft swarm, 32,000 control points, nonlinear (bench.py --workload swarm); strong scaling",
       "note": "N = 1: fused single-GPU path; N > 1: rows partitioned, LU panels exchanged inside the kernels over NVLink peer memory",
       "rows": rows}
with open(os.path.join(ROOT, "profiles", "r2_scaling_swarm.json"), "w") as fh:
    json.dump(out, fh, indent=1)
for r in rows:
    print(r["n_gpus"], "%.4f solves/s" % r["solves_per_s"], "x%.2f" % r.get("speedup_vs_1_gpu", 0), "eff %.2f" % r.get("efficiency", 0),
          "%.1f TF/GPU" % r["lu_tflops_per_gpu"], r["CL_sum"])
