#!/usr/bin/env python
"""Benchmark of the lifting-line solve path (BASELINE.json metric: lifting-line solves/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|testplane]

One "step" = one steady-state ``solve_forces()`` on an already-built scene: flow state, V_ji influence
build, linear solve, Newton iteration to 1e-10, force integration (reference: scene.py:1411-1532).
Workload at N=1: BASELINE.json configs[1] — Traditional aeroplane refined to 4,000 control points,
nonlinear solve.  At N>1 every rank solves the same scene at its own angle of attack (an alpha sweep
sharded by case): weak scaling, no data-path collective.

Prints ONE JSON line (rank 0).  `value` times the solve with the flight state already in HBM;
`e2e` times the public host-side call including the pinned host->device copy of the flight state and
the device->host read of the force sums, every step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# dram__bytes_read.sum + dram__bytes_write.sum of one lu_step_kernel launch (block step k = 8 of an N = 4000 factorisation,
# 54 x 54 = 2916 tiles) from the `ncu --set full` capture summarised in profiles/r1_lu_step_full.txt.  Algorithmic bytes
# of that launch: read + write of the trailing matrix, 2 x 8 x (54 x 64)^2 = 191 MB; the 126 MB L2 keeps part of it.
NCU_LU_TRAFFIC = {"bytes_per_launch": 138.1e6, "algorithmic_bytes_per_launch": 191.1e6,
                  "launch": "lu_step_kernel, block step 8 of 63 (2916 tiles) at N=4000", "source": "profiles/r1_lu_step_full.txt"}
METRIC = "lifting-line solves/sec"
UNIT = "solves/s"


# ------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------
def load_workload(name):
    """Returns (tables, ac_state, golden dict, meta).  Tables come from the committed fixtures generated
    from the reference's own JSON inputs (oracle/gen_golden.py); data are synthetic in the sense of
    BASELINE.json: a refined example scene, no measured flight data."""
    from golden_utils import load
    fixture = {"c2": "c2_traditional_n4000", "testplane": "testplane_nonlinear", "swarm": "swarm_nonlinear"}[name]
    return load(fixture)


def with_alpha(ac_state, alpha_deg, meta):
    """Flight state of rank r: same speed, angle of attack alpha_deg (airplane.py:300-316 convention)."""
    from machupx_b200 import layout as L
    st = np.array(ac_state, dtype=float)
    v = -st[0, L.AC_VTRANS:L.AC_VTRANS + 3]
    V = np.linalg.norm(v)
    a = np.radians(alpha_deg)
    st[0, L.AC_VTRANS:L.AC_VTRANS + 3] = -np.array([V * np.cos(a), 0.0, V * np.sin(a)])
    return st


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.samples = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            parts = [p.strip() for p in s.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nme, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(np.max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle port (numpy restatement of the reference's path)
# ------------------------------------------------------------------------------------------------
def cpu_sample(tables, ac_state, meta, threads):
    """Times a bounded sample of one solve with the oracle port and scales it to solves/s.

    sample: V_ji rows for 256 of N control points (scaled by N/256) + one linear-system assembly and solve
    + one residual + one Jacobian assembly + one dense solve; a full solve is 1 build + 1 linear solve
    + iters x (residual + Jacobian + solve) + 2 residual-class matvecs (scene.py:853-913, 943)."""
    from oracle import lls_oracle as O
    T = O.Tables(tables, ac_state)
    n = T.n
    iters = int(meta.get("iterations", 11)) or 1
    t0 = time.perf_counter()
    F = O.flow_state(T)
    rows = min(n, 256)
    i0 = (n - rows) // 2
    O.influence_rows(T, F, i0, i0 + rows)
    t_build = (time.perf_counter() - t0) * n / rows
    # remaining phases need a full V: build it (untimed beyond the sample above)
    V = O.influence(T, F)
    t0 = time.perf_counter()
    A, b = O.linear_system(T, F, V)
    t_lin_asm = time.perf_counter() - t0
    t0 = time.perf_counter()
    gamma = np.linalg.solve(A, b)
    t_solve = time.perf_counter() - t0
    t0 = time.perf_counter()
    S = O.residual(T, F, V, gamma)
    t_res = time.perf_counter() - t0
    t0 = time.perf_counter()
    J = O.jacobian(T, F, V, gamma, S)
    t_jac = time.perf_counter() - t0
    del J, A
    nonlinear = meta["solver"]["type"] == "nonlinear"
    total = t_build + t_lin_asm + t_solve + 2 * t_res
    if nonlinear:
        total += iters * (t_res + t_jac + t_solve) + t_res
    parts = {"build_s": t_build, "linear_assembly_s": t_lin_asm, "dense_solve_s": t_solve, "residual_s": t_res,
             "jacobian_s": t_jac, "newton_iterations": iters if nonlinear else 0}
    return 1.0 / total, total, parts


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    tables, ac_state, g, meta = load_workload(args.workload)
    vals = []
    for _ in range(max(1, min(args.steps, 2))):
        v, total, parts = cpu_sample(tables, ac_state, meta, threads)
        vals.append(v)
    value = float(np.median(vals))
    sample = ("oracle port (numpy/LAPACK restatement of scene.py:557-1117), %d BLAS threads: V_ji rows for 256 of N control points "
              "scaled to N, + 1 linear assembly/solve, 1 residual, 1 Jacobian, 1 dense solve, combined as "
              "1 build + 1 linear + iters x (residual+Jacobian+solve) + 3 residuals" % threads)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.workload, tables, meta),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample, "parts": parts},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(name, tables, meta):
    desc = {"c2": "Traditional airplane refined to 4,000 control points (grid.N=800 per segment), nonlinear Newton solve, "
                  "relaxation 0.9, convergence 1e-10 (BASELINE.json configs[1])",
            "testplane": "reference unit-test aeroplane, 200 control points, nonlinear",
            "swarm": "examples/Swarm, 4 aircraft, 800 control points, nonlinear"}[name]
    return {"workload": desc, "control_points": int(tables["n"]), "solver": meta["solver"],
            "l2_policy": "working set (V_ji %.0f MB + J %.0f MB) exceeds the 126 MB L2; no explicit flush" % (
                tables["n"] * 3 * tables["ld"] * 8 / 1e6, tables["ld"] ** 2 * 8 / 1e6)}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from machupx_b200 import _native
    from machupx_b200.device import DeviceScene
    from machupx_b200 import layout as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import copy
    import warnings
    from machupx_b200 import Scene
    from machupx_b200.tables import build_aircraft_state
    tables, ac_state, g, meta = load_workload(args.workload)
    sp = meta["solver"]
    nonlinear = sp["type"] == "nonlinear"
    # the public API object: JSON input of the fixture -> Scene (host geometry, O(N) tables) -> DeviceScene
    scene = Scene(copy.deepcopy(meta["input"]))
    if world > 1:   # alpha sweep sharded by case: rank r flies at alpha0 + 0.25 deg * r (first aircraft)
        name = list(scene._airplanes.keys())[0]
        scene._airplanes[name].set_aerodynamic_state(alpha=2.0 + 0.25 * rank)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        scene.solve_forces()
    dev = scene._device
    lib = dev.lib
    tables = scene._tables
    ac_state = build_aircraft_state(scene)

    def solve_resident():
        dev.flow_state()
        dev.build_influence(1e-10)
        dev.solve_linear()
        it = 0
        if nonlinear:
            it, err, status, hist = dev.solve_nonlinear(sp["convergence"], sp["relaxation"], sp["max_iterations"])
        with torch.cuda.device(dev.device):
            _native.check(lib.lls_integrate(dev.handle, dev.d_seg_fm.data_ptr(), dev._stream()))
        return it

    def solve_e2e():
        # the call a user makes: Scene.solve_forces() -> pinned host->device copy of the flight state, all kernels,
        # device->host read of the residual norms and the per-segment force sums, FM dictionary on the host
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return scene.solve_forces()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    dev.upload_state(ac_state)
    for _ in range(max(3, args.warmup)):
        iters = solve_resident()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    # ---- value: state resident in HBM ----
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    launches0 = lib.lls_launch_count()
    e0.record()
    for _ in range(args.steps):
        solve_resident()
    e1.record()
    barrier()
    launches = lib.lls_launch_count() - launches0
    ms_resident = e0.elapsed_time(e1)

    # ---- e2e: host buffers in, host buffers out, every step ----
    h2d0, d2h0 = dev.bytes_h2d, dev.bytes_d2h
    barrier()
    e0.record()
    for _ in range(args.steps):
        seg = solve_e2e()
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    h2d = (dev.bytes_h2d - h2d0) // args.steps
    d2h = (dev.bytes_d2h - d2h0) // args.steps + (8 * (iters + 1) + 4 if nonlinear else 4)   # force sums + residual norm per Newton iteration + status word

    clocks = sampler.stop() if rank == 0 else None

    if world > 1:
        t = torch.tensor([ms_resident, ms_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_resident, ms_e2e = t.tolist()

    # ---- per-phase device times (instrumented pass, outside the timed regions) ----
    dev.set_timing(True)
    for _ in range(args.steps):
        solve_resident()
    tm = dev.timings()
    dev.set_timing(False)
    for k in tm:
        tm[k] /= args.steps

    # ---- parity against the reference's stored solution (rank 0's state is the fixture's state) ----
    parity = None
    if world == 1:
        gam = dev.gamma()
        parity = {"gamma_max_rel_err_vs_reference": float(np.abs(gam - g["gamma"]).max() / np.abs(g["gamma"]).max()),
                  "newton_iterations": int(iters), "reference_newton_iterations": int(meta["iterations"])}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    n, ld = dev.n, dev.ld
    peaks = (ctypes_double4())
    _native.check(lib.lls_probe_peaks(peaks, None))
    fp64_dfma, fp64_dmma, copy_gbs = peaks[0], peaks[1], peaks[3]
    mp = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            mp = json.load(fh)
    except Exception:
        pass
    hbm_peak = float(mp.get("hbm_gbs", 6650.0))
    hbm_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in mp else "fallback 6650 GB/s (B200_PROFILING.md)"

    n_fact = (iters + 1) if nonlinear else 1
    lu_flops = n_fact * (2.0 / 3.0) * n ** 3
    lu_ms = tm["lu"]
    lu_tflops = lu_flops / (lu_ms * 1e-3) * 1e-12 if lu_ms > 0 else 0.0
    pairs = float(n) * n
    build_ms = tm["influence"]
    kernels = {
        "vji_build": {"bound": "fp64", "ms": build_ms, "pair_evals_per_s": pairs / (build_ms * 1e-3) if build_ms else None,
                      "achieved_tflops_at_166_flop_per_pair": 166.0 * pairs / (build_ms * 1e-3) * 1e-12 if build_ms else None,
                      "peak_tflops": fp64_dfma, "frac": (166.0 * pairs / (build_ms * 1e-3) * 1e-12 / fp64_dfma) if build_ms else None,
                      "write_gbs": 24.0 * pairs / (build_ms * 1e-3) * 1e-9 if build_ms else None},
        "matvec_residual": {"bound": "hbm", "ms_total": tm["residual"], "launches": (iters + 2) if nonlinear else 1,
                            "achieved_gbs": ((iters + 2) if nonlinear else 1) * 24.0 * pairs / (tm["residual"] * 1e-3) * 1e-9 if tm["residual"] else None,
                            "peak_gbs": hbm_peak},
        "assembly": {"bound": "hbm", "ms_total": tm["assemble"], "launches": n_fact,
                     "achieved_gbs": n_fact * 32.0 * pairs / (tm["assemble"] * 1e-3) * 1e-9 if tm["assemble"] else None, "peak_gbs": hbm_peak},
        "integrate_ms": tm["integrate"],
    }
    total_units = world * args.steps
    value = total_units / (ms_resident * 1e-3)
    e2e_value = total_units / (ms_e2e * 1e-3)

    cpu_v, cpu_total, cpu_parts = (None, None, None)
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cpu_v, cpu_total, cpu_parts = cpu_sample(tables, ac_state, meta, threads)
        cpu_baseline = {"value": cpu_v, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": "oracle port (numpy/LAPACK restatement of the reference path): V_ji rows for 256 of N control points "
                                  "scaled to N + 1 linear assembly/solve + 1 residual + 1 Jacobian + 1 dense solve, combined with the "
                                  "reference's Newton iteration count (%d) into one solve" % iters,
                        "seconds_per_solve_extrapolated": cpu_total, "parts": cpu_parts}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms_resident / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload, tables, meta),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"kernel": "lu_step_kernel: blocked FP64 LU factorisation, one launch per 64-column block step (trailing update on DMMA mma.sync m8n8k4, next panel factored in the same launch) + lu_backsolve_chain_kernel",
                     "bound": "tensor", "achieved": lu_tflops, "peak": fp64_dmma, "unit": "TFLOP/s",
                     "frac": lu_tflops / fp64_dmma if fp64_dmma else None, "traffic": NCU_LU_TRAFFIC["bytes_per_launch"],
                     "traffic_detail": NCU_LU_TRAFFIC,
                     "peak_source": "FP64 DMMA probe run live by this process (MEASURED_PEAKS.json has no FP64 entry)",
                     "flops_per_solve": lu_flops, "ms_per_solve": lu_ms, "factorisations_per_solve": n_fact},
        "kernels": kernels,
        "peaks": {"fp64_dfma_tflops": fp64_dfma, "fp64_dmma_tflops": fp64_dmma, "copy_gbs_live": copy_gbs, "hbm_gbs": hbm_peak, "hbm_source": hbm_src},
        "phase_ms": tm,
        "parity": parity,
        "cpu_baseline": cpu_baseline,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def ctypes_double4():
    import ctypes
    return (ctypes.c_double * 4)()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "testplane", "swarm"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
