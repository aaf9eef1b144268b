#!/usr/bin/env python
"""Times the UNMODIFIED reference (staged by oracle/stage_ref.py into oracle/_ref) on one JSON scene: the CPU arm of bench.py.
TEST / MEASUREMENT INFRASTRUCTURE -- the product never imports this.

    python oracle/ref_runner.py <scene_input.json> [--solves K] [--all-cores-too]

Run as its own process because the reference pins BLAS to ONE thread by exporting OMP/OPENBLAS/MKL_NUM_THREADS=1 at import
(machupX/__init__.py:1-6), which only works if it is imported before numpy: `import machupX` is therefore the first import
here ("as shipped").  The second setting raises the BLAS pool to every host core at run time (threadpoolctl) on the SAME
scene object, so the 46-49 s scene construction is paid once.  One "solve" = one steady-state Scene.solve_forces() on the
built scene.  Also reported: the time the reference spends in _calc_invariant_flow_properties per solve (the
per-solve part of the V_ji build, scene.py:557-668) and the scene construction (geometry + V_ji_const, scene.py:431-549).
Prints ONE JSON line."""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "ref_stubs"))
sys.path.insert(0, os.path.join(HERE, "_ref"))

import machupX as MX  # noqa: E402  FIRST import: the reference sets the BLAS thread count to 1 here

import warnings  # noqa: E402

import numpy as np  # noqa: E402


def main():
    path = sys.argv[1]
    solves = int(sys.argv[sys.argv.index("--solves") + 1]) if "--solves" in sys.argv else 1
    all_cores = "--all-cores-too" in sys.argv
    with open(path) as fh:
        scene_input = json.load(fh)
    warnings.simplefilter("ignore")
    cores = os.cpu_count() or 1
    out = {"cores": cores, "machupX": os.path.dirname(MX.__file__), "blas_env_as_shipped": os.environ.get("OPENBLAS_NUM_THREADS")}
    t0 = time.perf_counter()
    scene = MX.Scene(scene_input)
    out["construct_s"] = time.perf_counter() - t0
    out["control_points"] = int(scene._N)

    flow = {"s": 0.0}
    inner = scene._calc_invariant_flow_properties          # instrumented from outside; the reference's code is untouched

    def timed_flow(*a, **k):
        t = time.perf_counter()
        r = inner(*a, **k)
        flow["s"] += time.perf_counter() - t
        return r
    scene._calc_invariant_flow_properties = timed_flow

    def run(label):
        times = []
        for _ in range(solves):
            flow["s"] = 0.0
            t = time.perf_counter()
            FM = scene.solve_forces()
            times.append(time.perf_counter() - t)
        name = list(FM.keys())[0]
        tot = FM[name]["total"]
        out[label] = {"seconds_per_solve": float(np.median(times)), "solves": solves, "vji_flow_seconds": flow["s"],
                      "CL": tot.get("CL"), "CD": tot.get("CD"), "Cm": tot.get("Cm")}

    if "--sweep" in sys.argv:
        # BASELINE.json configs[2]: the reference's serial idiom set_aircraft_state -> solve_forces (scene.py:1535-1598) over a list of
        # flight states; as shipped (1 BLAS thread; the 180-point systems do not thread anyway)
        with open(sys.argv[sys.argv.index("--sweep") + 1]) as fh:
            states = json.load(fh)
        name = list(scene._airplanes.keys())[0]
        t = time.perf_counter()
        cl = []
        for st in states:
            scene.set_aircraft_state(state=dict(st), aircraft=name)
            cl.append(scene.solve_forces()[name]["total"]["CL"])
        dt = time.perf_counter() - t
        out["sweep"] = {"cases": len(states), "seconds": dt, "cases_per_second": len(states) / dt, "CL_first_last": [cl[0], cl[-1]]}
        print(json.dumps(out))
        return
    run("as_shipped_1_blas_thread")
    if all_cores:
        try:
            from threadpoolctl import threadpool_info, threadpool_limits
            threadpool_limits(limits=cores)
            out["blas_threads_all_cores"] = max([d.get("num_threads", 1) for d in threadpool_info()] or [1])
            run("all_cores")
        except Exception as e:   # noqa: BLE001
            out["all_cores_error"] = repr(e)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
