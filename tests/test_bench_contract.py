"""bench.py on CPU: the workload selection rule, and the reference arm's JSON line (the UNMODIFIED reference staged in oracle/_ref,
timed on the small unit-test aeroplane so the test stays short).  No GPU is touched."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class _Args:
    def __init__(self, workload="auto", gpus=1):
        self.workload, self.gpus = workload, gpus


def test_workload_selection(monkeypatch):
    import bench
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    monkeypatch.setattr(bench, "visible_gpus", lambda: 1)
    assert bench.pick_workload(_Args()) == "c2"                      # BASELINE.json configs[1]: the one-GPU headline
    assert bench.pick_workload(_Args(gpus=4)) == "swarm"             # N > 1: one large scene, rows partitioned (strong scaling)
    monkeypatch.setattr(bench, "visible_gpus", lambda: 8)
    assert bench.pick_workload(_Args()) == "swarm"                   # N = 1 on a multi-GPU box: base point of the scaling series
    assert bench.pick_workload(_Args(workload="c2")) == "c2"         # explicit choice wins
    monkeypatch.setenv("WORLD_SIZE", "2")
    monkeypatch.setattr(bench, "visible_gpus", lambda: 1)
    assert bench.pick_workload(_Args()) == "swarm"


def _staged():
    if os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "machupX", "scene.py")):
        return True
    if os.path.isdir("/root/reference/machupX"):
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import stage_ref
        return stage_ref.stage() is not None
    return False


def test_reference_arm_line_on_the_unit_test_aeroplane():
    if not _staged():
        pytest.skip("the reference is not staged (oracle/stage_ref.py needs /root/reference)")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "testplane", "--steps", "20",
                          "--warmup", "5"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "lifting-line solves/sec" and line["unit"] == "solves/s"
    assert line["higher_is_better"] is True and line["value"] > 0
    assert line["steps"] == 1 and line["steps_requested"] == 20          # what was timed, and what the driver asked for
    base = line["cpu_baseline"]
    assert base["kind"] == "reference" and base["cores"] >= 1 and "UNMODIFIED reference" in base["sample"]
    both = base["thread_settings"]
    assert both["as_shipped_1_blas_thread"]["seconds_per_solve"] > 0 and both["all_cores"]["seconds_per_solve"] > 0
    # the arm checks itself against the committed fixture: the staged reference gives the reference's CL
    assert abs(both["parity_with_fixture"]["CL"] - both["parity_with_fixture"]["CL_fixture"]) < 1e-12
    assert line["e2e"] == {"value": line["value"], "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["ms_per_step"] * line["steps"] / 1e3 <= line["wall_s"] + 1e-6       # consistent with the run's own clock
