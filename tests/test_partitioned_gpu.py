"""Row-partitioned solve on the GPU box.  With one GPU the partitioned kernels (owner panel, local L panel, DMMA update,
chunked back substitution, ownership-masked residual / integration) run with world size 1; with two or more GPUs the same
script runs under torchrun with NCCL and must reproduce the reference's circulation, forces and iteration counts."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIXTURES = ["testplane_nonlinear", "traditional_linear", "swarm_nonlinear"]


def _parse(stdout):
    rows = [json.loads(line) for line in stdout.splitlines() if line.startswith("{")]
    assert len(rows) == len(FIXTURES), stdout[-2000:]
    for r in rows:
        assert r["ok"], r
        assert r["gamma_rel_err"] < 1e-10 and r["force_rel_err"] < 1e-10 and r["iterations"] == r["reference_iterations"]
    return rows


def test_partitioned_world_size_1():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "dist_check.py")] + FIXTURES, capture_output=True, text=True,
                         timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    _parse(out.stdout)


def test_partitioned_two_gpus_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "scripts", "dist_check.py")] + FIXTURES
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    rows = _parse(out.stdout)
    assert all(r["world"] == 2 for r in rows)


def test_scene_api_partitioned_equals_regular():
    import copy
    import numpy as np
    from machupx_b200 import Scene
    with open(os.path.join(ROOT, "tests", "golden", "swarm_nonlinear.json")) as fh:
        inp = json.load(fh)["input"]
    scene = Scene(copy.deepcopy(inp))
    FM_a = scene.solve_forces()
    gamma_a = np.array(scene._gamma)
    FM_b = scene.solve_forces(partitioned=True)
    assert np.abs(np.asarray(scene._gamma) - gamma_a).max() <= 1e-12 * np.abs(gamma_a).max()
    for name in FM_a:
        for key, val in FM_a[name]["total"].items():
            assert abs(FM_b[name]["total"][key] - val) <= 1e-11 * max(1.0, abs(val)), (name, key)
    FM_c = scene.solve_forces()          # and back to the single-GPU path
    assert abs(FM_c["drone0"]["total"]["CL"] - FM_a["drone0"]["total"]["CL"]) < 1e-13
