"""Row-partitioned solve on the GPU box.  With one GPU the partitioned kernels run with world size 1;
whenever two or more GPUs are visible the same script runs under torchrun, one process per GPU, and must reproduce the
reference's circulation, forces and iteration counts -- through the peer-memory exchange (panels stored straight into the
other GPUs' mailboxes over NVLink by the owner kernel, look-ahead of one panel) and through the step-wise path with NCCL
broadcasts, which is what the gloo CPU tests (tests/test_partitioned_cpu.py) drive."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIXTURES = ["testplane_nonlinear", "traditional_linear", "swarm_nonlinear"]


def _parse(stdout, fixtures=FIXTURES):
    rows = [json.loads(line) for line in stdout.splitlines() if line.startswith("{")]
    assert len(rows) == len(fixtures), stdout[-2000:]
    for r in rows:
        assert r["ok"], r
        assert r["gamma_rel_err"] < 1e-10 and r["force_rel_err"] < 1e-10 and r["iterations"] == r["reference_iterations"]
    return rows


def _run(world, fixtures, p2p, port, pair_min_rem=None):
    env = dict(os.environ, LLS_PARTITIONED_P2P="1" if p2p else "0")
    if pair_min_rem is not None:      # rank-128 paired visits from this many block rows on (the default threshold is far above the fixtures)
        env["LLS_PARTITIONED_PAIR_MIN_REM"] = str(pair_min_rem)
    script = os.path.join(ROOT, "scripts", "dist_check.py")
    if world == 1:
        cmd = [sys.executable, script] + fixtures
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
               "--master-port", str(port), script] + fixtures
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    rows = _parse(out.stdout, fixtures)
    assert all(r["world"] == world and r["fused_exchange"] == p2p for r in rows), rows
    return rows


@pytest.mark.parametrize("p2p", [True, False])
def test_partitioned_world_size_1(p2p):
    _run(1, FIXTURES, p2p, 0)


def test_partitioned_paired_visits_world_size_1():
    """Thin launch + rank-128 visit for every pair of block steps (threshold forced down to 2 block rows), incl. N = 4000."""
    _run(1, FIXTURES + ["c2_traditional_n4000"], True, 0, pair_min_rem=2)


def _visible_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("p2p", [True, False])
def test_partitioned_multi_gpu(p2p):
    """Runs whenever the box shows at least two GPUs (the driver's 1-GPU box skips it; `gpurun --gpus 2/4/8` does not)."""
    n = _visible_gpus()
    if n < 2:
        pytest.skip("one GPU visible: the multi-GPU exchange needs `gpurun --gpus 2` or more")
    for world in sorted({2, min(4, n), min(8, n)}):
        if world < 2 or world > n:
            continue
        rows = _run(world, FIXTURES + (["c2_traditional_n4000"] if world == 2 else []), p2p, 29533 + world)
        if p2p:
            assert all(r["exchange_stats"]["block_steps"] > 0 for r in rows)
            _run(world, FIXTURES + (["c2_traditional_n4000"] if world == 2 else []), True, 29543 + world, pair_min_rem=2)     # paired visits


def test_partitioned_world_sizes_agree_on_a_lattice_swarm():
    """Sizes the reference cannot run: the 8-aircraft lattice swarm (N = 8000) must give the same forces for every world size the
    box offers, with a converged residual (cross-world-size agreement + residual + oracle rows)."""
    n = _visible_gpus()
    worlds = [w for w in (1, 2, 4, 8) if w <= n]
    sums = []
    for world in worlds:
        script = os.path.join(ROOT, "scripts", "swarm_scale.py")
        cmd = ([sys.executable, script] if world == 1 else
               [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                "--master-port", str(29550 + world), script]) + ["8", "200", "4"]
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=dict(os.environ, SWARM_ONE_SOLVE="1"))
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
        r = [json.loads(line) for line in out.stdout.splitlines() if line.startswith("{")][-1]
        assert r["control_points"] == 8000 and r["final_residual"] < 1e-10 and r["status"] == 0
        assert r["V_max_abs_err_over_scale_rank0"] < 1e-9
        sums.append((r["CL_sum"], r["CD_sum"], r["newton_iterations"]))
    for s in sums[1:]:
        assert abs(s[0] - sums[0][0]) < 1e-11 * abs(sums[0][0]) and abs(s[1] - sums[0][1]) < 1e-11 * abs(sums[0][1]) and s[2] == sums[0][2]


def test_one_rank_of_the_64k_swarm_partition_on_one_gpu():
    """BASELINE.json configs[4] (64 aircraft, 64,000 control points) needs several GPUs; ONE rank's share of its 8-way partition
    (8,000 rows of V_ji, 12 GB) fits one: the row-local kernels -- flow state, influence build with the ownership mapping,
    linear-system assembly -- against rows recomputed by the CPU oracle (restatement of the reference arithmetic)."""
    import numpy as np
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from workloads import swarm_input
    from machupx_b200 import Scene, layout as L
    from machupx_b200.partitioned import PartitionedDeviceScene
    from machupx_b200.tables import build_aircraft_state
    from oracle import lls_oracle as O
    scene = Scene(swarm_input(64, 200))
    assert scene._N == 64000
    world, rank = 8, 3
    dev = PartitionedDeviceScene(scene._tables, shard=(world, rank))
    ac = build_aircraft_state(scene)
    dev.upload_state(ac)
    dev.flow_and_influence()
    dev.assemble(False)
    torch.cuda.synchronize()
    T = O.Tables(scene._tables, ac)
    F = O.flow_state(T)
    n = scene._N
    worst, scale, worst_a = 0.0, 0.0, 0.0
    for lb in (0, 57, dev.local_rows // 64 - 1):
        for r in (0, 63):
            li = lb * 64 + r
            gi = (lb * world + rank) * 64 + r
            if gi >= n:
                continue
            Vo = O.influence_rows(T, F, gi, gi + 1)[0]                                   # (n, 3)
            Vg = dev.d_V[li, :, :n].T.cpu().numpy()
            scale, worst = max(scale, np.abs(Vo).max()), max(worst, np.abs(Vg - Vo).max())
            # row gi of the linear system (scene.py:811-824) from the oracle's V row
            a_row = -(F.V_inf_in_plane[gi] * F.CLa[gi] * T.DS[gi]) * (Vo @ T.UN[gi])
            a_row[gi] += 2.0 * np.linalg.norm(np.cross(F.v_inf_and_rot[gi], T.DL[gi]))
            Ag = dev.d_M[li, :n].cpu().numpy()
            worst_a = max(worst_a, np.abs(Ag - a_row).max() / np.abs(a_row).max())
    assert worst / scale < 1e-9 and worst_a < 1e-9
    del dev
    torch.cuda.empty_cache()


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
