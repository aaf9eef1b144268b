"""Case-sharded sweeps (machupx_b200/sweep.py): the N > 1 path on CPU with the gloo backend, world_size 2 and 3.
The per-case solver is replaced by a cheap analytic stand-in, so what is tested is the partition, the gather, the case
order, ragged shard sizes and the error policy — the parts that are not the GPU solve itself."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from machupx_b200.sweep import TOTAL_KEYS, alpha_beta_grid, run_sweep, shard_cases


def _fake_solver(i, state):
    if state.get("fail"):
        return None
    a, b = np.radians(state["alpha"]), np.radians(state["beta"])
    vals = np.array([np.sin(a + 0.01 * k) * np.cos(b) + i * 1e-6 for k in range(len(TOTAL_KEYS))])
    return vals, 10 + (i % 5)


def _expected(states):
    rows = [_fake_solver(i, s) for i, s in enumerate(states)]
    return np.array([r[0] for r in rows]), np.array([r[1] for r in rows])


def test_shard_cases_partitions_everything():
    for n in (0, 1, 7, 64, 4096, 4097):
        for world in (1, 2, 3, 8):
            spans = [shard_cases(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[r][1] == spans[r + 1][0] for r in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_single_process_sweep():
    states = alpha_beta_grid(np.linspace(-2, 10, 5), np.linspace(-6, 6, 3), 100.0)
    table, its, ok = run_sweep(states, _fake_solver)
    exp, exp_its = _expected(states)
    assert np.array_equal(table, exp) and np.array_equal(its, exp_its) and ok.all()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_alpha, n_beta, fail_case, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        states = alpha_beta_grid(np.linspace(-2, 10, n_alpha), np.linspace(-6, 6, n_beta), 100.0)
        if fail_case is not None:
            states[fail_case]["fail"] = True
        try:
            table, its, ok = run_sweep(states, _fake_solver, on_not_converged="raise" if fail_case is None else "nan")
            q.put((rank, table, its, ok))
        except Exception as e:   # noqa: BLE001
            q.put((rank, repr(e), None, None))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_alpha,n_beta,fail_case", [(2, 4, 4, None), (2, 5, 3, None), (3, 5, 5, None), (2, 3, 3, 4)])
def test_gloo_sharded_sweep(world, n_alpha, n_beta, fail_case):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_alpha, n_beta, fail_case, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    states = alpha_beta_grid(np.linspace(-2, 10, n_alpha), np.linspace(-6, 6, n_beta), 100.0)
    exp, exp_its = _expected(states)
    for rank, table, its, ok in results:
        assert not isinstance(table, str), table
        if fail_case is None:
            assert np.array_equal(table, exp) and np.array_equal(its, exp_its) and ok.all()     # every rank holds the full table, in case order
        else:
            good = np.arange(len(states)) != fail_case
            assert np.array_equal(table[good], exp[good]) and np.isnan(table[fail_case]).all()
            assert not ok[fail_case] and ok[good].all()


# ---- the batched block path: every rank's shard as ONE Scene.solve_forces_batch pass ---------------------------------------------
def _batch_worker(rank, world, port, q):
    import json
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import api_scenarios as S
        import machupx_b200.device as device
        import oracle_backend
        from machupx_b200.sweep import BatchSceneSolver
        device.BatchDeviceScene = oracle_backend.OracleBatchDeviceScene      # CPU stand-in of the device (tests only)
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "api_scenarios.json")) as fh:
            gold = json.load(fh)["c3"]
        inp = gold["input"]
        inp["solver"]["type"] = "nonlinear"
        inp["solver"].pop("relaxation", None)
        solver = BatchSceneSolver(inp)
        states = S.c3_states(S.C3_SUB[::2], S.C3_SUB[::2])
        table, its, ok = run_sweep(states, solver)
        q.put((rank, table, its, ok, solver.scene._batch_device.ncases))
    except Exception as e:   # noqa: BLE001
        q.put((rank, repr(e), None, None, None))
    finally:
        dist.destroy_process_group()


def test_gloo_sharded_batched_sweep_against_the_reference_c3_golden():
    """world_size 2 over gloo: rank r solves its contiguous half of a 4 x 4 sub-grid of BASELINE.json configs[2] as one batch
    (device replaced by the oracle stand-in), one all_gather, and every rank holds the reference's totals and iteration
    counts (tests/golden/api_scenarios.json "c3", generated by the reference's serial loop)."""
    import json
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import api_scenarios as S
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "api_scenarios.json")) as fh:
        gold = json.load(fh)["c3"]["nonlinear"]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    world = 2
    procs = [ctx.Process(target=_batch_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pick = [8 * a + b for a in range(0, 8, 2) for b in range(0, 8, 2)]          # positions of the 4 x 4 cases in the 8 x 8 golden
    ref = np.array([[gold["totals"][i][k] for k in TOTAL_KEYS] for i in pick])
    ref_its = np.array([gold["iterations"][i] for i in pick])
    for rank, table, its, ok, ncases in results:
        assert not isinstance(table, str), table
        assert ncases == 8                                                      # the rank's whole shard in one batch
        assert ok.all() and np.array_equal(its, ref_its)
        assert np.abs(table - ref).max() <= 1e-10 * max(1.0, np.abs(ref).max())
