"""Row-partitioned solve (machupx_b200/partitioned.py): the orchestration that runs on every rank — block ownership,
which rank factors / broadcasts / updates what and when, the Newton loop bookkeeping — driven on CPU with the gloo
backend (world_size 2 and 3) and a numpy implementation of the local work.  The device kernels behind the same ``ops``
interface are tested on the GPU box (scripts/dist_check.py under torchrun, tests/test_partitioned_gpu.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from machupx_b200.partitioned import (NB, LocalComm, PartitionedSolver, TorchComm, local_blocks, owner_of_block,
                                      padded_ld_partitioned, repad_tables)


def _system(n, seed=0):
    rng = np.random.default_rng(seed)
    A = rng.normal(size=(n, n)) / np.sqrt(n) + np.diag(3.0 + rng.random(n))
    b = rng.normal(size=n)
    c = 0.05 * rng.random(n)
    return A, b, c


class NumpyOps:
    """Local work of one rank in numpy: this rank holds the block rows ``local_blocks`` of the padded matrix.
    The toy nonlinear system is F(g) = A g + c g^3 - b (Jacobian A + diag(3 c g^2)); its linear start is A g = b."""

    def __init__(self, n, world, rank, seed=0):
        self.n, self.world, self.rank = n, world, rank
        self.ld = padded_ld_partitioned(n, world)
        self.nblk = self.ld // NB
        A, b, c = _system(n, seed)
        self.A = np.eye(self.ld); self.A[:n, :n] = A
        self.b = np.zeros(self.ld); self.b[:n] = b
        self.c = np.zeros(self.ld); self.c[:n] = c
        self.blocks = local_blocks(self.nblk, world, rank)
        self.rows = np.concatenate([np.arange(k * NB, (k + 1) * NB) for k in self.blocks])
        self.M = None                                  # local rows of the current matrix
        self.rhs = np.zeros(self.ld)                   # only owned entries are meaningful
        self.x = torch.zeros(self.ld, dtype=torch.float64)
        self.gamma = np.zeros(self.ld)
        self.step = torch.zeros(NB * NB + NB + self.nblk * NB * NB, dtype=torch.float64)
        self.log = []

    def flow_and_influence(self):
        self.log.append("influence")

    def assemble(self, newton):
        rows = self.rows
        if newton:
            self.M = self.A[rows].copy()
            self.M[np.arange(len(rows)), rows] += 3.0 * self.c[rows] * self.gamma[rows] ** 2
            # rhs = -F was stored by residual_local_sum
        else:
            self.M = self.A[rows].copy()
            self.rhs[rows] = self.b[rows]

    def residual_local_sum(self):
        rows = self.rows
        F = self.A[rows] @ self.gamma + self.c[rows] * self.gamma[rows] ** 3 - self.b[rows]
        self.rhs[rows] = -F
        return float(np.sum(F[rows < self.n] ** 2))

    def update_gamma(self, relaxation):
        x = self.x.numpy()
        self.gamma = x.copy() if relaxation < 0 else self.gamma + relaxation * x

    def lu_begin(self):
        pass

    def _local(self, k):
        return self.blocks.index(k) * NB

    def step_buffer(self, k):
        return self.step[:NB * NB + NB + (self.nblk - 1 - k) * NB * NB]

    def lu_owner(self, k, buf):
        assert owner_of_block(k, self.world) == self.rank
        r0 = self._local(k)
        D = self.M[r0:r0 + NB, k * NB:(k + 1) * NB]
        for p in range(NB):                             # unpivoted LU of the diagonal tile, in place
            D[p + 1:, p] /= D[p, p]
            D[p + 1:, p + 1:] -= np.outer(D[p + 1:, p], D[p, p + 1:])
        Lk = np.tril(D, -1) + np.eye(NB)
        Urow = np.linalg.solve(Lk, self.M[r0:r0 + NB, (k + 1) * NB:])
        self.M[r0:r0 + NB, (k + 1) * NB:] = Urow
        yk = np.linalg.solve(Lk, self.rhs[k * NB:(k + 1) * NB])
        self.rhs[k * NB:(k + 1) * NB] = yk
        out = buf.numpy()
        out[:NB * NB] = D.ravel()
        out[NB * NB:NB * NB + NB] = yk
        out[NB * NB + NB:] = Urow.reshape(NB, -1, NB).transpose(1, 0, 2).ravel()     # tile-packed, like the device

    def lu_update(self, k, buf):
        data = buf.numpy()
        U11 = np.triu(data[:NB * NB].reshape(NB, NB))
        yk = data[NB * NB:NB * NB + NB]
        ncols = self.nblk - 1 - k
        Urow = data[NB * NB + NB:].reshape(ncols, NB, NB).transpose(1, 0, 2).reshape(NB, ncols * NB) if ncols else np.zeros((NB, 0))
        for lb, g in enumerate(self.blocks):
            if g <= k:
                continue
            r0 = lb * NB
            Lik = np.linalg.solve(U11.T, self.M[r0:r0 + NB, k * NB:(k + 1) * NB].T).T
            self.M[r0:r0 + NB, k * NB:(k + 1) * NB] = Lik
            self.M[r0:r0 + NB, (k + 1) * NB:] -= Lik @ Urow
            self.rhs[g * NB:(g + 1) * NB] -= Lik @ yk

    def back(self, k):
        r0 = self._local(k)
        U = self.M[r0:r0 + NB]
        x = self.x.numpy()
        acc = self.rhs[k * NB:(k + 1) * NB] - U[:, (k + 1) * NB:] @ x[(k + 1) * NB:]
        x[k * NB:(k + 1) * NB] = np.linalg.solve(np.triu(U[:, k * NB:(k + 1) * NB]), acc)

    def x_block(self, k):
        return self.x[k * NB:(k + 1) * NB]


def _reference(n, seed, relaxation, convergence, max_iterations):
    A, b, c = _system(n, seed)
    g = np.linalg.solve(A, b)
    it, err = 0, 100.0
    while err > convergence:
        it += 1
        F = A @ g + c * g ** 3 - b
        err = np.linalg.norm(F)
        g = g + relaxation * np.linalg.solve(A + np.diag(3 * c * g ** 2), -F)
        if it >= max_iterations:
            break
    return g, it


def test_partition_helpers():
    assert padded_ld_partitioned(4000, 8) == 4096 and padded_ld_partitioned(64000, 8) == 64000 and padded_ld_partitioned(1, 3) == 192
    for world in (1, 2, 3, 8):
        nblk = 24 * world // world * world if world != 3 else 24
        seen = sorted(b for r in range(world) for b in local_blocks(nblk, world, r))
        assert seen == list(range(nblk))
        assert all(owner_of_block(b, world) == r for r in range(world) for b in local_blocks(nblk, world, r))
    tables = {"n": 100, "ld": 128, "tab": np.ones((60, 128)), "itab": np.ones((7, 128), dtype=np.int32)}
    out = repad_tables(tables, 192)
    assert out["ld"] == 192 and out["tab"].shape == (60, 192) and (out["itab"][2, 100:] == -1).all() and (out["tab"][:, :100] == 1).all()


def test_single_rank_solver_matches_numpy():
    n = 150
    ops = NumpyOps(n, 1, 0, seed=3)
    solver = PartitionedSolver(ops, LocalComm(), ops.nblk, 1, 0)
    it, err, failed, hist = solver.solve("nonlinear", 1e-11, 0.9, 50)
    g_ref, it_ref = _reference(n, 3, 0.9, 1e-11, 50)
    assert it == it_ref and not failed and np.abs(ops.gamma[:n] - g_ref).max() < 1e-12


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ops = NumpyOps(n, world, rank, seed=5)
        solver = PartitionedSolver(ops, TorchComm(), ops.nblk, world, rank)
        it, err, failed, hist = solver.solve("nonlinear", 1e-11, 0.9, 50)
        q.put((rank, ops.gamma[:n].copy(), it, err, failed, hist))
    except Exception as e:   # noqa: BLE001
        q.put((rank, repr(e), None, None, None, None))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 200), (3, 130), (2, 64)])
def test_gloo_partitioned_newton(world, n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g_ref, it_ref = _reference(n, 5, 0.9, 1e-11, 50)
    for rank, gamma, it, err, failed, hist in results:
        assert not isinstance(gamma, str), gamma
        assert it == it_ref and not failed
        assert np.abs(gamma - g_ref).max() < 1e-12       # every rank ends with the same, correct, replicated solution
    assert all(np.array_equal(results[0][1], r[1]) for r in results[1:])
