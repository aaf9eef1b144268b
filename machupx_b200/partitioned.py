"""One large scene with the rows of V_ji / A / J partitioned over the GPUs of a box (SURVEY.md section 8e, row 2;
BASELINE.json configs[3]-[4]).

Ownership: 64-row blocks dealt round-robin, rank r owns global blocks r, r + P, r + 2P ... (block-cyclic, so the trailing
update of the LU stays balanced as the active window shrinks).  Per-point tables, the circulation, the right-hand side
and the Newton step stay replicated (O(N)).  The library (include/lls.h, "partitioned" entry points) does the local work;
the collectives are issued here through ``torch.distributed`` (NCCL over NVLink on the box, gloo in the CPU tests):

  * LU block step k: the owner factors the diagonal tile and solves its U block row, packs [diag image | y_k | U tiles] into
    one step buffer, ONE broadcast of that buffer, then every rank solves its part of the L panel and updates its rows.
    Total broadcast volume per factorisation: 4 N^2 bytes (16 GB at N = 64k).
  * back substitution: the owner of block row k forms x_k from its rows of U and the replicated x_{>k}; 64 doubles are
    broadcast per block.
  * residual norm and per-segment force sums: one all-reduce each.

``PartitionedSolver`` holds the orchestration and is written against a small ``ops`` interface so that the CPU tests can
drive exactly the same control flow with a numpy implementation of the local work (tests/test_partitioned_cpu.py).
"""
import ctypes
import os

import numpy as np

from . import layout as L

NB = 64


def owner_of_block(k, world):
    return k % world


def local_blocks(nblk, world, rank):
    """Global block indices stored on ``rank``, in local order."""
    return list(range(rank, nblk, world))


def padded_ld_partitioned(n, world):
    q = NB * int(world)
    return (int(n) + q - 1) // q * q


def repad_tables(tables, ld_new):
    """Copy of ``tables`` with the leading dimension grown to ``ld_new`` (same benign padding as tables.build_tables)."""
    n, ld = int(tables["n"]), int(tables["ld"])
    if ld_new == ld:
        return tables
    assert ld_new > ld
    tab = np.zeros((L.NF, ld_new))
    itab = np.zeros((L.NI, ld_new), dtype=np.int32)
    tab[:, :n] = np.asarray(tables["tab"])[:, :n]
    itab[:, :n] = np.asarray(tables["itab"])[:, :n]
    tab[L.F_UA, n:] = 1.0
    tab[L.F_UN + 1, n:] = 1.0
    tab[L.F_US + 2, n:] = 1.0
    for f in (L.F_CBAR, L.F_DS, L.F_CSWI, L.F_RHO, L.F_NU, L.F_SOS):
        tab[f, n:] = 1.0
    tab[L.F_PC, n:] = 1e30
    itab[L.I_WING_BEGIN, n:] = -1
    itab[L.I_WING_END, n:] = -1
    itab[L.I_AIRCRAFT, n:] = -1
    out = dict(tables)
    out.update(tab=tab, itab=itab, ld=int(ld_new))
    return out


class PartitionedSolver:
    """Control flow of a partitioned solve.  ``ops`` supplies the local work (device: ``PartitionedDeviceScene``; tests: numpy),
    ``comm`` the collectives (``broadcast(tensor, src)``, ``all_reduce_sum(array) -> array``)."""

    def __init__(self, ops, comm, nblk, world, rank):
        self.ops, self.comm, self.nblk, self.world, self.rank = ops, comm, int(nblk), int(world), int(rank)

    def lu_solve(self):
        """Factor the matrix the ranks hold, carry the right-hand side along, back-substitute: x replicated on return."""
        ops, comm = self.ops, self.comm
        if getattr(ops, "fused_exchange", False):
            # device path: the whole factorisation + back substitution is enqueued by ONE library call; the panels travel between
            # the GPUs inside the kernels (peer-memory mailboxes), so there is nothing for the host to do per block step
            ops.lu_solve_fused()
            return
        ops.lu_begin()
        for k in range(self.nblk):
            owner = owner_of_block(k, self.world)
            buf = ops.step_buffer(k)
            if owner == self.rank:
                ops.lu_owner(k, buf)
            comm.broadcast(buf, owner)
            ops.lu_update(k, buf)
        for k in range(self.nblk - 1, -1, -1):
            owner = owner_of_block(k, self.world)
            if owner == self.rank:
                ops.back(k)
            comm.broadcast(ops.x_block(k), owner)

    def residual_norm(self):
        return float(np.sqrt(comm_sum(self.comm, self.ops.residual_local_sum())))

    def solve(self, solver_type, convergence, relaxation, max_iterations, linear_start=True):
        """Same loop as lls_solve_linear + lls_solve_nonlinear (reference scene.py:800-829, 851-913).
        Returns (iterations, final_error, not_converged, history)."""
        ops = self.ops
        ops.flow_and_influence()
        if solver_type == "linear" or linear_start:
            ops.assemble(False)
            self.lu_solve()
            ops.update_gamma(-1.0)
        it, error, hist, failed = 0, 100.0, [], False
        if solver_type == "nonlinear":
            while error > convergence:
                it += 1
                error = self.residual_norm()
                hist.append(error)
                ops.assemble(True)
                self.lu_solve()
                ops.update_gamma(relaxation)
                if it >= max_iterations:
                    failed = True
                    break
            error = self.residual_norm()
        return it, error, failed, hist


def comm_sum(comm, value):
    return float(comm.all_reduce_sum(np.array([value], dtype=np.float64))[0])


class LocalComm:
    """world_size 1 without a process group: the collectives are the identity."""

    def broadcast(self, tensor, src):
        pass

    def all_reduce_sum(self, array):
        return np.asarray(array, dtype=np.float64)


class TorchComm:
    """Collectives through torch.distributed on the tensors' own device (NCCL for CUDA tensors, gloo for CPU tensors)."""

    def __init__(self, group=None, device=None):
        import torch.distributed as dist
        self.dist, self.group, self.device = dist, group, device
        self.ranks = list(range(dist.get_world_size(group))) if group is None else dist.get_process_group_ranks(group)

    def broadcast(self, tensor, src):
        self.dist.broadcast(tensor, src=self.ranks[src], group=self.group)

    def all_reduce_sum(self, array):
        import torch
        t = torch.from_numpy(np.ascontiguousarray(array, dtype=np.float64))
        if self.device is not None:
            t = t.to(self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()


class PartitionedDeviceScene:
    """Device side of a partitioned scene on this rank's GPU; also the ``ops`` object of ``PartitionedSolver``."""

    def __init__(self, tables, group=None, device=None, shard=None):
        """``shard = (world, rank)`` builds ONE rank's share of a partition without a process group (no exchange possible: for
        checking the row-local kernels -- influence build, assembly, residual -- of scenes that need several GPUs on a single one)."""
        import torch
        import torch.distributed as dist
        from . import _native
        if not torch.cuda.is_available():
            raise _native.LlsError("machupx_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.torch, self._native = torch, _native
        self.lib = _native.load()
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if shard is not None:
            self.world, self.rank = int(shard[0]), int(shard[1])
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.n = int(tables["n"])
        sizes = _native.Sizes()
        _native.check(self.lib.lls_query_sizes_partitioned(self.n, int(tables["n_segments"]), self.world, ctypes.byref(sizes)))
        self.ld = int(sizes.ld)
        tables = repad_tables(tables, self.ld)
        self.nblk = self.ld // NB
        self.local_rows = self.ld // self.world
        self.n_aircraft, self.n_segments, self.flags = int(tables["n_aircraft"]), int(tables["n_segments"]), int(tables["flags"])
        af = np.ascontiguousarray(tables["airfoils"], dtype=np.float64).reshape(-1, L.AF_STRIDE)
        dev = self.device
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        with torch.cuda.device(dev):
            self.handle = ctypes.c_void_p()
            _native.check(self.lib.lls_create(ctypes.byref(self.handle), self.n, self.n_aircraft, self.n_segments, af.shape[0], self.flags))
            _native.check(self.lib.lls_set_partition(self.handle, self.world, self.rank))
            self.d_tab = torch.from_numpy(np.ascontiguousarray(tables["tab"], dtype=np.float64)).to(dev)
            self.d_itab = torch.from_numpy(np.ascontiguousarray(tables["itab"], dtype=np.int32)).to(dev)
            self.d_af = torch.from_numpy(af).to(dev)
            self.d_out = torch.zeros((L.NO, self.ld), dtype=torch.float64, device=dev)
            self.d_V = torch.empty((self.local_rows, 3, self.ld), dtype=torch.float64, device=dev)
            self.d_M = torch.empty((self.local_rows, self.ld), dtype=torch.float64, device=dev)
            self.d_work = torch.zeros(int(sizes.work_bytes), dtype=torch.uint8, device=dev)
            self.d_ac = torch.zeros((self.n_aircraft, L.AC_STRIDE), dtype=torch.float64, device=dev)
            self.d_seg_fm = torch.zeros((self.n_segments, 12), dtype=torch.float64, device=dev)
            _native.check(self.lib.lls_bind(self.handle, p(self.d_tab), p(self.d_itab), p(self.d_af), p(self.d_out), p(self.d_V), p(self.d_M),
                                            p(self.d_work), ctypes.c_size_t(int(sizes.work_bytes))))
            _native.check(self.lib.lls_set_state(self.handle, p(self.d_ac)))
            step = ctypes.c_size_t(0)
            _native.check(self.lib.lls_dist_lu_begin(self.handle, ctypes.byref(step), self._st()))
            self.d_step = torch.zeros(int(step.value), dtype=torch.float64, device=dev)
            rhs, x, gam = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
            _native.check(self.lib.lls_dist_vectors(self.handle, ctypes.byref(rhs), ctypes.byref(x), ctypes.byref(gam)))
            off = x.value - self.d_work.data_ptr()
            self.d_x = self.d_work[off:off + self.ld * 8].view(torch.float64)      # replicated Newton step / solution
        self.impingement_threshold = 1e-10
        self.comm = TorchComm(group, dev) if dist.is_initialized() and shard is None else LocalComm()
        self.fused_exchange = self._connect_mailboxes(group) if shard is None else False
        self.solver = PartitionedSolver(self, self.comm, self.nblk, self.world, self.rank)
        # Step-wise fall-back path only (LLS_PARTITIONED_P2P=0 or no peer access): one factorisation + back substitution = 3 launches
        # and 1-2 NCCL broadcasts per block step issued from here; optionally captured into a CUDA graph after a warm-up pass.
        self.use_graph = os.environ.get("LLS_PARTITIONED_GRAPH", "0") == "1"   # off by default: capture with NCCL hung on the 2-GPU box (round 1)
        self._lu_graph = None
        self._lu_calls = 0
        plain = self.solver.lu_solve

        def lu_solve_graphed():
            self._lu_calls += 1
            if not self.use_graph or self._lu_calls == 1:
                return plain()                      # first pass warms NCCL and the kernels up
            if self._lu_graph is None:
                try:
                    torch.cuda.synchronize(self.device)
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        plain()
                    self._lu_graph = g
                except Exception as exc:            # capture not possible in this environment: keep the plain loop
                    self.use_graph = False
                    self.graph_error = repr(exc)
                    torch.cuda.synchronize(self.device)
                    return plain()
            self._lu_graph.replay()

        self.solver.lu_solve = lu_solve_graphed

    def _connect_mailboxes(self, group):
        """Peer-memory exchange of the LU (include/lls.h, lls_dist_comm_*): every rank allocates its mailbox, the 64-byte CUDA IPC
        handles are gathered through torch.distributed, every rank maps the others' mailboxes.  Returns False -- on EVERY rank --
        if any rank cannot (no peer access between the GPUs, IPC not permitted): the solve then uses the step-wise path with
        NCCL broadcasts.  ``LLS_PARTITIONED_P2P=0`` forces that path."""
        torch = self.torch
        import torch.distributed as dist
        if os.environ.get("LLS_PARTITIONED_P2P", "1") != "1":
            return False
        self.multicast = False
        if self.world > 1 and os.environ.get("LLS_PARTITIONED_SYMM", "1") == "1" and self._attach_symmetric_memory(group):
            return True
        ok = 1
        handle = (ctypes.c_ubyte * 64)()
        try:
            self._call(self.lib.lls_dist_comm_create, int(os.environ.get("LLS_PARTITIONED_SLOTS", "8")), handle)
        except self._native.LlsError as exc:
            ok, self.exchange_error = 0, repr(exc)
        if self.world > 1:
            mine = torch.tensor(list(bytes(handle)) + [ok], dtype=torch.uint8, device=self.device)
            every = torch.empty(self.world * 65, dtype=torch.uint8, device=self.device)
            dist.all_gather_into_tensor(every, mine, group=group)
            every = every.cpu().numpy().reshape(self.world, 65)
            ok = int(every[:, 64].min())
            if ok:
                blob = (ctypes.c_ubyte * (64 * self.world)).from_buffer_copy(np.ascontiguousarray(every[:, :64]).tobytes())
                try:
                    self._call(self.lib.lls_dist_comm_connect, blob)
                except self._native.LlsError as exc:
                    ok, self.exchange_error = 0, repr(exc)
            flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            ok = int(flag.item())
        if not ok:
            self._native.check(self.lib.lls_dist_comm_destroy(self.handle))
        return bool(ok)

    def _attach_symmetric_memory(self, group):
        """Preferred plumbing of the mailboxes: torch.distributed's symmetric memory (every rank's buffer mapped into every process,
        plus -- on NVSwitch systems -- ONE multicast mapping of all of them, NVLS).  With the multicast mapping an owner kernel stores
        a panel once and the switch replicates it into all mailboxes.  Returns False on every rank if any rank cannot (old driver, no
        fabric support, API missing): the caller then falls back to CUDA IPC handles (unicast stores to every peer)."""
        torch = self.torch
        import torch.distributed as dist
        slots = int(os.environ.get("LLS_PARTITIONED_SLOTS", "8"))
        ok, hdl, buf = 1, None, None
        try:
            import torch.distributed._symmetric_memory as symm
            nbytes = ctypes.c_size_t(0)
            self._call(self.lib.lls_dist_comm_bytes, slots, ctypes.byref(nbytes))
            with torch.cuda.device(self.device):
                buf = symm.empty(int(nbytes.value), dtype=torch.uint8, device=self.device)
                buf.zero_()
                torch.cuda.synchronize(self.device)
                grp = group if group is not None else dist.group.WORLD
                hdl = symm.rendezvous(buf, grp)
            ptrs = [int(p) for p in hdl.buffer_ptrs]
            if len(ptrs) != self.world or int(hdl.rank) != self.rank or ptrs[self.rank] != buf.data_ptr():
                raise RuntimeError("symmetric memory handle does not match the process group")
        except Exception as exc:   # noqa: BLE001
            ok, self.exchange_error = 0, "symmetric memory: " + repr(exc)
        flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            return False
        use_mc = os.environ.get("LLS_PARTITIONED_MULTICAST", "1") == "1" and bool(getattr(hdl, "has_multicast_support", False))
        mc = int(hdl.multicast_ptr) if use_mc else 0
        flag = torch.tensor([1 if mc else 0], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)          # multicast on every rank or on none
        if int(flag.item()) == 0:
            mc = 0
        arr = (ctypes.c_void_p * self.world)(*ptrs)
        self._call(self.lib.lls_dist_comm_attach, slots, arr, ctypes.c_void_p(mc) if mc else None)
        dist.barrier(group=group)                                         # every mailbox is zero-filled and attached before anyone stores
        self._symm = (buf, hdl)                                           # keep the mappings alive as long as the scene
        self.multicast = bool(mc)
        return True

    def lu_solve_fused(self):
        self._call(self.lib.lls_dist_lu_solve, self._st())

    def exchange_stats(self):
        """{panel_wait_ms, slot_wait_ms, block_steps, back_substitutions} since the last call (fused exchange only)."""
        out = (ctypes.c_double * 4)()
        self._call(self.lib.lls_dist_comm_stats, out, self._st())
        return {"panel_wait_ms": out[0], "slot_wait_ms": out[1], "block_steps": int(out[2]), "back_substitutions": int(out[3])}

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.lls_destroy(self.handle)
                self.handle = None
            self._symm = None
        except Exception:
            pass

    def _st(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def _call(self, fn, *args):
        with self.torch.cuda.device(self.device):
            self._native.check(fn(self.handle, *args))

    def upload_tables(self, tables):
        tables = repad_tables(tables, self.ld)
        af = np.ascontiguousarray(tables["airfoils"], dtype=np.float64).reshape(-1, L.AF_STRIDE)
        self.d_tab.copy_(self.torch.from_numpy(np.ascontiguousarray(tables["tab"], dtype=np.float64)))
        self.d_itab.copy_(self.torch.from_numpy(np.ascontiguousarray(tables["itab"], dtype=np.int32)))
        self.d_af.copy_(self.torch.from_numpy(af))

    def upload_rows(self, tab, fields):
        n = self.n
        for f in fields:
            self.d_tab[f, :n].copy_(self.torch.from_numpy(np.ascontiguousarray(tab[f][:n], dtype=np.float64)))

    def upload_state(self, ac_state):
        self.d_ac.copy_(self.torch.from_numpy(np.ascontiguousarray(L.pad_ac_state(ac_state)).reshape(self.n_aircraft, L.AC_STRIDE)))

    # -- ops interface --------------------------------------------------------------------------------------------
    def flow_and_influence(self):
        self._call(self.lib.lls_flow_state, self._st())
        self._call(self.lib.lls_build_influence, float(self.impingement_threshold), self._st())

    def assemble(self, newton):
        self._call(self.lib.lls_dist_assemble, 1 if newton else 0, self._st())

    def residual_local_sum(self):
        s = ctypes.c_double(0.0)
        self._call(self.lib.lls_dist_residual, self._st(), ctypes.byref(s))
        return s.value

    def update_gamma(self, relaxation):
        self._call(self.lib.lls_dist_update_gamma, float(relaxation), self._st())

    def lu_begin(self):
        self._call(self.lib.lls_dist_lu_begin, None, self._st())

    def step_buffer(self, k):
        nd = ctypes.c_size_t(0)
        self._native.check(self.lib.lls_dist_lu_step_doubles(self.handle, int(k), ctypes.byref(nd)))
        return self.d_step[:int(nd.value)]

    def lu_owner(self, k, buf):
        self._call(self.lib.lls_dist_lu_owner, int(k), 1, ctypes.c_void_p(buf.data_ptr()), self._st())

    def lu_update(self, k, buf):
        self._call(self.lib.lls_dist_lu_update, int(k), 1, ctypes.c_void_p(buf.data_ptr()), self._st())

    def back(self, k):
        self._call(self.lib.lls_dist_back, int(k), self._st())

    def x_block(self, k):
        return self.d_x[k * NB:(k + 1) * NB]

    # -- whole solve ------------------------------------------------------------------------------------------------
    def solve(self, solver_type, convergence, relaxation, max_iterations, impingement_threshold=1e-10):
        """Returns (seg_fm (n_segments, 12) summed over ranks, iterations, final_error, status bits, history)."""
        torch = self.torch
        self.impingement_threshold = impingement_threshold
        it, err, failed, hist = self.solver.solve(solver_type, convergence, relaxation, max_iterations)
        self._call(self.lib.lls_integrate, ctypes.c_void_p(self.d_seg_fm.data_ptr()), self._st())
        seg = self.comm.all_reduce_sum(self.d_seg_fm.cpu().numpy())
        st = ctypes.c_int(0)
        self._call(self.lib.lls_get_status, self._st(), ctypes.byref(st))
        flags = self.comm.all_reduce_sum(np.array([float((st.value >> b) & 1) for b in range(L.STATUS_BITS)]))
        status = sum((1 << b) for b in range(L.STATUS_BITS) if flags[b] > 0)
        if failed:
            status |= L.STATUS_NOT_CONVERGED
        torch.cuda.synchronize(self.device)
        if status & L.STATUS_COMM:
            raise self._native.LlsError("row-partitioned solve: a rank waited 30 s for a peer's panel (LLS_STATUS_COMM); "
                                        "a peer process died or the ranks fell out of step")
        return seg.reshape(self.n_segments, 12), it, err, status, hist

    def gamma(self):
        return self.d_out[L.O_GAMMA, :self.n].cpu().numpy()
