"""Flight-state sweeps sharded by case over the GPUs of one box (BASELINE.json configs[2]).

Every case (one flight state of an alpha / beta / rate / control sweep, or one stencil point of a finite-difference
driver) is a complete, independent ``solve_forces``: the geometry tables are identical, only the 12-double aircraft
state block and the flap rows differ.  So the sweep shards with NO data-path collective: rank r owns a contiguous block
of cases, solves them on its own GPU, and one ``all_gather`` of the per-case totals at the end gives every rank the full
table.  The reference's idiom for the same job is a serial Python loop of ``set_aircraft_state`` / ``solve_forces``
(or a ``multiprocessing.Pool`` over cases, dev/studies/blending_distance_study.py:184).
"""
import numpy as np

TOTAL_KEYS = ("CL", "CD", "CS", "Cx", "Cy", "Cz", "Cl", "Cm", "Cn", "Cl_w", "Cm_w", "Cn_w",
              "FL", "FD", "FS", "Fx", "Fy", "Fz", "Mx", "My", "Mz", "Mx_w", "My_w", "Mz_w")


def shard_cases(n_cases, world_size, rank):
    """Contiguous block of cases owned by ``rank``: the first ``n_cases % world_size`` ranks get one extra case."""
    base, extra = divmod(int(n_cases), int(world_size))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def alpha_beta_grid(alphas, betas, velocity, **fixed):
    """Cartesian alpha x beta grid of state dictionaries (alpha varies slowest), e.g. the 64 x 64 grid of configs[2]."""
    return [dict(fixed, velocity=velocity, alpha=float(a), beta=float(b)) for a in alphas for b in betas]


class SceneCaseSolver:
    """Default per-case solver: one ``Scene`` kept resident on this rank's GPU, state swapped per case."""

    def __init__(self, scene_input, aircraft=None, control_states=None, solve_kwargs=None):
        from .scene import Scene
        self.scene = Scene(scene_input)
        self.aircraft = aircraft if aircraft is not None else list(self.scene._airplanes.keys())[0]
        self.control_states = control_states
        self.solve_kwargs = dict(solve_kwargs or {})

    def __call__(self, case_index, state):
        from .exceptions import SolverNotConvergedError
        self.scene.set_aircraft_state(state, aircraft=self.aircraft)
        if self.control_states is not None:
            self.scene.set_aircraft_control_state(self.control_states[case_index], aircraft=self.aircraft)
        try:
            FM = self.scene.solve_forces(**self.solve_kwargs)
        except SolverNotConvergedError:
            return None
        tot = FM[self.aircraft]["total"]
        return np.array([tot.get(k, np.nan) for k in TOTAL_KEYS]), self.scene._last_iterations


class BatchSceneSolver:
    """Block solver: the rank's whole shard of cases goes through ``Scene.solve_forces_batch`` -- one batched device pass
    (case index in blockIdx.z of every kernel) per ``max_batch`` cases instead of one solve per case.  This is the path
    BASELINE.json configs[2] is measured on."""

    def __init__(self, scene_input, aircraft=None, max_batch=4096):
        from .scene import Scene
        self.scene = Scene(scene_input)
        self.scene.set_err_state(not_converged="ignore")       # non-convergence is reported per case, raised by run_sweep
        self.aircraft = aircraft if aircraft is not None else list(self.scene._airplanes.keys())[0]
        self.max_batch = int(max_batch)

    def solve_block(self, first_case, states):
        """-> (values (m, len(TOTAL_KEYS)), newton iterations (m,), converged (m,)) for the m states of this rank's shard."""
        if len(states) == 0:
            return np.zeros((0, len(TOTAL_KEYS))), np.zeros(0, dtype=int), np.zeros(0, dtype=bool)
        res = self.scene.solve_forces_batch(states, aircraft=self.aircraft, max_batch=self.max_batch)
        tot = res["totals"][self.aircraft]
        values = np.stack([np.asarray(tot[k], dtype=float) for k in TOTAL_KEYS], axis=1)
        return values, np.asarray(res["iterations"], dtype=int), np.asarray(res["converged"], dtype=bool)


def run_sweep(states, solver, group=None, on_not_converged="raise"):
    """Solves ``states`` (list of state dicts) sharded over the ranks of ``group`` (``torch.distributed``; None = single
    process).  A solver with a ``solve_block(first_case, states)`` method (``BatchSceneSolver``) gets its rank's whole shard at
    once; otherwise ``solver(case_index, state)`` returns ``(values[len(TOTAL_KEYS)], newton_iterations)`` or None if the case
    did not converge.  Returns ``(table (n_cases, len(TOTAL_KEYS)), iterations (n_cases,), converged (n_cases,) bool)`` on
    every rank.  ``on_not_converged``: "raise" (the reference's default error policy; raised on every rank after the
    gather so no rank is left waiting in a collective) or "nan"."""
    import torch
    import torch.distributed as dist
    n = len(states)
    distributed = group is not None or (dist.is_available() and dist.is_initialized())
    world = dist.get_world_size(group) if distributed else 1
    rank = dist.get_rank(group) if distributed else 0
    lo, hi = shard_cases(n, world, rank)
    width = len(TOTAL_KEYS) + 2                       # values, iterations, converged flag
    local = np.full((hi - lo, width), np.nan)
    if hasattr(solver, "solve_block"):                 # the rank's shard as batched device passes
        values, its, ok = solver.solve_block(lo, states[lo:hi])
        local[:, :len(TOTAL_KEYS)] = values
        local[:, -2], local[:, -1] = its, ok
        local[~np.asarray(ok, dtype=bool), :len(TOTAL_KEYS)] = np.nan
    else:
        for i in range(lo, hi):
            res = solver(i, states[i])
            if res is None:
                local[i - lo, -2:] = (0.0, 0.0)
            else:
                local[i - lo, :len(TOTAL_KEYS)] = res[0]
                local[i - lo, -2:] = (float(res[1]), 1.0)
    if distributed:
        backend = dist.get_backend(group)
        dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        # equal-sized slots (the largest shard) so a plain all_gather works; the tail of short shards is padding
        slot = -(-n // world)
        send = torch.full((slot, width), float("nan"), dtype=torch.float64, device=dev)
        send[:hi - lo] = torch.from_numpy(local).to(dev)
        recv = [torch.empty_like(send) for _ in range(world)]
        dist.all_gather(recv, send, group=group)
        parts = []
        for r, t in enumerate(recv):
            a, b = shard_cases(n, world, r)
            parts.append(t[:b - a].cpu().numpy())
        full = np.concatenate(parts, axis=0) if parts else local
    else:
        full = local
    table, iterations, converged = full[:, :len(TOTAL_KEYS)], full[:, -2].astype(int), full[:, -1] > 0.5
    if on_not_converged == "raise" and not converged.all():
        from .exceptions import SolverNotConvergedError
        bad = int(np.argmin(converged))
        raise SolverNotConvergedError("nonlinear", float("nan")) from RuntimeError("case %d of the sweep did not converge" % bad)
    return table, iterations, converged
