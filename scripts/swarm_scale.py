"""Large swarm scenes, row-partitioned over the GPUs of one box (BASELINE.json configs[3] and [4]).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 --master-port 29512 \
        scripts/swarm_scale.py <n_aircraft> <grid_N_per_segment> [check_rows]

n_aircraft = 4  -> examples/Swarm as shipped (4 drones), e.g. grid 1000 -> 20,000 control points (configs[3])
n_aircraft = 64 -> the 8 x 8 lattice of SURVEY.md section 8d, grid 200 -> 64,000 control points (configs[4])
Prints one JSON line on rank 0.  Parity for sizes the reference cannot run: V_ji rows of a sample recomputed by the CPU
oracle, the converged residual norm, and agreement between different world sizes (compare the CL values of the runs)."""
import copy, json, os, sys, time
import numpy as np, torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from machupx_b200 import Scene
from machupx_b200.forces import assemble_fm
from machupx_b200.partitioned import PartitionedDeviceScene
from machupx_b200.tables import build_aircraft_state

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
n_ac = int(sys.argv[1]) if len(sys.argv) > 1 else 4
grid_n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
check_rows = int(sys.argv[3]) if len(sys.argv) > 3 else 0

from workloads import swarm_input
inp = swarm_input(n_ac, grid_n)
t0 = time.perf_counter(); scene = Scene(inp); t_host = time.perf_counter() - t0
dev = PartitionedDeviceScene(scene._tables)
dev.upload_state(build_aircraft_state(scene))
args = (scene._solver_type, scene._solver_convergence, scene._solver_relaxation, scene._max_solver_iterations)
torch.cuda.synchronize(); t0 = time.perf_counter()
seg, it, err, status, hist = dev.solve(*args)
torch.cuda.synchronize(); t_first = time.perf_counter() - t0
tm = None
if os.environ.get("LLS_TIMING"):
    import ctypes
    from machupx_b200 import _native
    _native.check(dev.lib.lls_set_timing(dev.handle, 1))
t_solve = None
if not os.environ.get("SWARM_ONE_SOLVE"):
    t0 = time.perf_counter()
    seg, it, err, status, hist = dev.solve(*args)
    torch.cuda.synchronize(); t_solve = time.perf_counter() - t0
if os.environ.get("LLS_TIMING"):
    ms = (ctypes.c_double * 6)(); _native.check(dev.lib.lls_get_timings(dev.handle, ms))
    tm = dict(zip(("influence", "assemble", "lu", "trisolve", "residual", "integrate"), list(ms)))
FM = assemble_fm(seg, scene._frames())
out = {"world": world, "control_points": scene._N, "ld": dev.ld, "aircraft": scene._num_aircraft, "host_setup_s": t_host,
       "first_solve_s": t_first, "steady_solve_s": t_solve, "newton_iterations": it, "final_residual": err, "status": status,
       "CL_first_last": [FM[k]["total"]["CL"] for k in (list(FM)[0], list(FM)[-1])],
       "CL_sum": float(sum(v["total"]["CL"] for v in FM.values())), "CD_sum": float(sum(v["total"]["CD"] for v in FM.values())),
       "gpu_mem_GB": torch.cuda.max_memory_allocated() / 1e9, "phase_ms": tm,
       "fused_exchange": bool(dev.fused_exchange), "multicast": bool(getattr(dev, "multicast", False)), "exchange_error": getattr(dev, "exchange_error", None), "exchange_stats": dev.exchange_stats() if dev.fused_exchange else None}
try:
    off = ctypes.c_size_t(0) if "ctypes" in dir() else None
    import ctypes as _ct
    off = _ct.c_size_t(0)
    dev.lib.lls_status_offset(dev.handle, _ct.byref(off))
    sc = dev.d_work[off.value - 256: off.value - 256 + 32].view(torch.float64).cpu().numpy()
    out["pivot_min_max"] = [float(sc[1]), float(sc[2])]
except Exception as e:
    out["pivot_min_max"] = repr(e)
if check_rows > 0:
    from oracle import lls_oracle as O      # checker only
    T = O.Tables(scene._tables, build_aircraft_state(scene)); F = O.flow_state(T)
    worst, scale = 0.0, 0.0
    mine = [g for g in np.unique(np.linspace(0, scene._N - 1, check_rows * world).astype(int)) if (g // 64) % world == rank][:check_rows]
    for gi in mine:
        Vo = O.influence_rows(T, F, int(gi), int(gi) + 1)[0]
        li = ((gi // 64) // world) * 64 + gi % 64
        Vg = dev.d_V[int(li), :, :scene._N].T.cpu().numpy()
        scale = max(scale, np.abs(Vo).max()); worst = max(worst, np.abs(Vg - Vo).max())
    out["V_rows_checked_rank0"] = len(mine); out["V_max_abs_err_over_scale_rank0"] = worst / max(scale, 1e-300)
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.destroy_process_group()
