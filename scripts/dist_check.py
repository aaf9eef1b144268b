"""Partitioned (row-sharded) solve against the golden vectors of the reference.  Run with
    python -m torch.distributed.run --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py [fixtures...]
or as a plain python script (world size 1: every partitioned kernel runs, collectives are the identity)."""
import json, os, sys, time
import numpy as np, torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from golden_utils import load, segment_sums
from machupx_b200.partitioned import PartitionedDeviceScene

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
names = sys.argv[1:] or ["testplane_nonlinear", "traditional_linear", "swarm_nonlinear", "c2_traditional_n4000"]
ok = True
for name in names:
    tables, ac_state, g, meta = load(name)
    sp = meta["solver"]
    dev = PartitionedDeviceScene(tables)
    dev.upload_state(ac_state)
    t0 = time.perf_counter()
    seg, it, err, status, hist = dev.solve(sp["type"], sp["convergence"], sp["relaxation"], sp["max_iterations"])
    t1 = time.perf_counter()
    seg, it, err, status, hist = dev.solve(sp["type"], sp["convergence"], sp["relaxation"], sp["max_iterations"])
    torch.cuda.synchronize(); t2 = time.perf_counter()
    gam = dev.gamma()
    e_g = float(np.abs(gam - g["gamma"]).max() / np.abs(g["gamma"]).max())
    if "dM_inv" in g:
        ref = segment_sums(g, tables)
        e_f = float(np.abs(seg - ref).max() / np.abs(ref).max())
    else:       # lean full-size fixture: circulation only
        e_f = 0.0
    good = e_g < 1e-10 and e_f < 1e-10 and it == int(meta["iterations"])
    ok = ok and good
    if rank == 0:
        print(json.dumps({"fixture": name, "world": world, "n": int(tables["n"]), "ld": dev.ld, "gamma_rel_err": e_g, "force_rel_err": e_f,
                          "iterations": it, "reference_iterations": int(meta["iterations"]), "final_error": err, "status": status,
                          "solve_s": t2 - t1, "first_solve_s": t1 - t0, "fused_exchange": bool(dev.fused_exchange), "multicast": bool(getattr(dev, "multicast", False)), "exchange_error": getattr(dev, "exchange_error", None),
                          "exchange_stats": dev.exchange_stats() if dev.fused_exchange else None, "ok": bool(good)}), flush=True)
    del dev
if world > 1:
    dist.destroy_process_group()
sys.exit(0 if ok else 1)
