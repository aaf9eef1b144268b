#!/usr/bin/env bash
# round 2, call C (2 GPUs): peer-memory exchange of the partitioned LU
set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2c_topo.txt 2>&1
timeout 900 python -m pytest tests/test_partitioned_gpu.py -q -x > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2c_tests.log
tail -25 gpurun_out/r2c_tests.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for mode in 1 0; do
  LLS_PARTITIONED_P2P=$mode timeout 600 $TR --nproc-per-node 2 --master-port 29601 bench.py --gpus 2 --steps 3 --warmup 3 --swarm-aircraft 16 > gpurun_out/r2c_bench_n2_p2p$mode.json 2> gpurun_out/r2c_bench_n2_p2p$mode.err; echo "bench n2 p2p=$mode rc=$?"
done
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --workload swarm --swarm-aircraft 16 > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err; echo "bench n1 rc=$?"
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --workload swarm --swarm-aircraft 16 --force-partitioned > gpurun_out/r2c_bench_n1_part.json 2> gpurun_out/r2c_bench_n1_part.err; echo "bench n1 partitioned rc=$?"
for f in gpurun_out/r2c_bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","n_gpus","phase_ms","per_rank_newton_iterations")}, d["e2e"]["value"], d["exchange"]["per_rank_wait_ms_per_step"], d["parity"]["CL_sum"], d["roofline"]["achieved"])
except Exception as e: print("ERR", e)
PY
done
tail -5 gpurun_out/r2c_bench_n2_p2p1.err
