#!/usr/bin/env bash
# Multi-GPU evidence run: NG = number of GPUs of the box, AIRCRAFT = size of the strong-scaling swarm (x 1000 control points),
# NS = space-separated world sizes to bench.  Raw outputs land in gpurun_out/scale_*.
set -u
cd "$GRAFT_REPO_ROOT"
NG=${NG:-8}; AIRCRAFT=${AIRCRAFT:-32}; NS=${NS:-"8 4"}; STEPS=${STEPS:-5}; TAG=${TAG:-scale}; C5=${C5:-1}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node $NG --master-port 29701 scripts/dist_check.py testplane_nonlinear swarm_nonlinear > gpurun_out/${TAG}_dist_check_w$NG.jsonl 2> gpurun_out/${TAG}_dist_check_w$NG.err; echo "dist_check world $NG rc=$?"
cat gpurun_out/${TAG}_dist_check_w$NG.jsonl | cut -c1-260
for n in $NS; do
  if [ "$n" = "1" ]; then
    timeout 900 python bench.py --gpus 1 --steps $STEPS --warmup 3 --workload swarm --swarm-aircraft $AIRCRAFT > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench n=1 rc=$?"
  else
    timeout 900 $TR --nproc-per-node $n --master-port 2971$n bench.py --gpus $n --steps $STEPS --warmup 3 --swarm-aircraft $AIRCRAFT > gpurun_out/${TAG}_bench_n$n.json 2> gpurun_out/${TAG}_bench_n$n.err; echo "bench n=$n rc=$?"
  fi
  python - gpurun_out/${TAG}_bench_n$n.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","n_gpus","per_rank_newton_iterations")}, "e2e", d["e2e"]["value"], "lu_ms", d["phase_ms"]["lu"], "TF/GPU", d["roofline"]["achieved"], d["exchange"]["per_rank_wait_ms_per_step"], d["parity"]["CL_sum"], d["parity"]["final_residual"], "setup", d["host_setup_s"])
except Exception as e: print("ERR", e)
PY
done
if [ "$C5" = "1" ]; then
  timeout 900 $TR --nproc-per-node $NG --master-port 29731 scripts/swarm_scale.py 64 200 4 > gpurun_out/${TAG}_c5_w$NG.json 2> gpurun_out/${TAG}_c5_w$NG.err; echo "C5 world $NG rc=$?"; cut -c1-900 gpurun_out/${TAG}_c5_w$NG.json
fi
timeout 600 $TR --nproc-per-node $NG --master-port 29741 bench.py --gpus $NG --workload c3 --steps 3 --warmup 3 > gpurun_out/${TAG}_c3_n$NG.json 2> gpurun_out/${TAG}_c3_n$NG.err; echo "c3 n=$NG rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_c3_n$NG.json').read().strip().splitlines()[-1]); print('c3', d['value'], d['e2e']['value'], d['parity'])"
if [ "${LATTICE:-1}" = "1" ]; then
  timeout 900 python -m pytest tests/test_partitioned_gpu.py -q -x -k "lattice or multi_gpu" > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log; tail -5 gpurun_out/${TAG}_tests.log
fi
