#!/usr/bin/env bash
# 2 GPUs: symmetric-memory mailboxes + NVLS multicast stores
set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 2 --master-port 29801 scripts/dist_check.py testplane_nonlinear swarm_nonlinear c2_traditional_n4000 > gpurun_out/r2k_dist.jsonl 2> gpurun_out/r2k_dist.err; echo "dist_check rc=$?"
python - <<'PY'
import json
for line in open('gpurun_out/r2k_dist.jsonl'):
    if line.startswith('{'):
        d=json.loads(line); print({k:d.get(k) for k in ('fixture','ok','gamma_rel_err','iterations','fused_exchange','multicast','exchange_error','solve_s')})
PY
tail -5 gpurun_out/r2k_dist.err
for mc in 1 0; do
LLS_PARTITIONED_MULTICAST=$mc timeout 600 $TR --nproc-per-node 2 --master-port 2981$mc bench.py --gpus 2 --steps 3 --warmup 3 --swarm-aircraft 16 > gpurun_out/r2k_bench_mc$mc.json 2> gpurun_out/r2k_bench_mc$mc.err; echo "bench mc=$mc rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r2k_bench_mc$mc.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['phase_ms']['lu'], d['roofline']['achieved'], d['exchange']['multicast'], d['exchange']['fallback_reason'], d['exchange']['per_rank_wait_ms_per_step'], d['parity']['CL_sum'])"
done
timeout 900 python -m pytest tests/test_partitioned_gpu.py -q -x -k "multi_gpu" > gpurun_out/r2k_tests.log 2>&1; tail -4 gpurun_out/r2k_tests.log
