#!/usr/bin/env bash
# 8 GPUs: multicast vs unicast panel exchange at N = 32,000, and the 64k swarm
set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for mc in 1 0; do
LLS_PARTITIONED_MULTICAST=$mc timeout 600 $TR --nproc-per-node 8 --master-port 2991$mc bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/scale8c_bench_n8_mc$mc.json 2> gpurun_out/scale8c_bench_n8_mc$mc.err; echo "bench n=8 mc=$mc rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/scale8c_bench_n8_mc$mc.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['phase_ms']['lu'], d['roofline']['achieved'], d['exchange']['multicast'], d['exchange']['fallback_reason'], d['exchange']['per_rank_wait_ms_per_step'], d['parity']['CL_sum'])"
done
timeout 600 $TR --nproc-per-node 4 --master-port 29921 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/scale8c_bench_n4.json 2> gpurun_out/scale8c_bench_n4.err; echo "bench n=4 rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/scale8c_bench_n4.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['phase_ms']['lu'], d['roofline']['achieved'], d['exchange']['multicast'], d['exchange']['per_rank_wait_ms_per_step'], d['parity']['CL_sum'])"
timeout 900 $TR --nproc-per-node 8 --master-port 29931 scripts/swarm_scale.py 64 200 4 > gpurun_out/scale8c_c5_w8.json 2> gpurun_out/scale8c_c5_w8.err; echo "C5 rc=$?"; cut -c1-1000 gpurun_out/scale8c_c5_w8.json
tail -3 gpurun_out/scale8c_bench_n8_mc1.err
