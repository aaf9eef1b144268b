#!/usr/bin/env bash
# refresh of the round-2 captures after the last kernel changes (1 GPU)
set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
P16="python bench.py --gpus 1 --steps 1 --warmup 3 --workload swarm --swarm-aircraft 16 --force-partitioned"
$P16 > gpurun_out/p3_part16.json 2> gpurun_out/p3_part16.err && \
ncu --set full --clock-control none --import-source on -k regex:lu_dist_step_kernel -s 262 -c 2 -o gpurun_out/p3_dist16 -f $P16 > gpurun_out/p3_ncu1.log 2>&1; echo "dist16 rc=$?"
C3="python bench.py --workload c3 --steps 1 --warmup 3"
$C3 > gpurun_out/p3_c3.json 2> gpurun_out/p3_c3.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv --log-file gpurun_out/p3_c3_launches.csv $C3 > gpurun_out/p3_ncu2.log 2>&1; echo "c3 launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:residual -s 12 -c 2 -o gpurun_out/p3_residual -f $C3 > gpurun_out/p3_ncu3.log 2>&1; echo "residual rc=$?"
ls -la gpurun_out/p3_*
