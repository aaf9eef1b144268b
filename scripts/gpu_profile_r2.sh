#!/usr/bin/env bash
# round 2 evidence (1 GPU): launch list of one C2 solve, ncu --set full captures of the dominant / new kernels, C3 phase times
set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/p2_plain.json 2> gpurun_out/p2_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:lls -s 3300 -c 1100 --csv --log-file gpurun_out/p2_launches.csv $CMD > gpurun_out/p2_ncu0.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:lu_step_kernel -s 198 -c 1 -o gpurun_out/p2_lu -f $CMD > gpurun_out/p2_ncu1.log 2>&1; echo "lu rc=$?"
ncu --set full --clock-control none --import-source on -k regex:influence_kernel -s 4 -c 2 -o gpurun_out/p2_influence -f $CMD > gpurun_out/p2_ncu2.log 2>&1; echo "influence rc=$?"
DC="python scripts/dist_check.py c2_traditional_n4000"
$DC > gpurun_out/p2_dist_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lu_dist_step_kernel -s 72 -c 1 -o gpurun_out/p2_dist -f $DC > gpurun_out/p2_ncu3.log 2>&1; echo "dist rc=$?"
C3="python bench.py --workload c3 --steps 1 --warmup 3"
$C3 > gpurun_out/p2_c3.json 2> gpurun_out/p2_c3.err && \
ncu --set full --clock-control none --import-source on -k regex:lu_small_case_kernel -s 30 -c 1 -o gpurun_out/p2_small -f $C3 > gpurun_out/p2_ncu4.log 2>&1; echo "small rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/p2_c3.json').read().strip().splitlines()[-1]); print('c3', d['value'], d['phase_ms'])"
ls -la gpurun_out/p2_*
