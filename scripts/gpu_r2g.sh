#!/usr/bin/env bash
# round 2, call G (1 GPU): one-CTA-per-case LU for large batches of small systems (C3)
set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_batch_gpu.py tests/test_scene_api_gpu.py tests/test_gpu_golden.py tests/test_large_scene_gpu.py -q -x > gpurun_out/r2g_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2g_tests.log
tail -6 gpurun_out/r2g_tests.log
for mode in 16 0; do
  LLS_LU_SMALL_BATCH_MIN=$mode timeout 300 python bench.py --workload c3 --steps 3 --warmup 3 > gpurun_out/r2g_c3_small$mode.json 2> gpurun_out/r2g_c3_small$mode.err; echo "c3 small=$mode rc=$?"
done
timeout 300 python bench.py --workload c3 --c3-solver linear --steps 3 --warmup 3 > gpurun_out/r2g_c3_linear.json 2> gpurun_out/r2g_c3_linear.err; echo "c3 linear rc=$?"
for f in gpurun_out/r2g_c3_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches'], d['parity'])"; done
tail -3 gpurun_out/r2g_c3_small16.err
for sp in 1 0; do echo -n "INF_SPLIT=$sp " >> gpurun_out/r2g_influence.txt; INF_SPLIT=$sp timeout 120 python scripts/influence_bench.py >> gpurun_out/r2g_influence.txt 2>&1; done
cat gpurun_out/r2g_influence.txt
