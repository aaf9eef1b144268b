#!/usr/bin/env bash
# round 2, call F (1 GPU): pipelined publication of the diagonal block
set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_scene_api_gpu.py tests/test_batch_gpu.py -q -x > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2f_tests.log
tail -6 gpurun_out/r2f_tests.log
for pipe in 1 0; do
  for n in 200 1000 2000 4000 8000 16000; do echo -n "pipe=$pipe " >> gpurun_out/r2f_lu.jsonl; LLS_LU_PIPE=$pipe timeout 120 python scripts/lu_bench.py $n 7 >> gpurun_out/r2f_lu.jsonl 2>> gpurun_out/r2f_lu.err; done
done
cat gpurun_out/r2f_lu.jsonl
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2f_bench.json')); print(d['value'], d['e2e']['value'], d['phase_ms'], d['parity'])"
