#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/i_lu.jsonl gpurun_out/i_lu.err
run() { echo "# $*" >> gpurun_out/i_lu.jsonl; env "$@" timeout 120 python scripts/lu_bench.py $N 5 >> gpurun_out/i_lu.jsonl 2>> gpurun_out/i_lu.err; }
for N in 64 200 1000 2000 4000; do
run LLS_LU_SPLIT4_MAX_REM=0
run LLS_LU_SPLIT4_MAX_REM=8
run LLS_LU_SPLIT4_MAX_REM=16
run LLS_LU_SPLIT4_MAX_REM=36
done
cat gpurun_out/i_lu.jsonl; tail -5 gpurun_out/i_lu.err
timeout 600 python -m pytest tests/test_gpu_golden.py tests/test_scene_api_gpu.py -m gpu -x -q 2>&1 | tail -3
