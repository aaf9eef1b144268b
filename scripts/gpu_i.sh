#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/i_lu.jsonl gpurun_out/i_lu.err
run() { echo "# $*" >> gpurun_out/i_lu.jsonl; env "$@" timeout 120 python scripts/lu_bench.py $N 3 >> gpurun_out/i_lu.jsonl 2>> gpurun_out/i_lu.err; }
N=8000
for d in 0 1 2 3 4 5 6 7; do run LLS_LU_WORKER_DBG=$d LLS_LU_WORKER_CHUNK=4; done
cat gpurun_out/i_lu.jsonl; tail -5 gpurun_out/i_lu.err
