#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/i_lu.jsonl
run() { echo "# $*" >> gpurun_out/i_lu.jsonl; env "$@" timeout 120 python scripts/lu_bench.py $N 5 >> gpurun_out/i_lu.jsonl 2>> gpurun_out/i_lu.err; }
for N in 4000 6000 8000; do
for m in 64 56 48 40 80 96; do run LLS_LU_PAIR_MIN_REM=$m; done
done
cat gpurun_out/i_lu.jsonl
