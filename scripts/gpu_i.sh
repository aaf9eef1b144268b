#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/i_lu.jsonl gpurun_out/i_lu.err
run() { echo "# $*" >> gpurun_out/i_lu.jsonl; env "$@" timeout 120 python scripts/lu_bench.py $N 5 >> gpurun_out/i_lu.jsonl 2>> gpurun_out/i_lu.err; }
for N in 2000 4000 8000; do
run LLS_LU_PANEL_GUARD=1,0
run LLS_LU_PANEL_GUARD=20,46
run LLS_LU_PANEL_GUARD=12,46
run LLS_LU_PANEL_GUARD=20,62
run LLS_LU_PANEL_GUARD=30,46
run LLS_LU_PANEL_GUARD=2,36
done
cat gpurun_out/i_lu.jsonl; tail -5 gpurun_out/i_lu.err
timeout 120 python scripts/lu_clocks.py 4000 > gpurun_out/i_clocks_pg.txt 2>&1; sed -n '10p;20p;30p;40p' gpurun_out/i_clocks_pg.txt | cut -c1-230
timeout 300 python -m pytest tests/test_gpu_golden.py -m gpu -x -q -k "lu or c2" 2>&1 | tail -3
