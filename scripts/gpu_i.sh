#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/i_lu.jsonl gpurun_out/i_lu.err
run() { echo "# $*" >> gpurun_out/i_lu.jsonl; env "$@" timeout 120 python scripts/lu_bench.py $N 5 >> gpurun_out/i_lu.jsonl 2>> gpurun_out/i_lu.err; }
for N in 64 200 1000 2000 4000 8000 16000; do
run LLS_LU_INV_CTA=0
run LLS_LU_INV_CTA=1
done
cat gpurun_out/i_lu.jsonl; tail -5 gpurun_out/i_lu.err
timeout 120 python scripts/lu_clocks.py 4000 > gpurun_out/i_clocks_inv.txt 2>&1; sed -n '10p;20p;30p;40p;55p' gpurun_out/i_clocks_inv.txt | cut -c1-230
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err; python -c "
import json; d=json.load(open('gpurun_out/i_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['phase_ms'], d['parity'])"
