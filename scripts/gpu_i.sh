#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/i_lu.jsonl gpurun_out/i_lu.err
run() { echo "# $*" >> gpurun_out/i_lu.jsonl; env "$@" timeout 120 python scripts/lu_bench.py $N 4 >> gpurun_out/i_lu.jsonl 2>> gpurun_out/i_lu.err; }
for N in 200 1000 4000 8000 16000; do
run LLS_LU_PDL=0
run LLS_LU_PDL=1
done
cat gpurun_out/i_lu.jsonl; tail -5 gpurun_out/i_lu.err
timeout 600 python -m pytest tests/test_gpu_golden.py tests/test_batch_gpu.py -m gpu -x -q > gpurun_out/i_tests.log 2>&1; tail -3 gpurun_out/i_tests.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err; python -c "
import json; d=json.load(open('gpurun_out/i_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['achieved'], d['phase_ms'], d['parity'])"
