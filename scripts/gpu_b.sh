#!/usr/bin/env bash
# GPU pass B: parity tests with the fused LU, LU micro-bench, bench line.
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/b_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/b_tests.log
for n in 1000 4000 8000 16000; do timeout 120 python scripts/lu_bench.py $n 5 >> gpurun_out/b_lu.jsonl 2>> gpurun_out/b_lu.err; done
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err
python scripts/probe_peaks.py > gpurun_out/b_peaks.json 2>&1
echo done
