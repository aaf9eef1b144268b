#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
python scripts/lu_clocks.py 256 > gpurun_out/e_clocks.json 2>&1
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/e_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/e_tests.log
rm -f gpurun_out/e_lu.jsonl
for n in 1000 4000 8000; do timeout 120 python scripts/lu_bench.py $n 5 >> gpurun_out/e_lu.jsonl 2>> gpurun_out/e_lu.err; done
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err
echo done
