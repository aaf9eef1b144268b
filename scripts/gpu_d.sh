#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/d_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/d_tests.log
rm -f gpurun_out/d_lu.jsonl
for n in 1000 4000 8000; do timeout 120 python scripts/lu_bench.py $n 5 >> gpurun_out/d_lu.jsonl 2>> gpurun_out/d_lu.err; done
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err
python scripts/lu_bench.py 4000 2 > /dev/null 2>&1 &&
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:lu_ -c 140 --csv --log-file gpurun_out/d_lu_launches.csv python scripts/lu_bench.py 4000 2 > gpurun_out/d_ncu1.log 2>&1
echo done
