#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_golden.py tests/test_batch_gpu.py tests/test_partitioned_gpu.py -m gpu -x -q > gpurun_out/h_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/h_tests.log
rm -f gpurun_out/h_lu.jsonl
for n in 1000 4000 8000 16000; do timeout 120 python scripts/lu_bench.py $n 5 >> gpurun_out/h_lu.jsonl 2>> gpurun_out/h_lu.err; done
for n in 4000 16000; do LLS_LU_PAIRS=0 timeout 120 python scripts/lu_bench.py $n 5 >> gpurun_out/h_lu.jsonl 2>> gpurun_out/h_lu.err; done
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err
echo done
