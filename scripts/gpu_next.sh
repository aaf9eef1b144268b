#!/usr/bin/env bash
# What the end of round 2 left unmeasured on a GPU (the round's GPU minutes were spent before these changes were written; all of
# them are covered on CPU).  One B200, about 6 minutes:  gpurun --timeout 600 -- 'bash scripts/gpu_next.sh'
set -uo pipefail
mkdir -p gpurun_out
# 1. the fuzzer's single cases on the CUDA path (18 finite cases; opt-in because never run on a device)
LLS_FUZZ_ON_GPU=1 timeout 200 python -m pytest tests/test_fuzz_cases.py -m gpu -q 2>&1 | tail -15 > gpurun_out/next_fuzz_gpu.log
# 2. C3 end to end after the column-wise sweep-state detection (host part 20.7 -> 5.3 ms per 4,096-state sweep on the build host)
timeout 200 python bench.py --workload c3 --steps 10 --warmup 3 > gpurun_out/next_bench_c3.json 2> gpurun_out/next_bench_c3.err
# 3. database airfoils: memory checker over the three device tests, then the launch list of one database solve
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_database_airfoils.py -m gpu -q -x \
  > gpurun_out/next_db_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/next_db_memcheck.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/next_db_launches.csv \
  python -m pytest tests/test_database_airfoils.py -m gpu -q -k golden_on_the_device > gpurun_out/next_db_ncu.log 2>&1
tail -3 gpurun_out/next_fuzz_gpu.log gpurun_out/next_db_memcheck.log; cut -c1-300 gpurun_out/next_bench_c3.json
