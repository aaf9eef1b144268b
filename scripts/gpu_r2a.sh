#!/usr/bin/env bash
# round 2, call A (1 GPU): parity tests, C2 bench (both arms), LU size sweep
set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2a_tests.log
tail -5 gpurun_out/r2a_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
for n in 200 1000 2000 4000 8000 16000; do timeout 120 python scripts/lu_bench.py $n 5 >> gpurun_out/r2a_lu.jsonl 2>> gpurun_out/r2a_lu.err; done
cat gpurun_out/r2a_lu.jsonl
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err; echo "ref rc=$?"
nproc; free -g | head -2
