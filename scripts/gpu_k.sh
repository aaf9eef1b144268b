#!/usr/bin/env bash
# final check of the committed build: all GPU tests, smoke, bench line, LU size sweep
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/k_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/k_tests.log
python __graft_entry__.py smoke > gpurun_out/k_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/k_smoke.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err
rm -f gpurun_out/k_lu.jsonl
for n in 200 1000 2000 4000 8000 16000; do timeout 120 python scripts/lu_bench.py $n 5 >> gpurun_out/k_lu.jsonl 2>> gpurun_out/k_lu.err; done
timeout 120 python scripts/lu_clocks.py 4000 > gpurun_out/k_clocks.txt 2>&1
tail -2 gpurun_out/k_tests.log; tail -1 gpurun_out/k_smoke.log; cat gpurun_out/k_lu.jsonl
python -c "
import json; d=json.load(open('gpurun_out/k_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['phase_ms'], d['cpu_baseline']['value'])"
