#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
for N in 64 200 1000 4000 8000 16000; do timeout 100 python scripts/lu_bench.py $N 5; done
timeout 120 python scripts/lu_clocks.py 4000 > gpurun_out/i_clocks_la.txt 2>&1; sed -n '2p;10p;30p;55p' gpurun_out/i_clocks_la.txt | cut -c1-230
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/i_tests.log 2>&1; tail -3 gpurun_out/i_tests.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err; python -c "
import json; d=json.load(open('gpurun_out/i_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['phase_ms'], d['parity'])"
