#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in 0 1; do LLS_LU_BS_SENTINEL=$v timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/i_bench.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('sentinel=$v', d['value'], d['e2e']['value'], d['phase_ms']['lu'], d['parity'])"; done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:backsolve -c 4 --csv --log-file gpurun_out/i_bs.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1; tail -3 gpurun_out/i_bs.csv | cut -c150-400
timeout 200 python scripts/c3_sweep_bench.py | cut -c1-250
