#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/g_plain.json 2> gpurun_out/g_plain.err &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:influence_kernel -s 2 -c 1 -o gpurun_out/g_influence $CMD > gpurun_out/g_ncu.log 2>&1
echo done
