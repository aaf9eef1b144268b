#!/usr/bin/env bash
# Round-1 GPU pass A: parity tests, bench, ncu launch list, ncu --set full of the streaming kernels.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/a_tests.log
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/a_bench_plain.json 2> gpurun_out/a_bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/a_launches.csv $CMD > gpurun_out/a_ncu1.log 2>&1
$CMD > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'influence_kernel|assemble_kernel|residual_kernel' -c 4 -o gpurun_out/a_stream $CMD > gpurun_out/a_ncu2.log 2>&1
python bench.py --steps 3 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
echo done
