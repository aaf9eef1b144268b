#!/usr/bin/env bash
# Round-1 evidence pass: all GPU tests, smoke, bench, ncu launch list of one full solve, ncu --set full of the top kernels.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/f_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/f_tests.log
python __graft_entry__.py smoke > gpurun_out/f_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/f_smoke.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
timeout 200 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/f_plain.json 2> gpurun_out/f_plain.err &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'lls::|lu_|influence|assemble|residual|integrate|flow_state|sum_kernel|newton|copy_gamma' -s 3500 -c 900 --csv --log-file gpurun_out/f_launches.csv $CMD > gpurun_out/f_ncu1.log 2>&1
$CMD > /dev/null 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'assemble_kernel|residual_kernel' -s 6 -c 2 -o gpurun_out/f_stream $CMD > gpurun_out/f_ncu2.log 2>&1
$CMD > /dev/null 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:influence_kernel -s 2 -c 1 -o gpurun_out/f_influence $CMD > gpurun_out/f_ncu4.log 2>&1
$CMD > /dev/null 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lu_step_kernel -s 198 -c 1 -o gpurun_out/f_lu $CMD > gpurun_out/f_ncu3.log 2>&1
echo done
