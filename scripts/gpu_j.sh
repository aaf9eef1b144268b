#!/usr/bin/env bash
# Round-1 closing evidence pass: all GPU tests, smoke, bench (both arms), ncu launch list of one full solve, ncu --set full of the
# LU block step and the influence kernel, LU size sweep, chain clocks, C3 sweep, C4 large scene.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/j_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/j_tests.log
python __graft_entry__.py smoke > gpurun_out/j_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/j_smoke.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err
timeout 200 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/j_bench_ref.json 2> gpurun_out/j_bench_ref.err
rm -f gpurun_out/j_lu.jsonl
for n in 1000 2000 4000 8000 16000; do timeout 120 python scripts/lu_bench.py $n 5 >> gpurun_out/j_lu.jsonl 2>> gpurun_out/j_lu.err; done
timeout 120 python scripts/lu_clocks.py 4000 > gpurun_out/j_clocks.txt 2>&1
timeout 300 python scripts/c3_sweep_bench.py > gpurun_out/j_c3.json 2> gpurun_out/j_c3.err
timeout 300 python scripts/c3_sweep_bench.py 64 64 linear >> gpurun_out/j_c3.json 2>> gpurun_out/j_c3.err
python scripts/influence_bench.py > gpurun_out/j_influence.json 2>&1
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/j_plain.json 2> gpurun_out/j_plain.err &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'lls::|lu_|influence|assemble|residual|integrate|flow_state|sum_kernel|newton|copy_gamma' -s 3500 -c 900 --csv --log-file gpurun_out/j_launches.csv $CMD > gpurun_out/j_ncu1.log 2>&1
$CMD > /dev/null 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:influence_kernel -s 2 -c 1 -o gpurun_out/j_influence $CMD > gpurun_out/j_ncu2.log 2>&1
$CMD > /dev/null 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lu_step_kernel -s 198 -c 1 -o gpurun_out/j_lu $CMD > gpurun_out/j_ncu3.log 2>&1
timeout 400 python scripts/large_scene.py > gpurun_out/j_c4.json 2> gpurun_out/j_c4.err
echo done
