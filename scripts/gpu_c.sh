#!/usr/bin/env bash
# GPU pass C: ncu evidence for the fused LU step kernel (launch list + full capture, large and small trailing matrix).
set -u
mkdir -p gpurun_out
python scripts/lu_bench.py 4000 2 > gpurun_out/c_plain4000.json 2>&1 &&
timeout 250 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:lu_ -c 200 --csv --log-file gpurun_out/c_lu_launches.csv python scripts/lu_bench.py 4000 2 > gpurun_out/c_ncu1.log 2>&1
python scripts/lu_bench.py 4000 2 > /dev/null 2>&1 &&
timeout 250 ncu --set full --clock-control none --import-source on -k regex:lu_step_kernel -s 70 -c 2 -o gpurun_out/c_lu_big python scripts/lu_bench.py 4000 2 > gpurun_out/c_ncu2.log 2>&1
python scripts/lu_bench.py 1000 2 > gpurun_out/c_plain1000.json 2>&1 &&
timeout 250 ncu --set full --clock-control none --import-source on -k regex:lu_step_kernel -s 20 -c 2 -o gpurun_out/c_lu_small python scripts/lu_bench.py 1000 2 > gpurun_out/c_ncu3.log 2>&1
echo done
