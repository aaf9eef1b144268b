#!/usr/bin/env bash
set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_partitioned_gpu.py tests/test_batch_gpu.py tests/test_scene_api_gpu.py -q -x > gpurun_out/r2i_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2i_tests.log
tail -5 gpurun_out/r2i_tests.log
timeout 300 python bench.py --workload c3 --steps 3 --warmup 3 > gpurun_out/r2i_c3.json 2> gpurun_out/r2i_c3.err
python -c "
import json; d=json.loads(open('gpurun_out/r2i_c3.json').read().strip().splitlines()[-1]); print('c3', d['value'], d['e2e']['value'], d['phase_ms'], d['parity'])"
timeout 300 python bench.py --gpus 1 --steps 3 --warmup 3 --workload swarm --swarm-aircraft 16 --force-partitioned > gpurun_out/r2i_part16.json 2> gpurun_out/r2i_part16.err
python -c "
import json; d=json.loads(open('gpurun_out/r2i_part16.json').read().strip().splitlines()[-1]); print('part16', d['ms_per_step'], d['phase_ms']['lu'], d['roofline']['achieved'], d['parity']['CL_sum'])"
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -s 3300 -c 1100 --csv --log-file gpurun_out/p2_launches.csv $CMD > gpurun_out/p2_ncu0.log 2>&1; echo "launch list rc=$?"; grep -c lu_step gpurun_out/p2_launches.csv
