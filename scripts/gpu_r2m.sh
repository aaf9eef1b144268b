#!/usr/bin/env bash
# final 1-GPU validation of the round: whole GPU suite, smoke, C2 bench (both arms in-line), C3 bench
set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2m_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2m_tests.log
tail -6 gpurun_out/r2m_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2m_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2m_smoke.log
timeout 300 python bench.py --workload c3 --steps 3 --warmup 3 > gpurun_out/r2m_c3.json 2> gpurun_out/r2m_c3.err
python -c "
import json; d=json.loads(open('gpurun_out/r2m_c3.json').read().strip().splitlines()[-1]); print('c3', d['value'], d['e2e']['value'], d['phase_ms'], d['parity'])"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r2m_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e'], d['roofline']['frac'], d['kernels']['vji_build']['ms'], d['cpu_baseline']['value'], d['cpu_baseline']['kind'], d['clocks'])"
