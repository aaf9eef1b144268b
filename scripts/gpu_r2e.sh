#!/usr/bin/env bash
set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
SWARM_ONE_SOLVE=1 timeout 300 python scripts/swarm_scale.py 8 200 4 > gpurun_out/r2e_swarm.json 2> gpurun_out/r2e_swarm.err; echo "rc=$?"; cat gpurun_out/r2e_swarm.json; tail -3 gpurun_out/r2e_swarm.err
timeout 300 python scripts/swarm_scale.py 8 200 0 > gpurun_out/r2e_swarm2.json 2> gpurun_out/r2e_swarm2.err; echo "rc=$?"; cat gpurun_out/r2e_swarm2.json
