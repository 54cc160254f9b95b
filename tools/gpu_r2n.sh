#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -x -q -m gpu --timeout 900 -p no:cacheprovider ) 2>&1 | tail -8
timeout 900 python tools/bench_sample_tasks.py > gpurun_out/sample_tasks_r02.json 2> gpurun_out/sample_tasks.err; tail -14 gpurun_out/sample_tasks.err | cut -c1-330
timeout 1500 bash tools/profile_round.sh r02 2>&1 | tail -5
