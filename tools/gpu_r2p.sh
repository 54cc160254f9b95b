#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_b200_parity.py tests/test_b200_workloads.py tests/test_b200_dispatcher.py -x -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | tail -3
GRID_B200_CREATE_TIMING=1 timeout 300 python bench.py --steps 5 --warmup 2 --no-cpu-baseline --no-reference-gpu 2>gpurun_out/c.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), d['create_task_list'])"
grep "grid_b200 create: [a-z]" gpurun_out/c.err | head -8
GRID_B200_CREATE_TIMING=1 timeout 300 python bench.py --workload H2O-1024 --steps 2 --warmup 1 --no-cpu-baseline --no-reference-gpu 2>gpurun_out/c2.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('H2O-1024', round(d['ms_per_step'],3), d['create_task_list'])"
grep "grid_b200 create: [a-z]" gpurun_out/c2.err | head -8
