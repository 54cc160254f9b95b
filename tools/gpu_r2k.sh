#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
GRID_B200_CREATE_TIMING=1 timeout 300 python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-reference-gpu 2>&1 >/dev/null | grep "grid_b200 create" | head -50
