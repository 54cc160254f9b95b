#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_b200_parity.py tests/test_b200_workloads.py -x -q -m gpu --timeout 300 -p no:cacheprovider -k "warptile or h2o_subset or h2o64_full" 2>&1 | tail -4
for g8 in 0 1; do
GRID_B200_COEF_G8=$g8 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('G8=$g8', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['roofline']['phase_ms_per_step'].items()})"
done
