#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for v in l3; do
export GRID_B200_LIB=$PWD/variants/libgrid_b200_$v.so
timeout 400 python -m pytest tests/test_b200_parity.py tests/test_b200_workloads.py -x -q -m gpu --timeout 200 -p no:cacheprovider -k "warptile or h2o_subset or h2o64_full" 2>&1 | tail -2
for extra in "--forces" "--virial" ""; do
timeout 200 python bench.py --steps 5 --warmup 2 --no-cpu-baseline --no-reference-gpu $extra 2>gpurun_out/b.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v $extra', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['roofline']['phase_ms_per_step'].items()})"
done
done
