#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for v in c5; do
export GRID_B200_LIB=$PWD/variants/libgrid_b200_$v.so
timeout 300 python -m pytest tests/test_b200_parity.py -x -q -m gpu --timeout 60 -p no:cacheprovider -k "warptile and (golden or multi_pair)" 2>&1 | tail -2
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu 2>gpurun_out/b.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['roofline']['phase_ms_per_step'].items()})"
done
