#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
export GRID_B200_LIB=$PWD/variants/libgrid_b200_ct1.so
timeout 300 python -m pytest tests/test_b200_parity.py -x -q -m gpu --timeout 60 -p no:cacheprovider -k "ctile and golden" 2>&1 | tail -3
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-reference-gpu --variant 3 2>gpurun_out/b.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ct1', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['roofline']['phase_ms_per_step'].items()})"
tail -2 gpurun_out/b.err
