#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 90 python tools/dbg/dbg_integrate.py 2>&1 | tail -30 || { echo "DEBUG SCRIPT FAILED/HUNG"; exit 1; }
timeout 400 python -m pytest tests/test_b200_parity.py -x -q -m gpu -k "ctile" --timeout 60 -p no:cacheprovider 2>&1 | tail -15
echo "=== bench variant 0 (ctile)"
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-reference-gpu --variant 3 2>gpurun_out/bench_v0.err > gpurun_out/bench_v0.json
python -c "
import json,sys; d=json.loads(open('gpurun_out/bench_v0.json').read()); print(d['ms_per_step'], d['roofline']['phase_ms_per_step'], d['e2e']['ms_per_step'], d['config']['task_block_pairs'])"
tail -3 gpurun_out/bench_v0.err
echo "=== ncu ctile (H2O-64)"
timeout 400 ncu --set full --import-source on --clock-control none -k regex:ctile_kernel -c 2 -o gpurun_out/ncu_ctile_h2o64 -f \
  python bench.py --workload H2O-64 --steps 1 --warmup 1 --no-cpu-baseline --no-reference-gpu --variant 3 > gpurun_out/ncu_ctile.log 2>&1
tail -2 gpurun_out/ncu_ctile.log | cut -c1-300
cat /tmp/grid_b200_abort.log 2>/dev/null | tail -5
