#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 300 python tools/dbg/dbg_integrate.py 2>&1 | tail -40
echo "=== ncu ctile collocate (H2O-64)"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:ctile_kernel -c 2 -o gpurun_out/ncu_ctile_h2o64 -f \
  python bench.py --workload H2O-64 --steps 1 --warmup 1 --no-cpu-baseline --no-reference-gpu > gpurun_out/ncu_ctile.log 2>&1
tail -3 gpurun_out/ncu_ctile.log
ls -la gpurun_out/
