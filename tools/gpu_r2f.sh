#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 120 python tools/dbg/dbg_refgpu.py 2>&1 | tail -12
echo "=== ncu full: warp-tile class-0 kernels, H2O-256"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"tiled_kernel<.*0, 2" -c 8 -o gpurun_out/ncu_tiled_h2o256_r02 -f \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-reference-gpu > gpurun_out/ncu_tiled.log 2>&1
tail -2 gpurun_out/ncu_tiled.log | cut -c1-200
echo "=== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_h2o256_r02.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-reference-gpu > gpurun_out/launches.log 2>&1
tail -1 gpurun_out/launches.log | cut -c1-200
ls -la gpurun_out | tail -8
