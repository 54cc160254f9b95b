#!/bin/bash
# round 2: full-size parity (H2O-1024, config 4), reference GPU backend comparator, a19 timing
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_reference_gpu.py tests/test_b200_workloads.py -x -q -m gpu --timeout 600 -p no:cacheprovider \
   -k "reference_gpu or h2o1024_full or nonortho_water_full" 2>&1 | tail -15
echo "=== reference GPU backend, H2O-256"
timeout 600 python bench.py --impl reference-gpu --steps 3 --warmup 1 2>gpurun_out/refgpu.err | tee gpurun_out/bench_refgpu_h2o256.json | cut -c1-1500
tail -3 gpurun_out/refgpu.err
