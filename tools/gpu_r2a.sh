#!/bin/bash
# round 2, first GPU check of the CTA-tile kernels: parity, then a short bench of both families
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_b200_parity.py -x -q -m gpu -k "ctile" --timeout 180 -p no:cacheprovider 2>&1 | tail -25
echo "=== bench variant 0 (ctile)"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --variant 0 2>gpurun_out/bench_v0.err | tee gpurun_out/bench_v0.json
tail -5 gpurun_out/bench_v0.err
echo "=== bench variant 2 (warp tile)"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --variant 2 2>gpurun_out/bench_v2.err | tee gpurun_out/bench_v2.json
tail -3 gpurun_out/bench_v2.err
cat /tmp/grid_b200_abort.log 2>/dev/null | tail -5
