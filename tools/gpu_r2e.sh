#!/bin/bash
# round 2: whole GPU suite (in-kernel exp in the warp-tile kernels, new full-size parity tests,
# reference GPU backend comparator), then the bench with its reference-gpu leg
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 900 -p no:cacheprovider 2>&1 | tail -15
echo "=== bench (default variant)"
timeout 900 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench.err > gpurun_out/bench_h2o256_r02a.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_h2o256_r02a.json').read())
print("ms/step", d['ms_per_step'], "e2e", d['e2e']['ms_per_step'], d['roofline']['phase_ms_per_step'])
print("cpu", d['cpu_baseline'])
print("refgpu", d['reference_gpu'])
PY
tail -3 gpurun_out/bench.err
cat /tmp/grid_b200_abort.log 2>/dev/null | tail -5
