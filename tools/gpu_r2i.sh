#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_b200_parity.py tests/test_reference_gpu.py -x -q -m gpu --timeout 300 -p no:cacheprovider -k "warptile or reference_gpu" 2>&1 | tail -5
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu 2>gpurun_out/bench.err > gpurun_out/bench_h2o256_r02b.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_h2o256_r02b.json').read())
print("ms/step", d['ms_per_step'], "e2e", d['e2e']['ms_per_step'], d['roofline']['phase_ms_per_step'], d['create_task_list'])
PY
tail -3 gpurun_out/bench.err
