#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_b200_parity.py -x -q -m gpu -k "warptile or resident or reuse or empty" 2>&1 | tail -2
timeout 200 python bench.py --no-reference-gpu --no-cpu-baseline > gpurun_out/bench_last.json 2> gpurun_out/bench_last.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_last.json').read().strip().splitlines()[-1]);print(d['ms_per_step'], d['e2e'])"
