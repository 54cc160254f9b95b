#!/bin/bash
mkdir -p gpurun_out
timeout 120 python bench.py --no-reference-gpu --no-cpu-baseline > gpurun_out/bench_head.json 2> gpurun_out/bench_head.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_head.json').read().strip().splitlines()[-1]);print(d['ms_per_step'], d['e2e']['ms_per_step'], d['create_task_list']['ms'], d['roofline']['frac'], d['clocks'])"
