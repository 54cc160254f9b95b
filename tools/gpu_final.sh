#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 600 gpurun_out/bench_final.json; echo
timeout 300 python bench.py --impl reference > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; tail -c 400 gpurun_out/bench_final_ref.json; echo
python -c "
import json;d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['create_task_list'], d['cpu_baseline'] and d['cpu_baseline']['ms_per_step_sample'], d.get('reference_gpu'))"
