#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_b200_parity.py tests/test_b200_workloads.py tests/test_b200_dispatcher.py tests/test_b200_pgf_product.py -x -q -m gpu 2>&1 | tail -3
show() { python -c "
import json,sys;d=json.loads(open('$1').read().strip().splitlines()[-1]);print('$1',d['ms_per_step'],d['e2e']['ms_per_step'],d['create_task_list'],d['roofline']['phase_ms_per_step'])"; }
timeout 300 python bench.py --workload H2O-64_nonortho --tau --virial --no-reference-gpu --no-cpu-baseline > gpurun_out/bench_z_c4.json 2> gpurun_out/bench_z_c4.err; show gpurun_out/bench_z_c4.json
timeout 300 python bench.py --forces --no-reference-gpu --no-cpu-baseline > gpurun_out/bench_z_f.json 2> gpurun_out/bench_z_f.err; show gpurun_out/bench_z_f.json
GRID_B200_CREATE_TIMING=1 timeout 300 python bench.py --no-reference-gpu --no-cpu-baseline > gpurun_out/bench_z256.json 2> gpurun_out/bench_z256.err; show gpurun_out/bench_z256.json
