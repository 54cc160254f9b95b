#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_b200_parity.py tests/test_b200_pgf_product.py -x -q -m gpu -k "generic or general or triclinic or nonortho" 2>&1 | tail -2
timeout 200 python -m pytest tests/test_b200_workloads.py -x -q -m gpu -k "nonortho" 2>&1 | tail -2
timeout 200 python bench.py --workload H2O-64_nonortho --tau --virial --no-reference-gpu --no-cpu-baseline > gpurun_out/bench_last_c4.json 2> gpurun_out/bench_last_c4.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_last_c4.json').read().strip().splitlines()[-1]);print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['phase_ms_per_step'])"
