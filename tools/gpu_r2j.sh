#!/bin/bash
# round 2: bench lines of every BASELINE config + ncu captures of the default kernels
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
run() { # name, args...
  local name=$1; shift
  GRID_B200_CREATE_TIMING=1 timeout 900 python bench.py "$@" 2>gpurun_out/$name.err > gpurun_out/$name.json
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/$name.json') if l.startswith('{')][-1])
    print("$name", "ms/step", round(d['ms_per_step'],2), "e2e", round(d['e2e']['ms_per_step'],2), "e2e(P/H only)", d['e2e'].get('ms_per_step_p_in_h_out_grids_resident'),
          "frac", round(d['roofline']['frac'],4), {k: round(v,2) for k,v in d['roofline']['phase_ms_per_step'].items()}, "create", d.get('create_task_list'),
          "cpu", (d.get('cpu_baseline') or {}).get('ms_per_step_sample'), "refgpu", (d.get('reference_gpu') or {}).get('ms_per_step'))
except Exception as e:
    print("$name FAILED", e)
PY
  grep "grid_b200 create" gpurun_out/$name.err | head -12
}
run bench_h2o256_r02c --steps 20 --warmup 3
run bench_h2o256_forces_r02c --steps 5 --warmup 2 --forces --no-cpu-baseline --no-reference-gpu
run bench_h2o64_nonortho_tau_virial_r02c --workload H2O-64_nonortho --tau --virial --steps 5 --warmup 2
run bench_h2o64_molopt_r02c --workload H2O-64 --basis DZVP-MOLOPT-SR-GTH --steps 10 --warmup 3
run bench_h2o1024_r02c --workload H2O-1024 --steps 5 --warmup 2 --no-reference-gpu
echo "=== ncu full: warp-tile kernels, H2O-256"
timeout 900 ncu --set full --import-source on --clock-control none -k tiled_kernel -c 6 -o gpurun_out/ncu_tiled_h2o256_r02 -f \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-reference-gpu > gpurun_out/ncu_tiled.log 2>&1
tail -2 gpurun_out/ncu_tiled.log | cut -c1-200
echo "=== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_h2o256_r02.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-reference-gpu > gpurun_out/launches.log 2>&1
ls -la gpurun_out | tail -12
