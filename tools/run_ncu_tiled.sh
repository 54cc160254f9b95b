set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python bench.py --workload H2O-64 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_h2o64_t2.json 2> gpurun_out/err.log; tail -c 900 gpurun_out/bench_h2o64_t2.json
ncu --set full --clock-control none --import-source on -k regex:tiled_kernel -s 4 -c 4 -o gpurun_out/prof_tiled -f python bench.py --workload H2O-64 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/ncu.log
