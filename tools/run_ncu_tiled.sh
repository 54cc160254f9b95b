# usage (on the GPU box): bash tools/run_ncu_tiled.sh [skip] [count] [workload]
cd $GRAFT_REPO_ROOT
SKIP=${1:-0}; COUNT=${2:-8}; WL=${3:-H2O-64}
ncu --set full --clock-control none --import-source on -k regex:tiled_kernel -s $SKIP -c $COUNT -o gpurun_out/prof_tiled -f \
  python bench.py --workload $WL --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log
