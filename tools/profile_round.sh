# Round profile (run under gpurun): launch list of the bench command + one full
# capture of the dominant kernels.  Usage: bash tools/profile_round.sh <tag>
cd $GRAFT_REPO_ROOT
TAG=${1:-r01}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tiled_kernel -s 8 -c 8 -o gpurun_out/prof_$TAG -f \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/prof_$TAG.log 2>&1
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/clocks_$TAG.csv
tail -2 gpurun_out/prof_$TAG.log
