# Round profile (run under gpurun, ONE GPU): launch list of the bench command +
# one full capture of the 16 tiled launches of one step + the coefficient kernels.
# Usage: bash tools/profile_round.sh <tag>
cd $GRAFT_REPO_ROOT
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks_$TAG.csv &
SMI=$!
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
kill $SMI
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"tiled_kernel" -s 16 -c 16 -o gpurun_out/prof_$TAG -f \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/prof_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"pab_to_coef|coef_to_hab" -s 12 -c 12 -o gpurun_out/prof_coef_$TAG -f \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/prof_coef_$TAG.log 2>&1
tail -c 400 gpurun_out/bench_$TAG.json
