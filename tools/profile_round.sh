# Round profile (run under gpurun, ONE GPU): the default bench line and the
# reference arm, the ncu launch list of the bench command, one full capture of
# the tiled launches of one step and of the coefficient kernels.  The
# .ncu-rep files are summarised on the box (they exceed gpurun's 64 MiB return
# limit) and only the text/csv summaries come back.
# Usage: bash tools/profile_round.sh <tag>
cd $GRAFT_REPO_ROOT
TAG=${1:-r02}
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-reference-gpu > gpurun_out/launches_$TAG.log 2>&1
# 14 tiled launches per step (7 collocate + 7 integrate): skip the warm-up step, capture the next
ncu --set full --clock-control none -k regex:"tiled_kernel" -s 14 -c 14 -o /tmp/prof_$TAG -f \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-reference-gpu > gpurun_out/prof_$TAG.log 2>&1
python tools/ncu_summary.py /tmp/prof_$TAG.ncu-rep > gpurun_out/ncu_tiled_$TAG.txt 2>&1
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/ncu_tiled_${TAG}_raw.csv 2>/dev/null
python tools/traffic_from_ncu.py gpurun_out/ncu_tiled_${TAG}_raw.csv H2O-256 > gpurun_out/traffic_$TAG.json
ncu --set full --clock-control none -k regex:"pab_to_coef|coef_to_hab" -s 6 -c 6 -o /tmp/prof_coef_$TAG -f \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-reference-gpu > gpurun_out/prof_coef_$TAG.log 2>&1
python tools/ncu_summary.py /tmp/prof_coef_$TAG.ncu-rep > gpurun_out/ncu_coef_$TAG.txt 2>&1
# the CTA-tile family (variant 3), same capture
ncu --set full --clock-control none -k regex:"ctile_kernel" -s 6 -c 6 -o /tmp/prof_ctile_$TAG -f \
  python bench.py --steps 1 --warmup 1 --variant 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/prof_ctile_$TAG.log 2>&1
python tools/ncu_summary.py /tmp/prof_ctile_$TAG.ncu-rep > gpurun_out/ncu_ctile_$TAG.txt 2>&1
python bench.py --variant 3 --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/bench_ctile_$TAG.json 2> gpurun_out/bench_ctile_$TAG.err
tail -c 300 gpurun_out/bench_$TAG.json; ls -la gpurun_out | tail -15
