cd $GRAFT_REPO_ROOT
ncu --set full --clock-control none --import-source on -k regex:"pab_to_coef|coef_to_hab" -s 2 -c 2 -o gpurun_out/prof_coef -f \
  python bench.py --workload H2O-64 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_coef.log 2>&1
tail -2 gpurun_out/ncu_coef.log
