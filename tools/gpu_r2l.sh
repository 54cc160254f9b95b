#!/bin/bash
# round 2, eight GPUs: replicated-grid (blocks) mode against the z-slab rs_grid partition with the
# library's NCCL halo exchange; both verify their result against a one-GPU run (untimed)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for mode in "--decomp slab --slab-compact" "--decomp blocks"; do
  tag=$(echo $mode | tr -d ' -')
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29721 \
     bench.py --gpus 8 --steps 20 --warmup 3 $mode 2>gpurun_out/bench8_$tag.err > gpurun_out/bench_h2o256_8gpu_${tag}_r02.json
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_h2o256_8gpu_${tag}_r02.json') if l.startswith('{')][-1])
    print("$mode", "ms/step", round(d['ms_per_step'],3), "e2e", round(d['e2e']['ms_per_step'],3), d['multi_gpu_parity'], {k: round(v,2) for k,v in d['roofline']['phase_ms_per_step'].items()})
except Exception as e:
    print("$mode FAILED", e); print(open('gpurun_out/bench8_$tag.err').read()[-1500:])
PY
done
