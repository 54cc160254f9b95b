#!/bin/bash
# round 2, two GPUs: NCCL + kernels + exchange parity vs one GPU (all decompositions), then H2O-256 timings
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | tail -15
for mode in "--decomp blocks" "--decomp slab" "--decomp slab --slab-compact" "--decomp slab --slab-compact --halo torch"; do
  tag=$(echo $mode | tr -d ' -')
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 \
     bench.py --gpus 2 --steps 10 --warmup 3 $mode 2>gpurun_out/bench2_$tag.err > gpurun_out/bench_h2o256_2gpu_${tag}_r02.json
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_h2o256_2gpu_${tag}_r02.json') if l.startswith('{')][-1])
    print("$mode", "ms/step", round(d['ms_per_step'],3), "e2e", round(d['e2e']['ms_per_step'],3), d['multi_gpu_parity'], {k: round(v,2) for k,v in d['roofline']['phase_ms_per_step'].items()})
except Exception as e:
    print("$mode FAILED", e); print(open('gpurun_out/bench2_$tag.err').read()[-1500:])
PY
done
