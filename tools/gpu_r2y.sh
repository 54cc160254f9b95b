#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
show() { python -c "
import json,sys;d=json.loads(open('$1').read().strip().splitlines()[-1]);print('$1',d['ms_per_step'],d['e2e']['ms_per_step'],d.get('multi_gpu_parity'),d.get('exchange_ms_per_step'),d.get('grid_sum'),d['roofline']['phase_ms_per_step'])"; }
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }

run --allreduce peer > gpurun_out/bench_y_peer_$N.json 2> gpurun_out/bench_y_peer_$N.err; show gpurun_out/bench_y_peer_$N.json || tail -30 gpurun_out/bench_y_peer_$N.err
