"""The multi-GPU product path on hardware: NCCL + the B200 kernels + the exchange step
(sum of replicated grids -- one grouped NCCL all-reduce, per-level all-reduces inside collocate, or
the copy-engine exchange through NVLink peer memory --, or the z-slab halo sum / fill and the owner
reduction of H)
against a single-GPU run of the full task list, on H2O-64 at 2 ranks, for every decomposition
bench.py offers.  Needs two GPUs on the box (skipped otherwise); the same check runs untimed
inside every `bench.py --gpus N` (N > 1) run on the benchmark's own workload."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("extra", [[], ["--allreduce", "fused"], ["--allreduce", "peer"], ["--decomp", "slab"],
                                   ["--decomp", "slab", "--slab-compact"]],
                         ids=["blocks", "blocks-fused-nccl", "blocks-peer-memory", "slab", "slab-compact"])
def test_two_ranks_match_one_gpu(b200, extra):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "bench.py"),
           "--gpus", "2", "--workload", "H2O-64", "--steps", "2", "--warmup", "1"] + extra
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    par = line["multi_gpu_parity"]
    assert par is not None and par["ok"] and par["grid_max_rel"] < 1e-10 and par["hab_max_rel"] < 1e-10, par
