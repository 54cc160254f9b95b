#!/usr/bin/env python
"""BASELINE config 1: the reference's 13 sample tasks (src/grid/sample_tasks/*.task, held as
tests/golden/*.npz), each replayed as a batch of `cycles` identical products on distinct matrix
blocks (grid_replay.c:361-412 with the aliasing of block offsets removed), timed per task for
collocate and integrate (+ forces + virial, as the batched replay always requests):
  * this backend, device-resident buffers (CUDA events),
  * the reference GPU backend and the reference CPU backend through the reference's public API
    (wall clock; host buffers, as grid_miniapp.x times them).
usage: python tools/bench_sample_tasks.py [--cycles 10000] [--repeat 5]   (needs a GPU)
Prints one JSON object; profiles/r02/sample_tasks_r02.json is its output on a B200."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cycles", type=int, default=10000)
    ap.add_argument("--cycles-per-block", type=int, default=100)
    ap.add_argument("--repeat", type=int, default=5)
    args = ap.parse_args()
    import torch

    from cp2k_b200 import OffloadBuffer, load_b200
    from cp2k_b200.grid_api import GRID_BACKEND_CPU
    from oracle import pyref
    from replay import TASK_NAMES, dummy_task_list, golden_grid, load_task

    lib = load_b200()
    refs = {}
    if pyref.have_reference():
        refs["reference_cpu"] = pyref.load_reference(GRID_BACKEND_CPU)
    if pyref.have_reference_gpu():
        refs["reference_gpu"] = pyref.load_reference_gpu(0)
    out = {"cycles": args.cycles, "cycles_per_block": args.cycles_per_block, "unit": "ms per call, best of %d" % args.repeat,
           "tasks": {}}
    for name in TASK_NAMES:
        t = load_task(name)
        n1, n2 = t["n1"], t["n2"]
        row = {}
        for who, L in [("b200", lib)] + list(refs.items()):
            dev = who != "reference_cpu"
            mk = OffloadBuffer.with_device if dev else OffloadBuffer
            if who == "b200":
                lib.set_device_resident(True)
            tl, nblocks = dummy_task_list(L, t, args.cycles, args.cycles_per_block)
            pab, hab = mk(nblocks * n1 * n2), mk(nblocks * n1 * n2)
            pab.host.reshape(nblocks, n2, n1)[:] = 0.5 * t["rscale"] * t["pab"]
            grid = mk(int(np.prod(t["npts_local"])))
            grid.host[:] = golden_grid(t)
            if dev:
                pab.device.copy_(torch.from_numpy(pab.host))
                grid.device.copy_(torch.from_numpy(grid.host))
            forces, virial = np.zeros((2, 3)), np.zeros((3, 3))
            best = {"collocate": 1e30, "integrate": 1e30}
            for rep in range(args.repeat + 1):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                tl.collocate(t["func"], pab, [grid])
                torch.cuda.synchronize()
                t1 = time.perf_counter()
                if who == "b200":
                    grid.device.copy_(torch.from_numpy(golden_grid(t)))
                    torch.cuda.synchronize()
                    t1b = time.perf_counter()
                else:
                    grid.host[:] = golden_grid(t)
                    t1b = time.perf_counter()
                tl.integrate(t["func"] == 200, pab, [grid], hab, forces, virial)
                torch.cuda.synchronize()
                t2 = time.perf_counter()
                if rep > 0:
                    best["collocate"] = min(best["collocate"], (t1 - t0) * 1e3)
                    best["integrate"] = min(best["integrate"], (t2 - t1b) * 1e3)
            tl.free()
            if who == "b200":
                lib.set_device_resident(False)
            row[who] = best
        out["tasks"][name] = row
        print(name, {k: {a: round(b, 3) for a, b in v.items()} for k, v in row.items()}, file=sys.stderr)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
