#!/usr/bin/env python
"""Convert the reference's golden `.task` vectors into compact .npz fixtures.

Run in the build container (needs /root/reference):
    python tests/golden/make_golden.py

Source: /root/reference/src/grid/sample_tasks/*.task ("#Grid task v10", written
by src/grid/cpu/grid_cpu_collocate.c:47-139, read by src/grid/grid_replay.c:242-349).
Each fixture keeps every input of the task plus the expected outputs: the
non-zero collocated grid points, the hab sub-block, force_a/force_b and virial.
"""
import glob
import os
import sys

import numpy as np

SRC = "/root/reference/src/grid/sample_tasks"
DST = os.path.dirname(os.path.abspath(__file__))


def ncoset(l):
    return (l + 1) * (l + 2) * (l + 3) // 6


def parse(path):
    with open(path) as fh:
        lines = fh.read().splitlines()
    assert lines[0] == "#Grid task v10" and lines[-1] == "#THE_END"
    it = iter(lines[1:-1])

    def take(key, n, conv):
        parts = next(it).split()
        assert parts[0] == key, (parts, key)
        vals = [conv(x) for x in parts[len(parts) - n:]]
        return vals[0] if n == 1 else vals

    t = {}
    for key in ("orthorhombic", "border_mask", "func", "la_max", "la_min", "lb_max", "lb_min"):
        t[key] = take(key, 1, int)
    for key in ("zeta", "zetb", "rscale"):
        t[key] = take(key, 1, float)
    t["dh"] = np.array([take("dh", 3, float) for _ in range(3)])
    t["dh_inv"] = np.array([take("dh_inv", 3, float) for _ in range(3)])
    t["ra"] = np.array(take("ra", 3, float))
    t["rab"] = np.array(take("rab", 3, float))
    for key in ("npts_global", "npts_local", "shift_local", "border_width"):
        t[key] = np.array(take(key, 3, int), dtype=np.int32)
    t["radius"] = take("radius", 1, float)
    for key in ("o1", "o2", "n1", "n2"):
        t[key] = take(key, 1, int)
    n1, n2 = t["n1"], t["n2"]
    pab = np.zeros((n2, n1))
    for i in range(n2):
        for j in range(n1):
            p = next(it).split()
            assert p[0] == "pab" and int(p[1]) == i and int(p[2]) == j
            pab[i, j] = float(p[3])
    t["pab"] = pab
    nnz = take("ngrid_nonzero", 1, int)
    nl = t["npts_local"]
    gidx = np.zeros(nnz, dtype=np.int64)
    gval = np.zeros(nnz)
    for n in range(nnz):
        p = next(it).split()
        assert p[0] == "grid"
        i, j, k = int(p[1]), int(p[2]), int(p[3])
        gidx[n] = (k * nl[1] + j) * nl[0] + i
        gval[n] = float(p[4])
    t["grid_idx"], t["grid_val"] = gidx, gval
    na, nb = ncoset(t["la_max"]), ncoset(t["lb_max"])
    hab = np.zeros((nb, na))
    for i in range(t["o2"], t["o2"] + nb):
        for j in range(t["o1"], t["o1"] + na):
            p = next(it).split()
            assert p[0] == "hab" and int(p[1]) == i and int(p[2]) == j
            hab[i - t["o2"], j - t["o1"]] = float(p[3])
    t["hab"] = hab
    t["force_a"] = np.array(take("force_a", 3, float))
    t["force_b"] = np.array(take("force_b", 3, float))
    t["virial"] = np.array([take("virial", 3, float) for _ in range(3)])
    assert next(it, None) is None
    return t


def main():
    files = sorted(glob.glob(os.path.join(SRC, "*.task")))
    if not files:
        sys.exit(f"no .task files under {SRC}")
    for f in files:
        t = parse(f)
        out = os.path.join(DST, os.path.basename(f).replace(".task", ".npz"))
        np.savez_compressed(out, **t)
        print(f"{os.path.basename(out):32s} nnz={t['grid_idx'].size:6d} "
              f"{os.path.getsize(out) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
