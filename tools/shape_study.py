#!/usr/bin/env python
"""Monte-Carlo study of register-tile shapes for the grid kernels (no GPU needed).

For the (radius index, lp) distribution of the H2O-64 TZV2P task list and random sub-block
offsets of the cube centre, counts per task: the (task, block) pairs, the warp-DFMAs a kernel
issues (per pair: ncoset(lp) + C*T2(lp) set-up, then C*(lp+1) per plane any lane needs; C =
columns per thread) and the useful ones (in-sphere points * (lp+1) / 32).  The sphere is the
reference's discretised-radius loop nest (src/grid/ref/grid_ref_collint.h:237-254, 144-147,
46-50).  DESIGN.md section 7 quotes the table this prints.
usage: python tools/shape_study.py"""
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cp2k_b200.workload import build_h2o_workload  # noqa: E402

rng = np.random.default_rng(1)
T2 = lambda l: (l + 1) * (l + 2) // 2  # noqa: E731
nco = lambda l: (l + 1) * (l + 2) * (l + 3) // 6  # noqa: E731


def task_distribution():
    w = build_h2o_workload("H2O-64")
    t = w.tasks
    lmax = [np.array(b.lmax) for b in w.basis_sets]
    ia, ja = t["iatom_list"] - 1, t["jatom_list"] - 1
    la = np.array([lmax[w.atom_kinds[a] - 1][s - 1] for a, s in zip(ia, t["iset_list"])])
    lb = np.array([lmax[w.atom_kinds[a] - 1][s - 1] for a, s in zip(ja, t["jset_list"])])
    n = np.zeros(w.ntasks, int)
    for lev in range(len(w.layouts)):
        m = t["level_list"] == lev + 1
        n[m] = np.maximum(1, np.ceil(t["radius_list"][m] / w.layouts[lev].dh[0, 0])).astype(int)
    dist = {}
    for k in zip(n, la + lb):
        dist[k] = dist.get(k, 0) + 1
    return dist


def sphere(n):
    R = float(max(1, n))
    nb = -math.ceil(-1e-8 - R)
    sz = 2 * nb + 2
    cube = np.zeros((sz, sz, sz), bool)
    for kd in range(nb + 1):
        krem = R * R - kd * kd
        js = math.ceil(-1e-8 - math.sqrt(max(0, krem)))
        for jd in range(-js + 1):
            jrem = krem - jd * jd
            is_ = math.ceil(-1e-8 - math.sqrt(max(0, jrem)))
            for id_ in range(-is_ + 1):
                for k in (-kd, kd + 1):
                    for j in (-jd, jd + 1):
                        for i in (-id_, id_ + 1):
                            cube[k + nb, j + nb, i + nb] = True
    return cube


def study(shape, C, lp, n, trials, halfskip):
    bx, by, bz = shape
    cube = sphere(n)
    sz = cube.shape[0]
    tot_pairs = tot_dfma = 0
    for _ in range(trials):
        ox, oy, oz = rng.integers(0, bx), rng.integers(0, by), rng.integers(0, bz)
        X, Y, Z = -(-(sz + ox) // bx) * bx, -(-(sz + oy) // by) * by, -(-(sz + oz) // bz) * bz
        A = np.zeros((Z, Y, X), bool)
        A[oz:oz + sz, oy:oy + sz, ox:ox + sz] = cube
        if not halfskip:
            plane_any = A.reshape(Z // bz, bz, Y // by, by, X // bx, bx).any(axis=(3, 5))
            npairs, nplanes = plane_any.any(axis=1).sum(), plane_any.sum()
            dfma = npairs * (nco(lp) + C * T2(lp)) + nplanes * C * (lp + 1)
        else:  # the two column sets (y halves) of a warp are skipped independently
            plane_any = A.reshape(Z // bz, bz, Y // by, 2, by // 2, X // bx, bx).any(axis=(4, 6))
            set_any = plane_any.any(axis=1)
            npairs = set_any.any(axis=2).sum()
            dfma = npairs * nco(lp) + set_any.sum() * T2(lp) + plane_any.sum() * (lp + 1)
        tot_pairs += npairs
        tot_dfma += dfma
    return tot_pairs / trials, tot_dfma / trials, cube.sum()


def main():
    dist = task_distribution()
    shapes = {"8x8x16, 2 columns/thread (both kernel families)": ((8, 8, 16), 2, False),
              "8x8x16, second column set skipped when idle": ((8, 8, 16), 2, True),
              "8x4x32, 1 column/thread": ((8, 4, 32), 1, False),
              "8x8x32, 2 columns/thread": ((8, 8, 32), 2, False),
              "8x8x32, with the skip": ((8, 8, 32), 2, True),
              "8x4x16, 1 column/thread": ((8, 4, 16), 1, False),
              "8x8x8, 2 columns/thread": ((8, 8, 8), 2, False)}
    print(f"{'shape':50s} {'pairs/task':>10s} {'DFMA/task':>10s} {'useful':>8s} {'eff':>6s}")
    for name, (shape, C, hs) in shapes.items():
        P = D = U = W = 0
        for (n, l), cnt in dist.items():
            if l > 2:
                continue
            p, d, pts = study(shape, C, l, n, 12, hs)
            P += p * cnt
            D += d * cnt
            U += pts * (l + 1) / 32 * cnt
            W += cnt
        print(f"{name:50s} {P / W:10.1f} {D / W:10.1f} {U / W:8.1f} {U / D:6.3f}")


if __name__ == "__main__":
    main()
