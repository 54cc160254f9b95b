"""Seeded synthetic task lists for parity tests (test helper, not product code).

Generates a small molecular-like system with the same structure a real
`generate_qs_task_list` (src/task_list_methods.F:117-501) hands to
`grid_create_task_list`: several atoms of two kinds, multi-set contracted basis
sets with dense random sphi, all (iset,jset,ipgf,jpgf) products of the chosen
atom pairs, radii from the Gaussian decay, grid level chosen from the exponent
(cf. src/pw_env/gaussian_gridlevels.F:147-167), one matrix block per atom pair.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List

import numpy as np

from cp2k_b200.grid_api import BasisSet, GridLayout, OffloadBuffer
from cp2k_b200.workload import Workload


def ncoset(l):
    return (l + 1) * (l + 2) * (l + 3) // 6 if l >= 0 else 0


def random_basis(rng, sets, zet_range=(0.15, 3.0)) -> BasisSet:
    """sets = [(lmin, lmax, npgf, nsgf_set), ...]"""
    nset = len(sets)
    lmin = [s[0] for s in sets]
    lmax = [s[1] for s in sets]
    npgf = [s[2] for s in sets]
    nsgf_set = [s[3] for s in sets]
    first_sgf = np.cumsum([1] + nsgf_set[:-1])
    nsgf = int(sum(nsgf_set))
    maxco = max(n * ncoset(l) for n, l in zip(npgf, lmax))
    maxpgf = max(npgf)
    sphi = np.zeros((nsgf, maxco))
    zet = np.zeros((nset, maxpgf))
    for i, (lo, hi, n, ns) in enumerate(sets):
        zet[i, :n] = np.exp(rng.uniform(np.log(zet_range[0]), np.log(zet_range[1]), n))
        block = rng.normal(size=(ns, n * ncoset(hi))) * 0.5
        # functions below lmin do not exist in the set: zero their columns like CP2K does
        for p in range(n):
            block[:, p * ncoset(hi): p * ncoset(hi) + ncoset(lo - 1)] = 0.0
        sphi[first_sgf[i] - 1: first_sgf[i] - 1 + ns, : n * ncoset(hi)] = block
    return BasisSet(lmin, lmax, npgf, nsgf_set, first_sgf, sphi, zet)


def make_cell(rng, edge, orthorhombic, skew=0.12):
    cell = np.diag(np.asarray(edge, dtype=np.float64))
    if not orthorhombic:
        cell = cell + rng.uniform(-skew, skew, size=(3, 3)) * np.mean(edge)
    return cell


def make_layouts(cell, npts_list, slab=None):
    """Periodic layouts, or (slab=(lo, hi, border)) a z-slab of a distributed
    grid: owned planes [lo,hi) of the FINEST level scaled to each level, plus
    `border` halo planes on both sides (src/grid/grid_api.F:501-547)."""
    layouts = []
    for ilev, npts in enumerate(npts_list):
        npts = np.asarray(npts, dtype=np.int32)
        dh = cell / npts[:, None].astype(np.float64)
        dh_inv = np.linalg.inv(dh)
        if slab is None:
            layouts.append(GridLayout(npts, npts.copy(), np.zeros(3, np.int32), np.zeros(3, np.int32), dh, dh_inv))
        else:
            lo_f, hi_f, border = slab
            lo = int(round(lo_f * npts[2])), int(round(hi_f * npts[2]))
            nloc = npts.copy()
            nloc[2] = (lo[1] - lo[0]) + 2 * border
            assert nloc[2] <= npts[2]
            shift = np.array([0, 0, lo[0] - border], dtype=np.int32)
            layouts.append(GridLayout(npts, nloc, shift, np.array([0, 0, border], np.int32), dh, dh_inv))
    return layouts


def make_workload(seed=0, natoms=6, orthorhombic=True, edge=(9.0, 10.0, 11.0),
                  npts_list=((45, 50, 54), (24, 25, 27)), sets_a=((0, 1, 3, 5), (2, 2, 2, 6)),
                  sets_b=((0, 0, 3, 2), (1, 1, 1, 3)), pair_fraction=0.6, eps=1e-10,
                  max_tasks=None, both_orders=True, border_mask_fraction=0.0) -> Workload:
    rng = np.random.default_rng(seed)
    cell = make_cell(rng, edge, orthorhombic)
    layouts = make_layouts(cell, npts_list)
    nlevels = len(layouts)
    basis = [random_basis(rng, list(sets_a)), random_basis(rng, list(sets_b))]
    kinds = rng.integers(1, 3, size=natoms).astype(np.int32)
    frac = rng.uniform(0, 1, size=(natoms, 3))
    pos = frac @ cell
    # resolution of each level ~ largest |dh| row
    hmax = [float(np.max(np.linalg.norm(l.dh, axis=1))) for l in layouts]

    cols = {k: [] for k in ("level_list", "iatom_list", "jatom_list", "iset_list", "jset_list", "ipgf_list",
                            "jpgf_list", "border_mask_list", "block_num_list", "radius_list", "rab_list")}
    block_offsets = []
    offset = 0
    nblocks = 0
    for i in range(natoms):
        for j in range(natoms):
            if j < i and not both_orders:
                continue
            if i != j and rng.uniform() > pair_fraction:
                continue
            bi, bj = basis[kinds[i] - 1], basis[kinds[j] - 1]
            # minimum-image like displacement plus an occasional lattice shift
            d = pos[j] - pos[i]
            if rng.uniform() < 0.3 and i != j:
                d = d + cell[rng.integers(0, 3)] * rng.choice([-1, 1])
            if np.linalg.norm(d) > 7.0:
                continue
            nblocks += 1
            block_offsets.append(offset)
            # block stored as (nsgf(row=min), nsgf(col=max)) like DBCSR upper triangle
            offset += bi.nsgf * bj.nsgf
            for iset in range(bi.nset):
                for jset in range(bj.nset):
                    for ipgf in range(bi.npgf[iset]):
                        for jpgf in range(bj.npgf[jset]):
                            za, zb = bi.zet[iset, ipgf], bj.zet[jset, jpgf]
                            zp = za + zb
                            pref = np.exp(-za * zb / zp * float(d @ d))
                            if pref < eps:
                                continue
                            radius = np.sqrt(max(1e-3, -np.log(eps / max(pref, eps)) / zp)) * rng.uniform(0.9, 1.2)
                            # coarsest level that still resolves the Gaussian with ~ >=4 points per radius
                            level = 1
                            for lev in range(nlevels, 0, -1):
                                if radius / hmax[lev - 1] >= 5.0:
                                    level = lev
                                    break
                            radius = min(radius, 14.0 * hmax[level - 1])
                            mask = 0
                            if border_mask_fraction > 0 and rng.uniform() < border_mask_fraction:
                                mask = int(rng.integers(1, 64))
                            cols["level_list"].append(level)
                            cols["iatom_list"].append(i + 1)
                            cols["jatom_list"].append(j + 1)
                            cols["iset_list"].append(iset + 1)
                            cols["jset_list"].append(jset + 1)
                            cols["ipgf_list"].append(ipgf + 1)
                            cols["jpgf_list"].append(jpgf + 1)
                            cols["border_mask_list"].append(mask)
                            cols["block_num_list"].append(nblocks)
                            cols["radius_list"].append(radius)
                            cols["rab_list"].append(d)
    tasks = {k: np.asarray(v) for k, v in cols.items()}
    n = tasks["level_list"].shape[0]
    if max_tasks is not None and n > max_tasks:
        keep = np.sort(rng.choice(n, size=max_tasks, replace=False))
        tasks = {k: v[keep] for k, v in tasks.items()}
    # shuffle: the backend must not rely on the caller's order
    perm = rng.permutation(tasks["level_list"].shape[0])
    tasks = {k: v[perm] for k, v in tasks.items()}
    return Workload(orthorhombic, natoms, pos, kinds, basis, layouts, np.asarray(block_offsets, dtype=np.int32),
                    tasks, offset, meta={"cell": cell})
