"""Test helper mirroring the reference's replay harness (src/grid/grid_replay.c):
feeds one golden `.task` vector through a backend, either as a single Gaussian
product (oracle only) or as a batched task list of `cycles` identical tasks,
and returns the reference's own error measure
    rel = |test - ref| / max(1, |ref|)          (grid_replay.c:449-456)
with forces/virial down-weighted by 1e-4 (:445,481,495).

Unlike the reference harness every matrix block gets its OWN offset
(grid_replay.c:175-176 aliases them all to 0, which races in integrate).
"""
from __future__ import annotations

import ctypes as C
import glob
import os

import numpy as np

from cp2k_b200.grid_api import BasisSet, GridLayout, OffloadBuffer, _dp, _ip, _i32, _f64

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TASK_NAMES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz"))
                    if not os.path.basename(p).startswith("ref_"))


def ncoset(l: int) -> int:
    return (l + 1) * (l + 2) * (l + 3) // 6 if l >= 0 else 0


def load_task(name: str) -> dict:
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    t = {k: z[k] for k in z.files}
    for k in ("orthorhombic", "border_mask", "func", "la_max", "la_min", "lb_max", "lb_min", "o1", "o2",
              "n1", "n2"):
        t[k] = int(t[k])
    for k in ("zeta", "zetb", "rscale", "radius"):
        t[k] = float(t[k])
    return t


def write_task_file(t: dict, path: str) -> None:
    """Writes a golden fixture back in the reference's "#Grid task v10" text format
    (written by src/grid/cpu/grid_cpu_collocate.c:47-139, read by
    src/grid/grid_replay.c:242-349) so that the reference's own replay harness can
    read it: the inverse of tests/golden/make_golden.py::parse."""
    e = lambda x: "%.21e" % float(x)  # noqa: E731  (doubles round-trip exactly)
    out = ["#Grid task v10"]
    for k in ("orthorhombic", "border_mask", "func", "la_max", "la_min", "lb_max", "lb_min"):
        out.append(f"{k} {int(t[k])}")
    for k in ("zeta", "zetb", "rscale"):
        out.append(f"{k} {e(t[k])}")
    for k in ("dh", "dh_inv"):
        for i in range(3):
            out.append(f"{k} {i} " + " ".join(e(x) for x in t[k][i]))
    for k in ("ra", "rab"):
        out.append(f"{k} " + " ".join(e(x) for x in t[k]))
    for k in ("npts_global", "npts_local", "shift_local", "border_width"):
        out.append(f"{k} " + " ".join(str(int(x)) for x in t[k]))
    out.append(f"radius {e(t['radius'])}")
    for k in ("o1", "o2", "n1", "n2"):
        out.append(f"{k} {int(t[k])}")
    n1, n2 = int(t["n1"]), int(t["n2"])
    for i in range(n2):
        for j in range(n1):
            out.append(f"pab {i} {j} {e(t['pab'][i, j])}")
    nl = [int(x) for x in t["npts_local"]]
    out.append(f"ngrid_nonzero {int(t['grid_idx'].size)}")
    for idx, val in zip(t["grid_idx"], t["grid_val"]):
        idx = int(idx)
        i, j, k = idx % nl[0], (idx // nl[0]) % nl[1], idx // (nl[0] * nl[1])
        out.append(f"grid {i} {j} {k} {e(val)}")
    na, nb = ncoset(int(t["la_max"])), ncoset(int(t["lb_max"]))
    for i in range(nb):
        for j in range(na):
            out.append(f"hab {int(t['o2']) + i} {int(t['o1']) + j} {e(t['hab'][i, j])}")
    out.append("force_a " + " ".join(e(x) for x in t["force_a"]))
    out.append("force_b " + " ".join(e(x) for x in t["force_b"]))
    for i in range(3):
        out.append(f"virial {i} " + " ".join(e(x) for x in t["virial"][i]))
    out.append("#THE_END")
    with open(path, "w") as fh:
        fh.write("\n".join(out) + "\n")


def golden_grid(t: dict) -> np.ndarray:
    g = np.zeros(int(np.prod(t["npts_local"])))
    g[t["grid_idx"]] = t["grid_val"]
    return g


def layout_of(t: dict) -> GridLayout:
    return GridLayout(t["npts_global"], t["npts_local"], t["shift_local"], t["border_width"], t["dh"],
                      t["dh_inv"])


def dummy_basis(size: int, lmin: int, lmax: int, zet: float) -> BasisSet:
    """grid_replay.c:114-148 -- one set, identity sphi, all exponents equal."""
    npgf = size // ncoset(lmax)
    assert size == npgf * ncoset(lmax)
    return BasisSet([lmin], [lmax], [npgf], [size], [1], np.eye(size), np.full((1, npgf), zet))


def rel_diff(test, ref) -> float:
    test, ref = np.asarray(test, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    if test.size == 0:
        return 0.0
    return float(np.max(np.abs(test - ref) / np.maximum(1.0, np.abs(ref))))


def norm_rel_diff(test, ref) -> float:
    """Norm-wise relative error ||test - ref||_2 / ||ref||_2 (north_star's "relative error",
    beside the reference validator's element-wise |d| / max(1, |ref|) of `rel_diff`)."""
    test, ref = np.asarray(test, dtype=np.float64).ravel(), np.asarray(ref, dtype=np.float64).ravel()
    den = float(np.linalg.norm(ref))
    if den == 0.0:
        return float(np.linalg.norm(test))
    return float(np.linalg.norm(test - ref)) / den


def assert_parity(got, ref, tol: float, what: str = "") -> None:
    """Both measures must hold: the reference validator's and the norm-wise relative one."""
    e1, e2 = rel_diff(got, ref), norm_rel_diff(got, ref)
    assert e1 < tol, f"{what}: |d|/max(1,|ref|) = {e1:.3e} >= {tol:g}"
    assert e2 < tol, f"{what}: ||d||/||ref|| = {e2:.3e} >= {tol:g}"


def dummy_task_list(lib, t: dict, cycles: int, cycles_per_block: int):
    """grid_replay.c:154-214 with distinct block offsets."""
    n1, n2 = t["n1"], t["n2"]
    nblocks = 1 if cycles == 1 else cycles // cycles_per_block + 1
    ra, rab = t["ra"], t["rab"]
    basis_a = dummy_basis(n1, t["la_min"], t["la_max"], t["zeta"])
    basis_b = dummy_basis(n2, t["lb_min"], t["lb_max"], t["zetb"])
    ipgf = t["o1"] // ncoset(t["la_max"]) + 1
    jpgf = t["o2"] // ncoset(t["lb_max"]) + 1
    ones = np.ones(cycles, dtype=np.int32)
    tl = lib.create_task_list(
        orthorhombic=bool(t["orthorhombic"]), natoms=2,
        block_offsets=np.arange(nblocks, dtype=np.int32) * (n1 * n2),
        atom_positions=np.array([ra, ra + rab]), atom_kinds=[1, 2], basis_sets=[basis_a, basis_b],
        level_list=ones, iatom_list=ones, jatom_list=2 * ones, iset_list=ones, jset_list=ones,
        ipgf_list=ipgf * ones, jpgf_list=jpgf * ones, border_mask_list=t["border_mask"] * ones,
        block_num_list=np.arange(cycles, dtype=np.int32) // cycles_per_block + 1,
        radius_list=np.full(cycles, t["radius"]), rab_list=np.tile(rab, (cycles, 1)),
        layouts=[layout_of(t)],
    )
    return tl, nblocks


def replay_batched(lib, t: dict, collocate: bool, cycles: int = 1, cycles_per_block: int = 1,
                   make_buffer=OffloadBuffer) -> float:
    """grid_replay.c:361-412 + comparison :443-503.  Returns max rel diff."""
    n1, n2 = t["n1"], t["n2"]
    tl, nblocks = dummy_task_list(lib, t, cycles, cycles_per_block)
    pab_blocks, hab_blocks = make_buffer(nblocks * n1 * n2), make_buffer(nblocks * n1 * n2)
    f = t["rscale"] if collocate else 1.0
    pab_blocks.host.reshape(nblocks, n2, n1)[:] = 0.5 * f * t["pab"]
    ntot = int(np.prod(t["npts_local"]))
    worst = 0.0
    if collocate:
        grid = make_buffer(ntot)
        grid.host[:] = 123.456  # must be overwritten, not accumulated
        tl.collocate(t["func"], pab_blocks, [grid])
        worst = rel_diff(grid.host, cycles * golden_grid(t))
    else:
        grid = make_buffer(ntot)
        grid.host[:] = golden_grid(t)
        forces, virial = np.full((2, 3), 7.0), np.full((3, 3), 7.0)
        hab_blocks.host[:] = 99.0  # must be overwritten
        tl.integrate(t["func"] == 200, pab_blocks, [grid], hab_blocks, forces, virial)
        hab = hab_blocks.host.reshape(nblocks, n2, n1)
        na, nb = ncoset(t["la_max"]), ncoset(t["lb_max"])
        counts = np.bincount(np.arange(cycles) // cycles_per_block, minlength=nblocks)
        for b in range(nblocks):
            ref = np.zeros((n2, n1))
            ref[t["o2"]:t["o2"] + nb, t["o1"]:t["o1"] + na] = counts[b] * t["hab"]
            worst = max(worst, rel_diff(hab[b], ref))
        worst = max(worst, 1e-4 * rel_diff(forces, cycles * np.stack([t["force_a"], t["force_b"]])))
        worst = max(worst, 1e-4 * rel_diff(virial, cycles * t["virial"]))
    tl.free()
    return worst


def replay_single_oracle(oracle, t: dict, collocate: bool) -> float:
    """grid_replay.c:413-441 with the oracle's single-product entry points."""
    L = oracle.lib
    dh, dh_inv = _f64(t["dh"]).reshape(-1), _f64(t["dh_inv"]).reshape(-1)
    ra, rab, pab = _f64(t["ra"]), _f64(t["rab"]), _f64(t["pab"]).reshape(-1)
    ng, nl = _i32(t["npts_global"]), _i32(t["npts_local"])
    sh, bw = _i32(t["shift_local"]), _i32(t["border_width"])
    n1, n2 = t["n1"], t["n2"]
    if collocate:
        grid = np.zeros(int(np.prod(nl)))
        L.grid_oracle_collocate_pgf_product(
            bool(t["orthorhombic"]), t["border_mask"], t["func"], t["la_max"], t["la_min"], t["lb_max"],
            t["lb_min"], t["zeta"], t["zetb"], t["rscale"], _dp(dh), _dp(dh_inv), _dp(ra), _dp(rab),
            _ip(ng), _ip(nl), _ip(sh), _ip(bw), t["radius"], t["o1"], t["o2"], n1, n2, _dp(pab), _dp(grid))
        return rel_diff(grid, golden_grid(t))
    grid = golden_grid(t)
    hab, forces, virials = np.zeros(n1 * n2), np.zeros(6), np.zeros(18)
    L.grid_oracle_integrate_pgf_product(
        bool(t["orthorhombic"]), t["func"] == 200, t["border_mask"], t["la_max"], t["la_min"], t["lb_max"],
        t["lb_min"], t["zeta"], t["zetb"], _dp(dh), _dp(dh_inv), _dp(ra), _dp(rab), _ip(ng), _ip(nl),
        _ip(sh), _ip(bw), t["radius"], t["o1"], t["o2"], n1, n2, _dp(grid), _dp(hab), _dp(pab),
        _dp(forces), _dp(virials))
    na, nb = ncoset(t["la_max"]), ncoset(t["lb_max"])
    ref = np.zeros((n2, n1))
    ref[t["o2"]:t["o2"] + nb, t["o1"]:t["o1"] + na] = t["hab"]
    worst = rel_diff(hab.reshape(n2, n1), ref)
    worst = max(worst, 1e-4 * rel_diff(forces.reshape(2, 3), np.stack([t["force_a"], t["force_b"]])))
    vir = virials.reshape(2, 3, 3).sum(axis=0)
    return max(worst, 1e-4 * rel_diff(vir, t["virial"]))
