"""Synthesises the task lists CP2K's `generate_qs_task_list` would hand to
`grid_create_task_list` for the water benchmarks (benchmarks/QS/H2O-N.inp),
without a Fortran compiler -- the "caller side" of the hot path.

What is restated (all read-only Fortran in /root/reference/src, SURVEY.md
Appendix C):
  thresholds            cp_control_utils.F:1049-1068   EPS_DEFAULT -> eps_pgf_orb, eps_rho_rspace
  basis normalisation   aobasis/basis_set_types.F:1182-1202 (normalise_gcc_orb), :1056-1098
                        (init_norm_cgf_orb), :817-890 (cphi/sphi), orbital_transformation_matrices.F:118-160
  primitive radii       qs_interactions.F:474-491, aobasis/ao_util.F:95-180 (exp_radius)
  neighbour pairs       qs_neighbor_lists.F:1408-1446 (symmetric checkerboard rule, all images)
  pair screening        task_list_methods.F:680-717 (task_list_inner_loop)
  grid level            pw_env/gaussian_gridlevels.F:147-167
  task radius           task_list_methods.F:776-783, aobasis/ao_util.F:207-290
  grid sizes            pw/pw_grid_info.F:222-238, pw/fft/fftw3_lib.F:255-264
  cell                  cell_methods.F:671-711
  blocks                task_list_methods.F:2091-2128, 2439-2484 (one block per atom pair, row<=col)
Input data (coordinates, exponents, contraction coefficients) come from the
fixtures under cp2k_b200/data written by tools/extract_benchmark_data.py.

The density matrix is synthetic: seeded, symmetric, decaying with distance
(values do not change the cost; the decay keeps grid values O(1)).
"""
from __future__ import annotations

import json
import math
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

from .grid_api import BasisSet, GridLayout, OffloadBuffer

ANGSTROM = 1.0 / 0.52917721067  # bohr per angstrom (CP2K's CODATA 2014 value, physcon.F)
_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def ncoset(l: int) -> int:
    return (l + 1) * (l + 2) * (l + 3) // 6 if l >= 0 else 0


def nco(l: int) -> int:
    return (l + 1) * (l + 2) // 2


def _dfac(n: int) -> float:  # double factorial, dfac(-1) = 1
    r = 1.0
    while n > 1:
        r *= n
        n -= 2
    return r


def cart_orbitals(l: int):
    """Cartesian (lx,ly,lz) of one shell in CP2K order (co index)."""
    return [(lx, ly, l - lx - ly) for lx in range(l, -1, -1) for ly in range(l - lx, -1, -1)]


def c2s_matrix(l: int) -> np.ndarray:
    """Cartesian -> real solid harmonics, [nso(l)][nco(l)]
    (orbital_transformation_matrices.F:118-160)."""
    fac = math.factorial
    comb = lambda n, k: math.comb(n, k) if 0 <= k <= n else 0
    orbs = cart_orbitals(l)
    out = np.zeros((2 * l + 1, len(orbs)))
    for ic, (lx, ly, lz) in enumerate(orbs):
        for m in range(-l, l + 1):
            ma = abs(m)
            j = lx + ly - ma
            if j < 0 or j % 2:
                continue
            j //= 2
            s1 = 0.0
            for i in range((l - ma) // 2 + 1):
                s2 = 0.0
                for k in range(j + 1):
                    if (m < 0 and abs(ma - lx) % 2 == 1) or (m > 0 and abs(ma - lx) % 2 == 0):
                        s = (-1.0) ** ((ma - lx + 2 * k) // 2) * math.sqrt(2.0)
                    elif m == 0 and lx % 2 == 0:
                        s = (-1.0) ** (k - lx // 2)
                    else:
                        s = 0.0
                    s2 += comb(j, k) * comb(ma, lx - 2 * k) * s
                s1 += comb(l, i) * comb(i, j) * (-1.0) ** i * fac(2 * l - 2 * i) / fac(l - ma - 2 * i) * s2
            out[l + m, ic] = math.sqrt(
                (fac(2 * lx) * fac(2 * ly) * fac(2 * lz) * fac(l) * fac(l - ma))
                / (fac(lx) * fac(ly) * fac(lz) * fac(2 * l) * fac(l + ma))) * s1 / (2.0 ** l * fac(l))
    return out


# ----------------------------------------------------------------------------
# exp_radius (ao_util.F:95-180), vectorised over numpy arrays
# ----------------------------------------------------------------------------
_C38, _C62 = float(np.float32(0.38)), float(np.float32(0.62))


def _exp_sp(x):
    """EXP(REAL(x, KIND=sp)) promoted back to double."""
    return np.exp(np.asarray(x, dtype=np.float32)).astype(np.float64)


def exp_radius(l: int, alpha, threshold: float, prefactor, epsabs: Optional[float] = None, rlow=None):
    a = np.abs(np.asarray(alpha, dtype=np.float64))
    d = np.abs(np.asarray(prefactor, dtype=np.float64))
    a, d = np.broadcast_arrays(a, d)
    t = abs(threshold)
    radius = np.zeros(a.shape) if rlow is None else np.array(np.broadcast_to(rlow, a.shape), dtype=np.float64)
    r = np.maximum(np.sqrt(0.5 * l / a), radius)
    g = d.copy()
    if l != 0:
        g = g * _exp_sp(-a * r * r) * r ** l
    active = (d != 0.0) & ~(g < t)  # others return rlow unchanged
    if not np.any(active):
        return radius
    radius = np.where(active, r * 2.0 + 1.0, radius)
    grow = active.copy()
    for _ in range(200):  # bracket
        g = d * _exp_sp(-a * radius * radius) * radius ** l
        grow &= ~(g < t)
        if not np.any(grow):
            break
        r = np.where(grow, radius, r)
        radius = np.where(grow, r * 2.0 + 1.0, radius)
    eps = float(np.float32(1.0e-12)) if epsabs is None else epsabs
    dr = np.zeros(a.shape)
    busy = active.copy()
    for _ in range(400):  # golden-section like contraction
        rd = radius - r
        busy &= ~((rd < eps) | (rd == dr))
        if not np.any(busy):
            break
        s1, s2 = r + rd * _C38, r + rd * _C62
        g1 = d * _exp_sp(-a * s1 * s1) * s1 ** l
        g2 = d * _exp_sp(-a * s2 * s2) * s2 ** l
        c1 = g1 < t
        c2 = ~c1 & (g2 < t)
        new_radius = np.where(c1, s1, np.where(c2, s2, radius))
        new_r = np.where(c1, r, s2)          # c2: r = s1 then radius = s2 ...
        new_r = np.where(c2, s1, new_r)      # ... neither: r = s2
        radius = np.where(busy, new_radius, radius)
        r = np.where(busy, new_r, r)
        dr = np.where(busy, rd, dr)
    return radius


def exp_radius_very_extended(la_max: int, lb_max: int, rad_a, rad_b, zetp, eps: float, prefactor):
    """ao_util.F:207-290 without pab screening (cutoff = 1), vectorised."""
    rad_a, rad_b = np.asarray(rad_a), np.asarray(rad_b)
    lp = la_max + lb_max
    polycoef = np.zeros((lp + 1,) + rad_a.shape)
    for lxa in range(la_max + 1):
        for lxb in range(lb_max + 1):
            coef = np.zeros_like(polycoef)
            bini, s1 = 1.0, np.ones_like(rad_a)
            for i in range(lxa + 1):
                binj, s2 = 1.0, np.ones_like(rad_b)
                for j in range(lxb + 1):
                    coef[lxa + lxb - i - j] += bini * binj * s1 * s2
                    binj = (binj * (lxb - j)) / (j + 1)
                    s2 = s2 * rad_b
                bini = (bini * (lxa - i)) / (i + 1)
                s1 = s1 * rad_a
            polycoef = np.maximum(polycoef, coef)
    polycoef = polycoef * (np.asarray(prefactor) * 1.0)
    radius = np.zeros(rad_a.shape)
    for i in range(lp + 1):
        radius = np.maximum(radius, exp_radius(i, zetp, eps, polycoef[i], epsabs=1.0e-2, rlow=radius))
    return radius


# ----------------------------------------------------------------------------
# basis sets
# ----------------------------------------------------------------------------
@dataclass
class KindBasis:
    """A CP2K orbital basis after init_orb_basis_set, in the layout
    grid_create_basis_set expects (grid_api.F:631-647)."""

    grid: BasisSet
    pgf_radius: np.ndarray  # [nset][maxpgf]
    set_radius: np.ndarray
    kind_radius: float


def make_kind_basis(sets: List[dict], eps_pgf_orb: float) -> KindBasis:
    nset = len(sets)
    lmin = [s["lmin"] for s in sets]
    lmax = [s["lmax"] for s in sets]
    npgf = [len(s["zet"]) for s in sets]
    maxpgf = max(npgf)
    maxco = max(n * ncoset(l) for n, l in zip(npgf, lmax))
    nsgf_set, shells = [], []
    for s in sets:
        ls = [l for l, n in zip(range(s["lmin"], s["lmax"] + 1), s["nshell"]) for _ in range(n)]
        shells.append(ls)
        nsgf_set.append(sum(2 * l + 1 for l in ls))
    nsgf = sum(nsgf_set)
    first_sgf = np.cumsum([1] + nsgf_set[:-1])
    zet = np.zeros((nset, maxpgf))
    sphi = np.zeros((nsgf, maxco))
    pgf_radius = np.zeros((nset, maxpgf))
    for iset, s in enumerate(sets):
        z = np.array(s["zet"])
        zet[iset, : len(z)] = z
        coef = np.array(s["coef"])  # [npgf][nshell_total]
        sgf = first_sgf[iset] - 1
        for ish, l in enumerate(shells[iset]):
            # normalise_gcc_orb: primitive normalisation folded into gcc
            gcc = 2.0 ** l * (2.0 / math.pi) ** 0.75 * z ** (0.25 * (2 * l + 3)) * coef[:, ish]
            # init_norm_cgf_orb: normalise the contracted function
            fnorm = 0.5 ** l * math.pi ** 1.5 * np.sum(
                gcc[:, None] * gcc[None, :] / (z[:, None] + z[None, :]) ** (0.5 * (2 * l + 3)))
            orbs = cart_orbitals(l)
            norm_cgf = np.array([1.0 / math.sqrt(_dfac(2 * lx - 1) * _dfac(2 * ly - 1) * _dfac(2 * lz - 1) * fnorm)
                                 for lx, ly, lz in orbs])
            # cphi[ico][icgf] = norm_cgf * gcc ; sphi = cphi * c2s^T
            c2s = c2s_matrix(l)  # [nso][nco]
            for ipgf in range(len(z)):
                base = ipgf * ncoset(lmax[iset]) + ncoset(l - 1)
                cphi = norm_cgf * gcc[ipgf]  # per Cartesian function of the shell
                sphi[sgf: sgf + 2 * l + 1, base: base + nco(l)] = c2s * cphi[None, :]
            sgf += 2 * l + 1
            # qs_interactions.F:474-491
            pgf_radius[iset, : len(z)] = np.maximum(
                pgf_radius[iset, : len(z)],
                exp_radius(l, z, eps_pgf_orb, gcc, rlow=pgf_radius[iset, : len(z)]))
    set_radius = pgf_radius.max(axis=1)
    grid = BasisSet(lmin, lmax, npgf, nsgf_set, first_sgf, sphi, zet)
    return KindBasis(grid, pgf_radius, set_radius, float(set_radius.max()))


def load_basis(element: str, name: str, eps_pgf_orb: float) -> KindBasis:
    data = json.load(open(os.path.join(_DATA, "basis_sets.json")))
    return make_kind_basis(data[f"{element}:{name}"], eps_pgf_orb)


# ----------------------------------------------------------------------------
# cell, grids
# ----------------------------------------------------------------------------
def make_cell(abc_angstrom, angles_deg=(90.0, 90.0, 90.0)) -> np.ndarray:
    """hmat with lattice vectors as ROWS, in bohr (cell_methods.F:685-700)."""
    al, be, ga = (math.radians(x) for x in angles_deg)
    snap = lambda v: 0.0 if abs(v) < 1e-12 else (math.copysign(1.0, v) if abs(abs(v) - 1) < 1e-12 else v)
    cg, sg, cb, ca = snap(math.cos(ga)), snap(math.sin(ga)), snap(math.cos(be)), snap(math.cos(al))
    v1 = np.array([1.0, 0.0, 0.0])
    v2 = np.array([cg, sg, 0.0])
    v3 = np.array([cb, (ca - cg * cb) / sg, 0.0])
    v3[2] = math.sqrt(1.0 - v3[0] ** 2 - v3[1] ** 2)
    abc = np.asarray(abc_angstrom, dtype=np.float64) * ANGSTROM
    return np.stack([v1 * abc[0], v2 * abc[1], v3 * abc[2]])


def _fft_sizes(limit=4096):
    out = set()
    for a in range(16):
        for b in range(4):
            for c in range(3):
                for d in range(2):
                    for e in range(2):
                        n = 2 ** a * 3 ** b * 5 ** c * 7 ** d * 11 ** e
                        if n <= limit:
                            out.add(n)
    return sorted(out)


_FFT_SIZES = _fft_sizes()


def grid_npts(cell: np.ndarray, cutoff_ha: float) -> np.ndarray:
    """pw_grid_n_from_cutoff (pw_grid_info.F:222-238) rounded up to the next
    FFTW-friendly size (fftw3_lib.F:255-264)."""
    alat = np.sum(cell ** 2, axis=1)
    n0 = 2 * np.floor(np.sqrt(2.0 * cutoff_ha * alat) / (2.0 * math.pi)).astype(int) + 1
    return np.array([next(s for s in _FFT_SIZES if s >= n) for n in n0], dtype=np.int32)


# ----------------------------------------------------------------------------
@dataclass
class Workload:
    """Everything `grid_create_task_list` needs plus the buffers' sizes."""

    orthorhombic: bool
    natoms: int
    atom_positions: np.ndarray
    atom_kinds: np.ndarray
    basis_sets: List[BasisSet]
    layouts: List[GridLayout]
    block_offsets: np.ndarray
    tasks: Dict[str, np.ndarray]
    pab_len: int
    meta: dict = field(default_factory=dict)

    @property
    def ntasks(self) -> int:
        return int(self.tasks["level_list"].shape[0])

    @property
    def nblocks(self) -> int:
        return int(self.block_offsets.shape[0])

    def create(self, lib):
        return lib.create_task_list(
            orthorhombic=self.orthorhombic, natoms=self.natoms, block_offsets=self.block_offsets,
            atom_positions=self.atom_positions, atom_kinds=self.atom_kinds, basis_sets=self.basis_sets,
            layouts=self.layouts, **self.tasks)

    def new_grids(self, make=OffloadBuffer):
        return [make(l.npts_local_total) for l in self.layouts]

    def random_pab(self, seed=1, make=OffloadBuffer):
        buf = make(self.pab_len)
        decay = self.meta.get("block_decay")
        rng = np.random.default_rng(seed)
        if decay is None:
            buf.host[:] = rng.normal(size=self.pab_len)
        else:
            vals = rng.normal(size=self.pab_len)
            buf.host[:] = vals * np.repeat(decay, self.meta["block_sizes"])
        return buf

    def subset(self, keep: np.ndarray, compact_blocks: bool = False) -> "Workload":
        """Same system, only the tasks selected by the boolean/index array.  With
        `compact_blocks` the matrix blocks no kept task refers to are dropped and
        the rest renumbered (what a rank owns in a distributed run: its P and H
        block buffers hold its blocks only)."""
        tasks = {k: v[keep] for k, v in self.tasks.items()}
        if not compact_blocks:
            return Workload(self.orthorhombic, self.natoms, self.atom_positions, self.atom_kinds, self.basis_sets,
                            self.layouts, self.block_offsets, tasks, self.pab_len, dict(self.meta))
        sizes = np.diff(np.append(self.block_offsets.astype(np.int64), self.pab_len))
        used = np.unique(tasks["block_num_list"] - 1)
        new_num = np.full(self.nblocks, -1, dtype=np.int64)
        new_num[used] = np.arange(used.size)
        tasks["block_num_list"] = (new_num[tasks["block_num_list"] - 1] + 1).astype(np.int32)
        new_sizes = sizes[used]
        offsets = np.concatenate([[0], np.cumsum(new_sizes)[:-1]]).astype(np.int32) if used.size else np.zeros(0, np.int32)
        meta = dict(self.meta)
        if meta.get("block_decay") is not None:
            meta["block_decay"] = np.asarray(meta["block_decay"])[used]
        meta["block_sizes"] = new_sizes
        # where the kept blocks live in the parent's P/H buffers (for block_index_map)
        meta["parent_block_ids"] = used
        meta["parent_block_offsets"] = self.block_offsets.astype(np.int64)[used]
        return Workload(self.orthorhombic, self.natoms, self.atom_positions, self.atom_kinds, self.basis_sets,
                        self.layouts, offsets, tasks, int(new_sizes.sum()), meta)


def block_index_map(sub: Workload) -> np.ndarray:
    """For a `subset(compact_blocks=True)` workload: the index in the parent's
    P/H block buffer of every element of the subset's (compacted) buffer, so that
    `sub_buf[:] = parent_buf[idx]` gathers a rank's P blocks and
    `parent_buf[idx] += sub_buf` scatters its H blocks back."""
    sizes = np.asarray(sub.meta["block_sizes"], dtype=np.int64)
    starts = np.asarray(sub.meta["parent_block_offsets"], dtype=np.int64)
    if sizes.size == 0:
        return np.zeros(0, dtype=np.int64)
    local0 = np.concatenate([[0], np.cumsum(sizes)[:-1]])
    return np.repeat(starts - local0, sizes) + np.arange(int(sizes.sum()), dtype=np.int64)


def load_system(name: str):
    z = np.load(os.path.join(_DATA, "h2o_systems.npz"))
    key = name.replace("-", "_")
    xyz = z[key + "__xyz_angstrom"] * ANGSTROM
    is_o = z[key + "__is_oxygen"]
    cell = make_cell(z[key + "__abc_angstrom"], z[key + "__alpha_beta_gamma"])
    return xyz, is_o, cell


def build_h2o_workload(name: str = "H2O-64", basis: str = "TZV2P-GTH", cutoff_ry: float = 280.0,
                       rel_cutoff_ry: float = 30.0, ngrids: int = 4, progression: float = 3.0,
                       eps_default: float = 1.0e-12, max_atoms: Optional[int] = None) -> Workload:
    """Task list of one MPI rank owning everything (replicated grids, one GPU).

    `max_atoms` truncates the system to its first atoms IN THE SAME CELL (a
    cheaper, sparser case for tests)."""
    xyz, is_o, cell = load_system(name)
    if max_atoms is not None:
        xyz, is_o = xyz[:max_atoms], is_o[:max_atoms]
    natoms = xyz.shape[0]
    orthorhombic = bool(np.allclose(cell, np.diag(np.diag(cell))))
    eps_pgf_orb = math.sqrt(eps_default)
    eps_rho_rspace = eps_default
    kinds = [load_basis("H", basis, eps_pgf_orb), load_basis("O", basis, eps_pgf_orb)]  # kind 1 = H, 2 = O
    atom_kinds = np.where(is_o, 2, 1).astype(np.int32)

    # positions folded like pbc(r, cell): fractional in [-1/2, 1/2)
    hinv = np.linalg.inv(cell)
    frac = xyz @ hinv
    frac -= np.rint(frac)
    pos = frac @ cell

    # multigrid
    cutoffs = [0.5 * cutoff_ry / progression ** i for i in range(ngrids)]  # Hartree
    rel_cutoff = 0.5 * rel_cutoff_ry
    layouts = []
    for c in cutoffs:
        npts = grid_npts(cell, c)
        dh = cell / npts[:, None].astype(np.float64)
        layouts.append(GridLayout(npts, npts.copy(), np.zeros(3, np.int32), np.zeros(3, np.int32), dh,
                                  np.linalg.inv(dh)))

    # neighbour pairs: all ordered (a,b,image) passing the checkerboard rule
    kr = np.array([kinds[k - 1].kind_radius for k in atom_kinds])
    rmax = 2.0 * kr.max()
    heights = 1.0 / np.linalg.norm(hinv, axis=0)  # distance between lattice planes
    nimg = np.ceil(rmax / heights).astype(int)
    shifts = np.array([[i, j, k] for i in range(-nimg[0], nimg[0] + 1) for j in range(-nimg[1], nimg[1] + 1)
                       for k in range(-nimg[2], nimg[2] + 1)])
    shift_vecs = shifts @ cell
    ia_all, ib_all, rab_all = [], [], []
    idx = np.arange(natoms)
    a1, b1 = idx[:, None] + 1, idx[None, :] + 1  # one-based like the Fortran rule
    allowed = np.where(a1 > b1, (a1 + b1) % 2 != 0, (a1 + b1) % 2 == 0)
    cut2 = (kr[:, None] + kr[None, :]) ** 2
    for sv in shift_vecs:
        d = pos[None, :, :] + sv[None, None, :] - pos[:, None, :]
        d2 = np.einsum("abk,abk->ab", d, d)
        m = allowed & (d2 < cut2)
        ia, ib = np.nonzero(m)
        ia_all.append(ia)
        ib_all.append(ib)
        rab_all.append(d[ia, ib])
    ia = np.concatenate(ia_all)
    ib = np.concatenate(ib_all)
    rab = np.concatenate(rab_all)
    dab = np.linalg.norm(rab, axis=1)

    # blocks: one per unordered atom pair (row = min, col = max), images share it
    row, col = np.minimum(ia, ib), np.maximum(ia, ib)
    pair_key = row.astype(np.int64) * natoms + col
    uniq, block_of_pair = np.unique(pair_key, return_inverse=True)
    nsgf = np.array([kinds[k - 1].grid.nsgf for k in atom_kinds])
    urow, ucol = uniq // natoms, uniq % natoms
    block_sizes = nsgf[urow] * nsgf[ucol]
    block_offsets = np.concatenate([[0], np.cumsum(block_sizes)[:-1]]).astype(np.int64)
    pab_len = int(block_sizes.sum())
    assert pab_len < 2 ** 31
    # decay of the synthetic density with the minimum-image distance of the pair
    dmin = np.full(uniq.shape[0], np.inf)
    np.minimum.at(dmin, block_of_pair, dab)

    # tasks, vectorised per (kind_a, kind_b, iset, jset, ipgf, jpgf)
    cols = {k: [] for k in ("level_list", "iatom_list", "jatom_list", "iset_list", "jset_list", "ipgf_list",
                            "jpgf_list", "block_num_list", "radius_list", "rab_list")}
    level_cut = np.array(cutoffs)
    for ka in (1, 2):
        for kb in (1, 2):
            sel = np.nonzero((atom_kinds[ia] == ka) & (atom_kinds[ib] == kb))[0]
            if sel.size == 0:
                continue
            A, B = kinds[ka - 1], kinds[kb - 1]
            d_sel = dab[sel]
            for iset in range(A.grid.nset):
                m_i = ~(A.set_radius[iset] + B.kind_radius < d_sel)
                for jset in range(B.grid.nset):
                    m_ij = m_i & ~(A.set_radius[iset] + B.set_radius[jset] < d_sel)
                    for ipgf in range(A.grid.npgf[iset]):
                        m_ijp = m_ij & ~(A.pgf_radius[iset, ipgf] + B.set_radius[jset] < d_sel)
                        for jpgf in range(B.grid.npgf[jset]):
                            m = m_ijp & ~(A.pgf_radius[iset, ipgf] + B.pgf_radius[jset, jpgf] < d_sel)
                            s = sel[m]
                            if s.size == 0:
                                continue
                            za, zb = A.grid.zet[iset, ipgf], B.grid.zet[jset, jpgf]
                            zetp = za + zb
                            needed = abs(zetp) * rel_cutoff
                            level = 1
                            for i in range(ngrids):
                                if level_cut[i] + 1e-6 >= needed:
                                    level = i + 1
                            f = zb / zetp
                            rab2 = dab[s] ** 2
                            prefactor = np.exp(-za * f * rab2)
                            rad_a = f * dab[s]            # |ra - rp|
                            rad_b = (1.0 - f) * dab[s]    # |rb - rp|
                            radius = exp_radius_very_extended(int(A.grid.lmax[iset]), int(B.grid.lmax[jset]),
                                                              rad_a, rad_b, zetp, eps_rho_rspace, prefactor)
                            n = s.size
                            cols["level_list"].append(np.full(n, level, np.int32))
                            cols["iatom_list"].append(ia[s] + 1)
                            cols["jatom_list"].append(ib[s] + 1)
                            cols["iset_list"].append(np.full(n, iset + 1, np.int32))
                            cols["jset_list"].append(np.full(n, jset + 1, np.int32))
                            cols["ipgf_list"].append(np.full(n, ipgf + 1, np.int32))
                            cols["jpgf_list"].append(np.full(n, jpgf + 1, np.int32))
                            cols["block_num_list"].append(block_of_pair[s] + 1)
                            cols["radius_list"].append(radius)
                            cols["rab_list"].append(rab[s])
    tasks = {k: np.concatenate(v) for k, v in cols.items()}
    tasks["border_mask_list"] = np.zeros(tasks["level_list"].shape[0], np.int32)
    # CP2K's order: level, then atom pair, then set / pgf indices (tasks_less_than, :2931-2960)
    order = np.lexsort((tasks["jpgf_list"], tasks["ipgf_list"], tasks["jset_list"], tasks["iset_list"],
                        tasks["jatom_list"], tasks["iatom_list"], tasks["level_list"]))
    tasks = {k: v[order] for k, v in tasks.items()}
    for k in ("level_list", "iatom_list", "jatom_list", "iset_list", "jset_list", "ipgf_list", "jpgf_list",
              "block_num_list", "border_mask_list"):
        tasks[k] = tasks[k].astype(np.int32)
    meta = {"name": name, "basis": basis, "cell": cell, "cutoffs_ha": cutoffs, "npairs": int(ia.size),
            "block_decay": np.exp(-dmin / 3.0), "block_sizes": block_sizes,
            "npts": [tuple(int(x) for x in l.npts_global) for l in layouts]}
    return Workload(orthorhombic, natoms, pos, atom_kinds, [k.grid for k in kinds], layouts,
                    block_offsets.astype(np.int32), tasks, pab_len, meta)
