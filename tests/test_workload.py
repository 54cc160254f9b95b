"""Host-side workload synthesiser (cp2k_b200/workload.py): the pieces restated
from CP2K's task-list generation, checked on CPU."""
import math

import numpy as np
import pytest

from cp2k_b200 import workload as W


def test_c2s_is_orthonormal_on_normalised_cartesians():
    """c2s maps normalised Cartesians to real solid harmonics; rows are orthogonal
    w.r.t. the Cartesian overlap of a shell (l = 0..3)."""
    for l in range(4):
        c2s = W.c2s_matrix(l)
        assert c2s.shape == (2 * l + 1, W.nco(l))
        orbs = W.cart_orbitals(l)

        def ovl(a, b):  # angular overlap of normalised Cartesian monomials
            s = [x + y for x, y in zip(a, b)]
            if any(v % 2 for v in s):
                return 0.0
            num = np.prod([W._dfac(v - 1) for v in s])
            den = math.sqrt(np.prod([W._dfac(2 * v - 1) for v in a]) * np.prod([W._dfac(2 * v - 1) for v in b]))
            return num / den

        S = np.array([[ovl(a, b) for b in orbs] for a in orbs])
        G = c2s @ S @ c2s.T
        assert np.allclose(G, np.diag(np.diag(G)), atol=1e-12)
        assert np.allclose(np.diag(G), np.diag(G)[0], rtol=1e-12)


def test_exp_radius_brackets_the_threshold():
    rng = np.random.default_rng(0)
    for l in range(4):
        a = np.exp(rng.uniform(-2, 2, 50))
        pre = np.exp(rng.uniform(-3, 1, 50))
        r = W.exp_radius(l, a, 1e-8, pre, epsabs=1e-4)
        g = lambda x: pre * x ** l * np.exp(-a * x * x)
        ok = r > 0
        assert np.all(g(r[ok]) < 1e-8 * 1.01)            # below threshold at the radius
        assert np.all(g(np.maximum(r[ok] - 2e-4, 0)) >= 1e-8 * 0.5)  # and tight


def test_fft_grid_sizes_match_survey():
    """126/75/42/25 for H2O-64, 200/120/70/40 for H2O-256."""
    for name, expect in (("H2O-64", [126, 75, 42, 25]), ("H2O-256", [200, 120, 70, 40])):
        _, _, cell = W.load_system(name)
        got = [int(W.grid_npts(cell, 0.5 * 280.0 / 3.0 ** i)[0]) for i in range(4)]
        assert got == expect


def test_h2o64_task_list_statistics():
    wl = W.build_h2o_workload("H2O-64")
    assert wl.natoms == 192 and wl.orthorhombic
    assert 250_000 < wl.ntasks < 450_000          # SURVEY Appendix D estimate: 3.6e5 +- 30 %
    lv = np.bincount(wl.tasks["level_list"], minlength=5)[1:]
    assert np.all(lv > 0) and lv.sum() == wl.ntasks
    # exp_radius returns 0 when the product never exceeds the threshold; the grid
    # library then skips the task (src/grid/ref/grid_ref_collint.h:929-937)
    assert wl.tasks["radius_list"].min() >= 0
    assert np.mean(wl.tasks["radius_list"] == 0) < 0.05
    # every task points at a valid block; blocks are row<=col atom pairs
    assert wl.tasks["block_num_list"].min() >= 1 and wl.tasks["block_num_list"].max() <= wl.nblocks
    assert np.all(np.diff(wl.block_offsets) > 0)
    # basis shapes of TZV2P-GTH: H nsgf 9 / maxco 8, O nsgf 22 / maxco 20
    h, o = wl.basis_sets
    assert (h.nsgf, h.maxco, o.nsgf, o.maxco) == (9, 8, 22, 20)


def test_nonortho_cell_is_triclinic():
    _, _, cell = W.load_system("H2O-64_nonortho")
    assert not np.allclose(cell, np.diag(np.diag(cell)))
    assert abs(np.linalg.norm(cell[0]) - 12.4138 * W.ANGSTROM) < 1e-9


def test_compact_subset_is_the_same_task_list(oracle):
    """Workload.subset(compact_blocks=True) -- a rank's share in a distributed run --
    renumbers the blocks it keeps; with the matching P blocks the collocated
    density is identical to the uncompacted subset's."""
    from cp2k_b200.grid_api import OffloadBuffer
    from cp2k_b200.workload import build_h2o_workload

    w = build_h2o_workload("H2O-64", max_atoms=24)
    keep = (w.tasks["block_num_list"] % 2) == 0
    a, b = w.subset(keep), w.subset(keep, compact_blocks=True)
    assert b.nblocks < a.nblocks and b.pab_len < a.pab_len and a.ntasks == b.ntasks
    sizes = np.diff(np.append(w.block_offsets.astype(np.int64), w.pab_len))
    used = np.unique(a.tasks["block_num_list"] - 1)
    pa = a.random_pab(1)
    pb = OffloadBuffer(b.pab_len)
    pos = 0
    for u in used:
        pb.host[pos:pos + sizes[u]] = pa.host[w.block_offsets[u]:w.block_offsets[u] + sizes[u]]
        pos += sizes[u]

    def collocate(wl, p):
        tl = wl.create(oracle)
        g = wl.new_grids()
        tl.collocate(100, p, g)
        tl.free()
        return np.concatenate([x.host for x in g])

    ga, gb = collocate(a, pa), collocate(b, pb)
    assert np.abs(ga).max() > 0 and np.array_equal(ga, gb)
