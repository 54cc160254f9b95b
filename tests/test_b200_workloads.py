"""GPU parity on realistic task lists (water benchmarks, TZV2P-GTH) and on
distributed (z-slab) grid layouts, plus size-independent properties at the full
benchmark size: collocate and integrate are adjoint linear maps,
    sum_blocks w_b <P_b, H_b(V)>  ==  <rho(P), V>,   w_b = 1 (same atom) or 2,
(the factor is the reference's rscale, src/grid/ref/grid_ref_task_list.c:369)."""
import numpy as np
import pytest

from cp2k_b200.grid_api import GRID_BACKEND_CPU, OffloadBuffer
from cp2k_b200 import rsgrid
from cp2k_b200.workload import build_h2o_workload
from replay import assert_parity, rel_diff
from synth import make_workload

pytestmark = pytest.mark.gpu


def _run(lib, wl, pab, forces=True, func=100, tau=False):
    tl = wl.create(lib)
    grids = wl.new_grids()
    tl.collocate(func, pab, grids)
    hab = OffloadBuffer(wl.pab_len)
    f = np.zeros((wl.natoms, 3)) if forces else None
    v = np.zeros((3, 3)) if forces else None
    tl.integrate(tau, pab if forces else None, grids, hab, f, v)
    tl.free()
    return [g.host.copy() for g in grids], hab.host.copy(), f, v


@pytest.fixture(scope="module")
def h2o_small():
    return build_h2o_workload("H2O-64", max_atoms=36)


def test_h2o_subset_against_reference_cpu_backend(b200, reference, h2o_small):
    wl = h2o_small
    assert wl.ntasks > 5000
    pab = wl.random_pab(11)
    ref = _run(reference.load_reference(GRID_BACKEND_CPU), wl, pab)
    got = _run(b200, wl, pab)
    for a, b in zip(got[0], ref[0]):
        assert rel_diff(a, b) < 1e-10
    assert rel_diff(got[1], ref[1]) < 1e-10
    assert rel_diff(got[2], ref[2]) < 1e-8 and rel_diff(got[3], ref[3]) < 1e-8


@pytest.mark.parametrize("func,tau", [(200, True), (502, False)])
def test_h2o_subset_tau_and_gradient(b200, oracle, h2o_small, func, tau):
    wl = h2o_small.subset(np.arange(0, h2o_small.ntasks, 7))
    pab = wl.random_pab(12)
    ref = _run(oracle, wl, pab, forces=False, func=func, tau=tau)
    got = _run(b200, wl, pab, forces=False, func=func, tau=tau)
    for a, b in zip(got[0], ref[0]):
        assert rel_diff(a, b) < 1e-10
    assert rel_diff(got[1], ref[1]) < 1e-10


@pytest.mark.parametrize("world", [2, 3])
def test_slab_layouts_every_rank(b200, oracle, world):
    """npts_local != npts_global with halos: every rank's local task list."""
    wl = make_workload(seed=51, natoms=6, max_tasks=700, edge=(9.0, 10.0, 18.0),
                       npts_list=((40, 45, 80), (20, 24, 40)))
    levels = rsgrid.make_slab_levels(wl, world)
    assert any(l.distributed for l in levels)
    pab = wl.random_pab(13)
    for rank in range(world):
        mine = rsgrid.local_workload(wl, levels, rank, world)
        ref = _run(oracle, mine, pab, forces=False)
        got = _run(b200, mine, pab, forces=False)
        for a, b in zip(got[0], ref[0]):
            assert rel_diff(a, b) < 1e-10
        assert rel_diff(got[1], ref[1]) < 1e-10


def _adjoint_defect(lib, wl, seed):
    tl = wl.create(lib)
    pab = wl.random_pab(seed)
    rho = wl.new_grids()
    tl.collocate(100, pab, rho)
    rng = np.random.default_rng(seed + 1)
    pot = wl.new_grids()
    for g in pot:
        g.host[:] = rng.normal(size=g.host.size)
    hab = OffloadBuffer(wl.pab_len)
    tl.integrate(False, None, pot, hab)
    # block weights: a block (atom pair) is diagonal iff its tasks have iatom == jatom
    diag = np.zeros(wl.nblocks, dtype=bool)
    t = wl.tasks
    diag[t["block_num_list"][t["iatom_list"] == t["jatom_list"]] - 1] = True
    sizes = np.diff(np.append(wl.block_offsets, wl.pab_len))
    w = np.repeat(np.where(diag, 1.0, 2.0), sizes)
    lhs = float(np.sum(w * pab.host * hab.host))
    rhs = float(sum(np.dot(r.host, p.host) for r, p in zip(rho, pot)))
    # linearity on the way: collocate(2 P) == 2 collocate(P)
    pab2 = OffloadBuffer(wl.pab_len)
    pab2.host[:] = 2.0 * pab.host
    rho2 = wl.new_grids()
    tl.collocate(100, pab2, rho2)
    lin = max(rel_diff(a.host, 2.0 * b.host) for a, b in zip(rho2, rho))
    tl.free()
    scale = float(np.sqrt(sum(np.dot(r.host, r.host) for r in rho) * sum(np.dot(p.host, p.host) for p in pot)))
    return abs(lhs - rhs) / scale, lin


def test_adjointness_h2o64_full(b200):
    defect, lin = _adjoint_defect(b200, build_h2o_workload("H2O-64"), 21)
    assert defect < 1e-12 and lin < 1e-12


def test_adjointness_h2o256_full_size(b200):
    """BASELINE.json's headline configuration, all 1.25 M tasks."""
    defect, lin = _adjoint_defect(b200, build_h2o_workload("H2O-256"), 22)
    assert defect < 1e-12 and lin < 1e-12


def test_adjointness_h2o1024_full_size(b200):
    """BASELINE.json config 5: H2O-1024, 5.0 M tasks on four levels (315/189/105/63)^3."""
    defect, lin = _adjoint_defect(b200, build_h2o_workload("H2O-1024"), 23)
    assert defect < 1e-12 and lin < 1e-12


def test_h2o256_full_against_reference_cpu_backend(b200, reference):
    """The headline configuration itself (H2O-256, 1.25 M tasks, with forces): every
    grid value and every H element against the unmodified reference CPU backend."""
    wl = build_h2o_workload("H2O-256")
    pab = wl.random_pab(17)
    ref = _run(reference.load_reference(GRID_BACKEND_CPU), wl, pab, forces=True)
    got = _run(b200, wl, pab, forces=True)
    _check_full(got, ref)


def test_h2o64_molopt_subset_against_reference_cpu_backend(b200, reference):
    """BASELINE.json config 2's basis (DZVP-MOLOPT-SR: one l = 0..2 set per oxygen, so
    every O-O product has lp = 4 -- the lp 3-4 kernel class carries the weight)."""
    wl = build_h2o_workload("H2O-64", basis="DZVP-MOLOPT-SR-GTH", max_atoms=24)
    assert wl.ntasks > 2000
    pab = wl.random_pab(14)
    ref = _run(reference.load_reference(GRID_BACKEND_CPU), wl, pab)
    got = _run(b200, wl, pab)
    for a, b in zip(got[0], ref[0]):
        assert rel_diff(a, b) < 1e-10
    assert rel_diff(got[1], ref[1]) < 1e-10
    assert rel_diff(got[2], ref[2]) < 1e-8 and rel_diff(got[3], ref[3]) < 1e-8


@pytest.mark.parametrize("func,tau", [(100, False), (200, True)])
def test_nonortho_water_metagga_forces_virial(b200, reference, func, tau):
    """BASELINE.json config 4: triclinic cell (general-cell path), density and
    tau collocation, integrate with forces + virial (lp up to 9 with tau)."""
    full = build_h2o_workload("H2O-64_nonortho", max_atoms=18)
    assert not full.orthorhombic
    wl = full.subset(np.arange(0, full.ntasks, 3))
    pab = wl.random_pab(15)
    ref = _run(reference.load_reference(GRID_BACKEND_CPU), wl, pab, forces=True, func=func, tau=tau)
    got = _run(b200, wl, pab, forces=True, func=func, tau=tau)
    for a, b in zip(got[0], ref[0]):
        assert rel_diff(a, b) < 1e-10
    assert rel_diff(got[1], ref[1]) < 1e-10
    assert rel_diff(got[2], ref[2]) < 1e-8 and rel_diff(got[3], ref[3]) < 1e-8


def _check_full(got, ref):
    for lvl, (a, b) in enumerate(zip(got[0], ref[0])):
        assert_parity(a, b, 1e-10, f"grid level {lvl}")
    assert_parity(got[1], ref[1], 1e-10, "hab blocks")
    assert_parity(got[2], ref[2], 1e-8, "forces")
    assert_parity(got[3], ref[3], 1e-8, "virial")


def test_h2o1024_full_against_reference_cpu_backend(b200, reference):
    """BASELINE.json config 5 itself: H2O-1024, all 5.0 M tasks on (315/189/105/63)^3 grids, with
    forces + virial -- every grid value and every H element against the unmodified reference CPU
    backend (element-wise and norm-wise), not only the adjointness property."""
    wl = build_h2o_workload("H2O-1024")
    assert wl.ntasks > 4_000_000
    pab = wl.random_pab(18)
    ref = _run(reference.load_reference(GRID_BACKEND_CPU), wl, pab, forces=True)
    got = _run(b200, wl, pab, forces=True)
    _check_full(got, ref)


@pytest.mark.parametrize("func,tau", [(100, False), (200, True)])
def test_nonortho_water_full_metagga_forces_virial(b200, reference, func, tau):
    """BASELINE.json config 4 at full size: every task of the 64-molecule triclinic cell
    (general-cell path), density and tau collocation, integrate with forces + virial."""
    wl = build_h2o_workload("H2O-64_nonortho")
    assert not wl.orthorhombic and wl.ntasks > 300_000
    pab = wl.random_pab(19)
    ref = _run(reference.load_reference(GRID_BACKEND_CPU), wl, pab, forces=True, func=func, tau=tau)
    got = _run(b200, wl, pab, forces=True, func=func, tau=tau)
    _check_full(got, ref)


def test_h2o64_full_against_reference_cpu_backend(b200, reference):
    """The complete H2O-64 task list (362 k tasks: hundreds of pairs per grid block,
    several work items per block) against the unmodified reference CPU backend."""
    wl = build_h2o_workload("H2O-64")
    pab = wl.random_pab(16)
    ref = _run(reference.load_reference(GRID_BACKEND_CPU), wl, pab, forces=True)
    got = _run(b200, wl, pab, forces=True)
    _check_full(got, ref)
