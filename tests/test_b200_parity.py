"""Parity of the CUDA backend (through the C ABI) against the oracle, the
golden vectors and -- where built -- the unmodified reference REF backend.

Tolerances (north_star): relative error <= 1e-10 on grid values and hab blocks
with the reference's own measure |d|/max(1,|ref|) (src/grid/grid_task_list.c:241),
forces/virial 1e-8 (:386,:410).  The golden `.task` vectors are held to the
reference unit test's 1e-12 per cycle (src/grid/grid_unittest.c:51-56)."""
import numpy as np
import pytest

from cp2k_b200.grid_api import ALL_GRID_FUNCS, OffloadBuffer
from replay import TASK_NAMES, load_task, rel_diff, replay_batched
from synth import make_workload

pytestmark = pytest.mark.gpu

GRID_TOL, HAB_TOL, FV_TOL = 1e-10, 1e-10, 1e-8
VARIANTS = [3, 2, 1]  # 3 = CTA-tile kernels, 2 = warp-tile kernels (both with generic fallback), 1 = generic only


@pytest.fixture(params=VARIANTS, ids=["ctile", "warptile", "generic"])
def lib(b200, request):
    b200.set_kernel_variant(request.param)
    yield b200
    b200.set_kernel_variant(0)


@pytest.mark.parametrize("name", TASK_NAMES)
@pytest.mark.parametrize("collocate", [True, False], ids=["collocate", "integrate"])
def test_golden_vectors(lib, name, collocate):
    err = replay_batched(lib, load_task(name), collocate)
    assert err < 1e-12, err


@pytest.mark.parametrize("name", ["ortho_density_l3333", "ortho_density_l0122", "general_tau",
                                  "ortho_non_periodic", "general_subpatch16"])
@pytest.mark.parametrize("collocate", [True, False], ids=["collocate", "integrate"])
def test_golden_vectors_many_cycles(lib, name, collocate):
    cycles = 50
    err = replay_batched(lib, load_task(name), collocate, cycles=cycles, cycles_per_block=7)
    assert err < 1e-12 * cycles, err


def _collocate(L, wl, func, pab):
    tl = wl.create(L)
    grids = wl.new_grids()
    for g in grids:
        g.host[:] = 42.0  # overwrite semantics
    tl.collocate(func, pab, grids)
    tl.free()
    return [g.host.copy() for g in grids]


def _integrate(L, wl, tau, pab, grids, forces, virial):
    tl = wl.create(L)
    hab = OffloadBuffer(wl.pab_len)
    hab.host[:] = 42.0
    f = np.full((wl.natoms, 3), 42.0) if forces else None
    v = np.full((3, 3), 42.0) if virial else None
    tl.integrate(tau, pab if forces else None, grids, hab, f, v)
    tl.free()
    return hab.host.copy(), f, v


@pytest.fixture(scope="module")
def wl_ortho():
    return make_workload(seed=21, natoms=5, max_tasks=600)


@pytest.fixture(scope="module")
def wl_general():
    return make_workload(seed=22, natoms=4, orthorhombic=False, max_tasks=300, border_mask_fraction=0.15)


@pytest.fixture(scope="module")
def wl_masked_ortho():
    return make_workload(seed=23, natoms=4, orthorhombic=True, max_tasks=300, border_mask_fraction=0.3)


@pytest.mark.parametrize("func", ALL_GRID_FUNCS)
def test_collocate_all_funcs(lib, oracle, wl_ortho, func):
    pab = wl_ortho.random_pab(3)
    ref = _collocate(oracle, wl_ortho, func, pab)
    got = _collocate(lib, wl_ortho, func, pab)
    for a, b in zip(got, ref):
        assert np.abs(b).max() > 0
        assert rel_diff(a, b) < GRID_TOL


@pytest.mark.parametrize("func", [100, 200, 413, 502, 703, 803, 905, 1001])
@pytest.mark.parametrize("which", ["general", "masked_ortho"])
def test_collocate_general(lib, oracle, wl_general, wl_masked_ortho, func, which):
    wl = wl_general if which == "general" else wl_masked_ortho
    pab = wl.random_pab(4)
    ref = _collocate(oracle, wl, func, pab)
    got = _collocate(lib, wl, func, pab)
    for a, b in zip(got, ref):
        assert rel_diff(a, b) < GRID_TOL


@pytest.mark.parametrize("tau", [False, True], ids=["notau", "tau"])
@pytest.mark.parametrize("fv", [(False, False), (True, False), (True, True)], ids=["hab", "forces", "virial"])
@pytest.mark.parametrize("which", ["ortho", "general", "masked_ortho"])
def test_integrate(lib, oracle, wl_ortho, wl_general, wl_masked_ortho, tau, fv, which):
    wl = {"ortho": wl_ortho, "general": wl_general, "masked_ortho": wl_masked_ortho}[which]
    pab = wl.random_pab(5)
    grids = wl.new_grids()
    rng = np.random.default_rng(6)
    for g in grids:
        g.host[:] = rng.normal(size=g.host.size)
    hab_r, f_r, v_r = _integrate(oracle, wl, tau, pab, grids, *fv)
    hab_g, f_g, v_g = _integrate(lib, wl, tau, pab, grids, *fv)
    assert np.abs(hab_r).max() > 0
    assert rel_diff(hab_g, hab_r) < HAB_TOL
    if fv[0]:
        assert rel_diff(f_g, f_r) < FV_TOL
    if fv[1]:
        assert rel_diff(v_g, v_r) < FV_TOL


def test_against_unmodified_reference(lib, reference, wl_ortho):
    """Same task list through the reference's own REF and CPU backends."""
    from cp2k_b200.grid_api import GRID_BACKEND_CPU, GRID_BACKEND_REF

    pab = wl_ortho.random_pab(7)
    got = _collocate(lib, wl_ortho, 100, pab)
    for backend in (GRID_BACKEND_REF, GRID_BACKEND_CPU):
        ref = _collocate(reference.load_reference(backend), wl_ortho, 100, pab)
        for a, b in zip(got, ref):
            assert rel_diff(a, b) < GRID_TOL
    grids = wl_ortho.new_grids()
    for g, v in zip(grids, got):
        g.host[:] = v
    hab_r, f_r, v_r = _integrate(reference.load_reference(GRID_BACKEND_REF), wl_ortho, False, pab, grids, True, True)
    hab_g, f_g, v_g = _integrate(lib, wl_ortho, False, pab, grids, True, True)
    assert rel_diff(hab_g, hab_r) < HAB_TOL and rel_diff(f_g, f_r) < FV_TOL and rel_diff(v_g, v_r) < FV_TOL


def test_empty_task_list_and_handle_reuse(b200):
    """grid_task_list.c:65-70,185-189,290-311: empty lists zero the outputs; a
    non-NULL handle is reused by create."""
    wl = make_workload(seed=31, natoms=3, max_tasks=50)
    empty = {k: v[:0] for k, v in wl.tasks.items()}
    tl = b200.create_task_list(orthorhombic=True, natoms=wl.natoms, block_offsets=wl.block_offsets,
                               atom_positions=wl.atom_positions, atom_kinds=wl.atom_kinds,
                               basis_sets=wl.basis_sets, layouts=wl.layouts, **empty)
    grids = wl.new_grids()
    for g in grids:
        g.host[:] = 1.0
    tl.collocate(100, wl.random_pab(), grids)
    assert all(np.all(g.host == 0.0) for g in grids)
    hab = OffloadBuffer(wl.pab_len)
    hab.host[:] = 1.0
    f, v = np.ones((wl.natoms, 3)), np.ones((3, 3))
    tl.integrate(False, wl.random_pab(), grids, hab, f, v)
    assert np.all(hab.host == 0) and np.all(f == 0) and np.all(v == 0)
    tl.free()


def test_device_resident_mode_matches_host_mode(b200, oracle, wl_ortho):
    """device_buffer authoritative, no copies in the call."""
    import torch

    pab_h = wl_ortho.random_pab(9)
    ref = _collocate(oracle, wl_ortho, 100, pab_h)
    tl = wl_ortho.create(b200)
    pab = OffloadBuffer.with_device(wl_ortho.pab_len)
    pab.device.copy_(torch.from_numpy(pab_h.host))
    grids = [OffloadBuffer.with_device(l.npts_local_total) for l in wl_ortho.layouts]
    hab = OffloadBuffer.with_device(wl_ortho.pab_len)
    b200.set_device_resident(True)
    try:
        tl.collocate(100, pab, grids)
        tl.integrate(False, None, grids, hab)
        torch.cuda.synchronize()
    finally:
        b200.set_device_resident(False)
    for g, r in zip(grids, ref):
        assert rel_diff(g.device.cpu().numpy()[: r.size], r) < GRID_TOL
    gh = wl_ortho.new_grids()
    for g, r in zip(gh, ref):
        g.host[:] = r
    hab_r, _, _ = _integrate(oracle, wl_ortho, False, pab_h, gh, False, False)
    assert rel_diff(hab.device.cpu().numpy()[: hab_r.size], hab_r) < HAB_TOL
    tl.free()


def test_resident_mode_with_host_only_buffers(b200, oracle, wl_ortho):
    """Mixed mode: the resident flag is on, but P and H are pinned host-only buffers (NULL
    device_buffer) -- the library copies them inside the call and must not return before the copies
    are done: host P may be changed right after collocate returns, host H is read right after
    integrate returns, without any synchronisation by the caller.  Grids once device-authoritative,
    once host-only as well."""
    import torch

    wl = wl_ortho
    pab_h = wl.random_pab(9)
    ref = _collocate(oracle, wl, 100, pab_h)
    gh = wl.new_grids()
    for g, r in zip(gh, ref):
        g.host[:] = r
    hab_r, _, _ = _integrate(oracle, wl, False, pab_h, gh, False, False)
    tl = wl.create(b200)
    pab = OffloadBuffer(wl.pab_len, pinned=True)
    hab = OffloadBuffer(wl.pab_len, pinned=True)
    grids_dev = [OffloadBuffer.with_device(l.npts_local_total) for l in wl.layouts]
    grids_host = [OffloadBuffer(l.npts_local_total, pinned=True) for l in wl.layouts]
    b200.set_device_resident(True)
    try:
        for grids in (grids_dev, grids_host):
            pab.host[:] = pab_h.host
            hab.host[:] = np.nan
            tl.collocate(100, pab, grids)
            pab.host[:] = 0.0  # the caller owns host P again
            tl.integrate(False, None, grids, hab)
            assert rel_diff(hab.host[: hab_r.size].copy(), hab_r) < HAB_TOL
            for g, r in zip(grids, ref):
                got = g.device.cpu().numpy() if g.device is not None else g.host
                assert rel_diff(got[: r.size], r) < GRID_TOL
    finally:
        b200.set_device_resident(False)
    tl.free()


def test_stats_match_oracle_counters(b200, oracle, wl_ortho, wl_general):
    for wl in (wl_ortho, wl_general):
        tl = wl.create(b200)
        st = b200.stats(tl)
        tl.free()
        oracle.reset_counters()
        _collocate(oracle, wl, 100, wl.random_pab())
        c = oracle.counters()
        assert st["npts_model"] == c["npts"]
        assert abs(st["flops_collocate"] - c["flops"]) <= 1e-9 * c["flops"]


def _lp0_of(wl):
    t, kinds = wl.tasks, wl.atom_kinds
    lmax = [np.asarray(b.lmax) for b in wl.basis_sets]
    la = np.array([lmax[kinds[a - 1] - 1][s - 1] for a, s in zip(t["iatom_list"], t["iset_list"])])
    lb = np.array([lmax[kinds[a - 1] - 1][s - 1] for a, s in zip(t["jatom_list"], t["jset_list"])])
    return la + lb


@pytest.mark.parametrize("lp0", [0, 1, 2, 3, 4])
def test_multi_pair_items(b200, oracle, wl_ortho, lp0):
    """Several (task, block) pairs per work item in every lp-specialised pair
    loop of the tiled kernels: the same task repeated (identical geometry, so
    every block sees it several times) plus a few different ones, for every l
    growth the API can ask for (collocate dl = 0, 1, 2; integrate dl = 0..3)."""
    ids = np.nonzero(_lp0_of(wl_ortho) == lp0)[0]
    if ids.size == 0:
        pytest.skip("no such tasks in the synthetic workload")
    sel = np.array([ids[0]] * 3 + list(ids[:5]))
    wl = wl_ortho.subset(sel)
    pab = wl.random_pab(8)
    for func in (100, 301, 200):
        ref = _collocate(oracle, wl, func, pab)
        got = _collocate(b200, wl, func, pab)
        for a, b in zip(got, ref):
            assert rel_diff(a, b) < GRID_TOL, (func, lp0)
    grids = wl.new_grids()
    rng = np.random.default_rng(9)
    for g in grids:
        g.host[:] = rng.normal(size=g.host.size)
    for tau in (False, True):
        for fv in ((False, False), (True, False), (True, True)):
            hab_r, f_r, v_r = _integrate(oracle, wl, tau, pab, grids, *fv)
            hab_g, f_g, v_g = _integrate(b200, wl, tau, pab, grids, *fv)
            assert rel_diff(hab_g, hab_r) < HAB_TOL, (tau, fv, lp0)
            if fv[0]:
                assert rel_diff(f_g, f_r) < FV_TOL
            if fv[1]:
                assert rel_diff(v_g, v_r) < FV_TOL
