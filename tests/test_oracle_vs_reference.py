"""Second pin of the oracle: against the UNMODIFIED reference (REF backend,
oracle/_ref/libgrid_ref.so) on seeded multi-task lists -- all 35 grid_func
values, tau / forces / virial, orthorhombic and triclinic cells, border masks.
The golden `.task` vectors only cover func 100 and 200.  Tolerances are the
reference validator's (src/grid/grid_task_list.c:241,367,386,410)."""
import numpy as np
import pytest

from cp2k_b200.grid_api import ALL_GRID_FUNCS, OffloadBuffer
from replay import rel_diff
from synth import make_workload


def run_collocate(lib, wl, func, pab):
    tl = wl.create(lib)
    grids = wl.new_grids()
    tl.collocate(func, pab, grids)
    tl.free()
    return [g.host.copy() for g in grids]


def run_integrate(lib, wl, tau, pab, grids, forces=True, virial=True):
    tl = wl.create(lib)
    hab = OffloadBuffer(wl.pab_len)
    f = np.zeros((wl.natoms, 3)) if forces else None
    v = np.zeros((3, 3)) if virial else None
    tl.integrate(tau, pab if forces else None, grids, hab, f, v)
    tl.free()
    return hab.host.copy(), f, v


@pytest.fixture(scope="module")
def wl_ortho():
    return make_workload(seed=11, natoms=4, max_tasks=400)


@pytest.fixture(scope="module")
def wl_general():
    return make_workload(seed=12, natoms=4, orthorhombic=False, max_tasks=300, border_mask_fraction=0.15)


@pytest.mark.parametrize("func", ALL_GRID_FUNCS)
def test_collocate_all_funcs_ortho(oracle, reference, wl_ortho, func):
    pab = wl_ortho.random_pab(3)
    ref = run_collocate(reference.load_reference(), wl_ortho, func, pab)
    ora = run_collocate(oracle, wl_ortho, func, pab)
    for a, b in zip(ora, ref):
        assert np.abs(b).max() > 0
        assert rel_diff(a, b) < 1e-12


@pytest.mark.parametrize("func", [100, 200, 412, 503, 702, 801, 902, 1003])
def test_collocate_general_and_masked(oracle, reference, wl_general, func):
    pab = wl_general.random_pab(4)
    ref = run_collocate(reference.load_reference(), wl_general, func, pab)
    ora = run_collocate(oracle, wl_general, func, pab)
    for a, b in zip(ora, ref):
        assert rel_diff(a, b) < 1e-12


@pytest.mark.parametrize("tau", [False, True])
@pytest.mark.parametrize("fv", [(False, False), (True, False), (True, True)])
@pytest.mark.parametrize("which", ["ortho", "general"])
def test_integrate(oracle, reference, wl_ortho, wl_general, tau, fv, which):
    wl = wl_ortho if which == "ortho" else wl_general
    pab = wl.random_pab(5)
    grids = wl.new_grids()
    rng = np.random.default_rng(6)
    for g in grids:
        g.host[:] = rng.normal(size=g.host.size)
    hab_r, f_r, v_r = run_integrate(reference.load_reference(), wl, tau, pab, grids, *fv)
    hab_o, f_o, v_o = run_integrate(oracle, wl, tau, pab, grids, *fv)
    assert np.abs(hab_r).max() > 0
    assert rel_diff(hab_o, hab_r) < 1e-12
    if fv[0]:
        assert rel_diff(f_o, f_r) < 1e-8
    if fv[1]:
        assert rel_diff(v_o, v_r) < 1e-8


def test_flop_counter_matches_survey_examples(oracle):
    """SURVEY.md Appendix A worked examples (points / flops of the REF loop nest)."""
    from replay import load_task, replay_single_oracle

    expect = {"ortho_density_l2200": (88, 1024), "ortho_density_l0000": (8552, 26692),
              "ortho_density_l3300": (23288, 230472), "ortho_density_l3333": (55984, 945712)}
    for name, (pts, flops) in expect.items():
        oracle.reset_counters()
        replay_single_oracle(oracle, load_task(name), True)
        c = oracle.counters()
        assert c["npts"] == pts
        assert c["flops"] == flops
