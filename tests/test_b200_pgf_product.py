"""The ad-hoc Gaussian-product entry points (module
grid_api's collocate_pgf_product / integrate_pgf_product,
src/grid/grid_api.F:110-236, 267-490) of the CUDA backend through the C ABI:

* every golden `.task` vector replayed as ONE product, the way the reference's
  harness does (src/grid/grid_replay.c:413-441), held to the unit test's 1e-12;
* the batched form against the oracle's single-product functions on seeded
  random products (accumulate semantics, sub-block offsets, forces).
"""
import ctypes as C

import numpy as np
import pytest

from cp2k_b200.grid_api import GridLayout, _dp, _f64, _i32, _ip
from replay import TASK_NAMES, golden_grid, layout_of, load_task, ncoset, rel_diff

pytestmark = pytest.mark.gpu


def _args(t):
    return dict(orthorhombic=bool(t["orthorhombic"]), border_mask=t["border_mask"], la_max=t["la_max"],
                la_min=t["la_min"], lb_max=t["lb_max"], lb_min=t["lb_min"], zeta=t["zeta"], zetb=t["zetb"],
                layout=layout_of(t), ra=t["ra"], rab=t["rab"], radius=t["radius"], o1=t["o1"], o2=t["o2"])


@pytest.mark.parametrize("name", TASK_NAMES)
def test_golden_collocate_single_product(b200, name):
    t = load_task(name)
    grid = np.full(int(np.prod(t["npts_local"])), 0.25)  # the call accumulates
    b200.collocate_pgf_product(func=t["func"], rscale=t["rscale"], pab=t["pab"], grid=grid, **_args(t))
    assert rel_diff(grid - 0.25, golden_grid(t)) < 1e-12


@pytest.mark.parametrize("name", TASK_NAMES)
def test_golden_integrate_single_product(b200, name):
    t = load_task(name)
    n1, n2 = t["n1"], t["n2"]
    hab = np.full((n2, n1), 0.5)
    forces = np.full((2, 3), -1.0)
    b200.integrate_pgf_product(compute_tau=(t["func"] == 200), grid=golden_grid(t), hab=hab, pab=t["pab"],
                               forces=forces, **_args(t))
    na, nb = ncoset(t["la_max"]), ncoset(t["lb_max"])
    ref = np.zeros((n2, n1))
    ref[t["o2"]:t["o2"] + nb, t["o1"]:t["o1"] + na] = t["hab"]
    assert rel_diff(hab - 0.5, ref) < 1e-12
    assert 1e-4 * rel_diff(forces + 1.0, np.stack([t["force_a"], t["force_b"]])) < 1e-12
    # hab only (no pab, no forces)
    hab2 = np.zeros((n2, n1))
    b200.integrate_pgf_product(compute_tau=(t["func"] == 200), grid=golden_grid(t), hab=hab2, **_args(t))
    assert rel_diff(hab2, ref) < 1e-12


def _random_products(rng, n, layout, lmax=2):
    cell = np.asarray(layout.dh) * np.asarray(layout.npts_global)[:, None]
    prods = []
    for _ in range(n):
        la_max, lb_max = int(rng.integers(0, lmax + 1)), int(rng.integers(0, lmax + 1))
        la_min, lb_min = int(rng.integers(0, la_max + 1)), int(rng.integers(0, lb_max + 1))
        pad1, pad2 = int(rng.integers(0, 3)), int(rng.integers(0, 3))
        n1, n2 = ncoset(la_max) + pad1, ncoset(lb_max) + pad2
        prods.append(dict(
            la_max=la_max, la_min=la_min, lb_max=lb_max, lb_min=lb_min,
            zeta=float(rng.uniform(0.4, 2.5)), zetb=float(rng.uniform(0.4, 2.5)),
            rscale=float(rng.choice([1.0, 2.0])), ra=rng.uniform(0, 1, 3) @ cell,
            rab=rng.normal(0, 0.6, 3), radius=float(rng.uniform(1.2, 2.6)),
            o1=int(rng.integers(0, pad1 + 1)), o2=int(rng.integers(0, pad2 + 1)),
            pab=rng.normal(size=(n2, n1)), border_mask=0))
    return prods


def _oracle_collocate(L, ortho, func, p, lay, grid):
    L.grid_oracle_collocate_pgf_product(
        ortho, p["border_mask"], func, p["la_max"], p["la_min"], p["lb_max"], p["lb_min"], p["zeta"],
        p["zetb"], p["rscale"], _dp(_f64(lay.dh).reshape(-1)), _dp(_f64(lay.dh_inv).reshape(-1)),
        _dp(_f64(p["ra"])), _dp(_f64(p["rab"])), _ip(_i32(lay.npts_global)), _ip(_i32(lay.npts_local)),
        _ip(_i32(lay.shift_local)), _ip(_i32(lay.border_width)), p["radius"], p["o1"], p["o2"],
        p["pab"].shape[1], p["pab"].shape[0], _dp(_f64(p["pab"]).reshape(-1)), _dp(grid))


def _oracle_integrate(L, ortho, tau, p, lay, grid, hab, forces):
    virials = np.zeros(18)
    L.grid_oracle_integrate_pgf_product(
        ortho, tau, p["border_mask"], p["la_max"], p["la_min"], p["lb_max"], p["lb_min"], p["zeta"],
        p["zetb"], _dp(_f64(lay.dh).reshape(-1)), _dp(_f64(lay.dh_inv).reshape(-1)), _dp(_f64(p["ra"])),
        _dp(_f64(p["rab"])), _ip(_i32(lay.npts_global)), _ip(_i32(lay.npts_local)),
        _ip(_i32(lay.shift_local)), _ip(_i32(lay.border_width)), p["radius"], p["o1"], p["o2"],
        p["pab"].shape[1], p["pab"].shape[0], _dp(grid), _dp(hab.reshape(-1)),
        _dp(_f64(p["pab"]).reshape(-1)), _dp(forces.reshape(-1)), _dp(virials))


@pytest.mark.parametrize("ortho", [True, False], ids=["ortho", "triclinic"])
@pytest.mark.parametrize("func", [100, 200, 302, 413, 503], ids=["AB", "DADB", "ADBmDAB_Y", "ARDBmDARB_XZ", "DABpADB_Z"])
def test_batched_collocate_vs_oracle(b200, oracle, ortho, func):
    rng = np.random.default_rng(7 + func + ortho)
    npts = np.array([30, 32, 36])
    cell = np.diag([6.0, 6.4, 7.2])
    if not ortho:
        cell = cell + np.array([[0, 0.3, 0.1], [0.2, 0, -0.2], [0.1, 0.3, 0]])
    dh = cell / npts[:, None]
    lay = GridLayout(npts, npts, [0, 0, 0], [0, 0, 0], dh, np.linalg.inv(dh))
    prods = _random_products(rng, 23, lay)
    want = np.full(int(np.prod(npts)), 0.125)
    for p in prods:
        _oracle_collocate(oracle.lib, ortho, func, p, lay, want)
    got = np.full(int(np.prod(npts)), 0.125)
    keys = ("border_mask", "la_max", "la_min", "lb_max", "lb_min", "zeta", "zetb", "rscale", "radius", "o1",
            "o2")
    b200.collocate_pgf_products(orthorhombic=ortho, func=func, layout=lay, grid=got,
                                ra=np.stack([p["ra"] for p in prods]), rab=np.stack([p["rab"] for p in prods]),
                                pab=[p["pab"] for p in prods], **{k: [p[k] for p in prods] for k in keys})
    assert rel_diff(got, want) < 1e-10


@pytest.mark.parametrize("ortho", [True, False], ids=["ortho", "triclinic"])
@pytest.mark.parametrize("tau", [False, True], ids=["v", "tau"])
def test_batched_integrate_vs_oracle(b200, oracle, ortho, tau):
    rng = np.random.default_rng(11 + 2 * tau + ortho)
    npts = np.array([30, 32, 36])
    cell = np.diag([6.0, 6.4, 7.2])
    if not ortho:
        cell = cell + np.array([[0, 0.3, 0.1], [0.2, 0, -0.2], [0.1, 0.3, 0]])
    dh = cell / npts[:, None]
    lay = GridLayout(npts, npts, [0, 0, 0], [0, 0, 0], dh, np.linalg.inv(dh))
    prods = _random_products(rng, 19, lay)
    grid = rng.normal(size=int(np.prod(npts)))
    want_h = [np.full(p["pab"].shape, 0.5) for p in prods]
    want_f = np.full((len(prods), 2, 3), 0.25)
    for p, h, f in zip(prods, want_h, want_f):
        _oracle_integrate(oracle.lib, ortho, tau, p, lay, grid, h, f)
    got_h = [np.full(p["pab"].shape, 0.5) for p in prods]
    got_f = np.full((len(prods), 2, 3), 0.25)
    keys = ("border_mask", "la_max", "la_min", "lb_max", "lb_min", "zeta", "zetb", "radius", "o1", "o2")
    b200.integrate_pgf_products(orthorhombic=ortho, compute_tau=tau, layout=lay, grid=grid, hab=got_h,
                                ra=np.stack([p["ra"] for p in prods]), rab=np.stack([p["rab"] for p in prods]),
                                pab=[p["pab"] for p in prods], forces=got_f,
                                **{k: [p[k] for p in prods] for k in keys})
    for g, w in zip(got_h, want_h):
        assert rel_diff(g, w) < 1e-10
    assert rel_diff(got_f, want_f) < 1e-8
