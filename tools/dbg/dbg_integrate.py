import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from cp2k_b200 import load_b200, OffloadBuffer
from replay import load_task, dummy_task_list, golden_grid, rel_diff, ncoset
from synth import make_workload
lib = load_b200()
def run(variant, t, with_fv=True, cycles=1):
    lib.set_kernel_variant(variant)
    n1, n2 = t["n1"], t["n2"]
    tl, nblocks = dummy_task_list(lib, t, cycles, 1)
    pab, hab = OffloadBuffer(nblocks*n1*n2), OffloadBuffer(nblocks*n1*n2)
    pab.host.reshape(nblocks, n2, n1)[:] = 0.5 * t["pab"]
    grid = OffloadBuffer(int(np.prod(t["npts_local"]))); grid.host[:] = golden_grid(t)
    f, v = (np.zeros((2,3)), np.zeros((3,3))) if with_fv else (None, None)
    tl.integrate(t["func"] == 200, pab if with_fv else None, [grid], hab, f, v)
    st = lib.stats(tl)
    tl.free()
    return hab.host.copy(), f, v, st
for name in ["ortho_density_l0000", "ortho_density_l0122", "ortho_density_l3333"]:
    t = load_task(name)
    print(name, "npts", t["npts_local"], "radius", t["radius"], "la/lb", t["la_max"], t["lb_max"])
    for fv in (False, True):
        ref = run(1, t, fv)
        for var in (3, 2):
            out = run(var, t, fv)
            msg = f"  fv={fv} variant={var} visits={out[3]['npairs']} hab err {rel_diff(out[0], ref[0]):.2e}"
            if fv:
                msg += f" forces err {rel_diff(out[1], ref[1]):.2e} virial err {rel_diff(out[2], ref[2]):.2e}"
            print(msg)
# collocate with l growth on a synthetic list
wl = make_workload(seed=5, natoms=6, max_tasks=800)
pab = wl.random_pab(1)
for func in (100, 200, 301, 411, 501):
    res = {}
    for var in (1, 3, 2):
        lib.set_kernel_variant(var)
        tl = wl.create(lib); grids = wl.new_grids(); tl.collocate(func, pab, grids); tl.free()
        res[var] = [g.host.copy() for g in grids]
    print("collocate func", func, "ctile err", max(rel_diff(a, b) for a, b in zip(res[3], res[1])),
          "warptile err", max(rel_diff(a, b) for a, b in zip(res[2], res[1])))
lib.set_kernel_variant(0)
