import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from cp2k_b200 import load_b200, OffloadBuffer
from cp2k_b200.grid_api import GRID_BACKEND_CPU, GRID_BACKEND_GPU, GRID_BACKEND_REF
from cp2k_b200.workload import build_h2o_workload
from synth import make_workload
from replay import TASK_NAMES, load_task, replay_batched, rel_diff
from oracle import pyref
lib = load_b200()
ref = pyref.load_reference_gpu(0)
mk = OffloadBuffer.with_device
print("golden vectors through the reference GPU backend (max rel diff):")
for name in TASK_NAMES:
    try:
        ec = replay_batched(ref, load_task(name), True, make_buffer=mk)
        ei = replay_batched(ref, load_task(name), False, make_buffer=mk)
        print(f"  {name:28s} collocate {ec:.2e} integrate {ei:.2e}")
    except Exception as e:
        print("  ", name, "EXC", e)
def run(L, wl, make=mk):
    tl = wl.create(L); pab = wl.random_pab(1, make=make); grids = wl.new_grids(make=make)
    tl.collocate(100, pab, grids)
    hab = make(wl.pab_len); tl.integrate(False, None, grids, hab, None, None); tl.free()
    return [g.host.copy() for g in grids], hab.host.copy()
for label, wl in (("synth seed 11", make_workload(seed=11, natoms=6, max_tasks=1500)),
                  ("synth both_orders=False", make_workload(seed=11, natoms=6, max_tasks=1500, both_orders=False)),
                  ("H2O-64 36 atoms", build_h2o_workload("H2O-64", max_atoms=36))):
    ref.set_backend(GRID_BACKEND_GPU); g_gpu = run(ref, wl)
    ref.set_backend(GRID_BACKEND_CPU); g_cpu = run(ref, wl)
    g_b2 = run(lib, wl)
    print(label, "ntasks", wl.ntasks)
    for l in range(len(g_cpu[0])):
        print(f"   level {l}: refgpu-vs-refcpu {rel_diff(g_gpu[0][l], g_cpu[0][l]):.2e}  b200-vs-refcpu {rel_diff(g_b2[0][l], g_cpu[0][l]):.2e}  max|cpu| {np.abs(g_cpu[0][l]).max():.3e}")
    print(f"   hab: refgpu-vs-refcpu {rel_diff(g_gpu[1], g_cpu[1]):.2e}  b200-vs-refcpu {rel_diff(g_b2[1], g_cpu[1]):.2e}")
