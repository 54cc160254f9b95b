import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from cp2k_b200 import load_b200, OffloadBuffer
from synth import make_workload
lib = load_b200()
wl = make_workload(seed=11, natoms=6, max_tasks=1500)
def run(L, make):
    tl = wl.create(L); pab = wl.random_pab(1, make=make); grids = wl.new_grids(make=make)
    tl.collocate(100, pab, grids); tl.free()
    return [float(np.abs(g.host).max()) for g in grids]
print("b200 host-only   ", run(lib, OffloadBuffer))
print("b200 with device ", run(lib, OffloadBuffer.with_device))
from oracle import pyref
ref = pyref.load_reference_gpu(0)
print("after loading refgpu:")
print("b200 host-only   ", run(lib, OffloadBuffer))
print("b200 with device ", run(lib, OffloadBuffer.with_device))
print("refgpu           ", run(ref, OffloadBuffer.with_device))
print("b200 with device ", run(lib, OffloadBuffer.with_device))
