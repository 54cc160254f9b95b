import time, sys
sys.path.insert(0, '/root/repo')
import torch
from cp2k_b200 import load_b200
from cp2k_b200.workload import build_h2o_workload
t=time.time(); wl=build_h2o_workload("H2O-256"); print("workload gen %.2fs"%(time.time()-t), wl.ntasks)
lib=load_b200()
for i in range(3):
    torch.cuda.synchronize(); t=time.time(); tl=wl.create(lib); torch.cuda.synchronize(); print("create %.3fs"%(time.time()-t))
    st=lib.stats(tl); 
    if i<2: tl.free()
print(st)
print(torch.cuda.memory_allocated()/1e9, torch.cuda.mem_get_info())
