import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import socialmediamodel_b200 as smm
os.environ["SMM_PDE_CLUSTER"] = "1"
p = smm.PDEconstruct(); q = smm.parameters()
uz, qq = smm.constructinitial("4inf", (p, q))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
smm.pde_solve_raw(p, [qq], uz.x[0][..., None], uz.x[1][..., None], uz.x[2][..., None], dt=0.001, nsteps=n, savesteps=[n], precision="f32")
import ctypes
from socialmediamodel_b200 import _lib
lib = _lib.load()
buf = (ctypes.c_ulonglong * 64)()
print("rc", lib.smm_debug_cluster_times(buf))
t = [buf[i] for i in range(16)]
names = {1: "(start) -> P2 entry", 4: "P2 fill+fwd", 5: "P2 mul+inv", 6: "P2 scatter", 7: "cluster barrier", 8: "P3 fill+halo+ODE", 9: "P3 inverse", 10: "P3 force", 11: "syncthreads",
         12: "P4 stencil", 13: "P4 moments+scatter", 14: "P1' rows fwd + barrier", 15: "end"}
order = [3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15]
for a_, b_ in zip(order[:-1], order[1:]):
    print(f"{names.get(b_, ''):26s} {t[b_] - t[a_]:6d} ns")
print("stage total", t[15] - t[3])
