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
names = ["P1 fill", "P1 fft", "P1 scatter", "sync1", "P2 fill+fwd", "P2 mul+inv", "P2 scatter", "sync2", "P3 fill+halo", "P3 inv", "P3 F", "sync3", "P4 stage", "P4 moments+scatter", "sync4", "P5"]
for i in range(1, 16):
    print(f"{names[i-1]:22s} {t[i] - t[i-1]:6d} ns")
print("stage total", t[15] - t[0])
