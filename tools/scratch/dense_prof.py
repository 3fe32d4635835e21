import os, sys, ctypes
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import socialmediamodel_b200 as smm
from socialmediamodel_b200 import _lib
p, q = smm.ABMconstruct(), smm.parameters()
st = smm.savetimes_abm(200, 4)
for R in (1, 128, 1024):
    a = smm.abm_ensemble_raw(200, R, p=p, q=q, savesteps=st, first_realization=1, precision="f32", want=("ys", "sums"))
    buf = (ctypes.c_longlong * 16)()
    _lib.load().smm_debug_dense_times(buf)
    t = [buf[i] for i in range(9)]
    names = ["fill xy2", "bin sums + barrier", "P2 (thread 0's share)", "pair tiles (thread 0)", "barrier", "epilogue", "barrier", "snapshot"]
    print("R", R, {n: t[i + 1] - t[i] for i, n in enumerate(names)}, "step total", t[8] - t[0])
