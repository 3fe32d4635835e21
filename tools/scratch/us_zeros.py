import sys; sys.path.insert(0, "/root/repo")
import numpy as np
import socialmediamodel_b200 as smm
p, q = smm.ABMconstruct(), smm.parameters()
st = smm.savetimes_abm(200, 4)
o = smm.abm_ensemble_raw(200, 64, p=p, q=q, savesteps=st, first_realization=1, precision="f32", us_dtype=np.float32)
us = o["us"]
print("shape", us.shape, "zeros fraction", float((us == 0).mean()))
for s in range(us.shape[4]):
    print("save", s, "zeros", float((us[:, :, :, :, s, :] == 0).mean()), "below 1e-30", float((np.abs(us[:, :, :, :, s, :]) < 1e-30).mean()))
# zero runs along x (fastest axis): fraction of 16-float groups that are entirely zero
g = us.reshape(-1, order="F")
g = g[: g.size // 16 * 16].reshape(-1, 16)
print("16-groups all zero", float((g == 0).all(axis=1).mean()))
