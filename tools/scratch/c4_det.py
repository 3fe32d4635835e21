import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import socialmediamodel_b200 as smm
os.environ["SMM_FFT_FAST"] = "1"; os.environ["SMM_PDE_STAGE2"] = "1"
p = smm.PDEconstruct(dx=0.0125)
qs = [smm.parameters(a=0.5, sigma=0.75), smm.parameters(a=2.0, b=6.0, eta=1.0), smm.parameters(c=4.0, sigma=0.25)]
inits = [smm.constructinitial("4inf", (p, q)) for q in qs]
ql = [qq for _, qq in inits]
rng = np.random.default_rng(3)
U = np.stack([uz.x[0] * (1 + 0.2 * rng.random(uz.x[0].shape)) for uz, _ in inits], axis=-1)
Z = np.stack([uz.x[1] for uz, _ in inits], axis=-1); Y = np.stack([uz.x[2] for uz, _ in inits], axis=-1)
dt = 5e-5
def solve(idx, n=8):
    return smm.pde_solve_raw(p, [ql[i] for i in idx], U[..., idx], Z[..., idx], Y[..., idx], dt=dt, nsteps=n, savesteps=[n], precision="f32")["us"]
for n in (1, 8):
    b3 = solve([0, 1, 2], n); b3b = solve([0, 1, 2], n)
    print(n, "batch repeat equal:", np.array_equal(b3, b3b))
    for b in range(3):
        s1 = solve([b], n); s2 = solve([b], n)
        print(n, b, "solo repeat equal:", np.array_equal(s1, s2), " solo == batch:", np.array_equal(s1[..., 0, 0], b3[..., 0, b]), np.abs(s1[..., 0, 0] - b3[..., 0, b]).max())
    b2 = solve([0, 1], n)
    print(n, "pair == batch:", np.array_equal(b2[..., 0, 0], b3[..., 0, 0]), np.array_equal(b2[..., 0, 1], b3[..., 0, 1]))
