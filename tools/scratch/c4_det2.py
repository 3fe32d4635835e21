import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import socialmediamodel_b200 as smm
os.environ.setdefault("SMM_FFT_FAST", "1"); os.environ.setdefault("SMM_PDE_STAGE2", "1")
p = smm.PDEconstruct(dx=0.0125)
qs = [smm.parameters(a=0.5, sigma=0.75), smm.parameters(a=2.0, b=6.0, eta=1.0), smm.parameters(c=4.0, sigma=0.25)]
inits = [smm.constructinitial("4inf", (p, q)) for q in qs]
ql = [qq for _, qq in inits]
rng = np.random.default_rng(3)
U = np.stack([uz.x[0] * (1 + 0.2 * rng.random(uz.x[0].shape)) for uz, _ in inits], axis=-1)
Z = np.stack([uz.x[1] for uz, _ in inits], axis=-1); Y = np.stack([uz.x[2] for uz, _ in inits], axis=-1)
def solve(idx, n=8):
    return smm.pde_solve_raw(p, [ql[i] for i in idx], U[..., idx], Z[..., idx], Y[..., idx], dt=5e-5, nsteps=n, savesteps=[n], precision="f32")["us"]
runs = [solve([0, 1, 2]) for _ in range(8)]
print(os.environ.get("TAG"), [np.array_equal(runs[0], r) for r in runs[1:]], [float(np.abs(runs[0] - r).max()) for r in runs[1:]],
      "per instance:", [[np.array_equal(runs[0][..., b], r[..., b]) for b in range(3)] for r in runs[1:]])
