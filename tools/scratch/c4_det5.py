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
scheme = os.environ.get("SCHEME", "rk4")
def solve(prec, n):
    return smm.pde_solve_raw(p, ql, U, Z, Y, dt=5e-5, nsteps=n, savesteps=[n], scheme=scheme, precision=prec)
for n in (2, 3, 8):
    ref = solve("f64", n)
    for k in range(4):
        r = solve("f32", n)
        d = np.abs(r["us"] - ref["us"])[..., 0, :]      # [Nx,Ny,2,L,B]
        bad = d > 2e-6 * np.abs(ref["us"]).max()
        msg = ""
        if bad.any():
            idx = np.argwhere(bad)
            msg = f" BAD {len(idx)} x {idx[:,0].min()}-{idx[:,0].max()} y {idx[:,1].min()}-{idx[:,1].max()} inst {sorted(set(idx[:,4].tolist()))} planes {len(set(map(tuple, idx[:,2:4].tolist())))}"
        print(scheme, n, k, f"err {d.max() / np.abs(ref['us']).max():.2e} ys err {np.abs(r['ys'] - ref['ys']).max():.2e} zs err {np.abs(r['zs'] - ref['zs']).max():.2e}" + msg)
