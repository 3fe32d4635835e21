"""Generate the golden fixtures of tests/golden/ from the Float64 oracle (seeded NumPy inputs).

The reference ships no golden data and cannot be run here (Julia absent), so these vectors are
ORACLE outputs: they freeze the restatement of abm.jl / pde.jl / ensemble.jl at the commit that wrote
them, so that later edits of the oracle cannot drift silently, and they give the GPU tests fixed
inputs and expected outputs.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import abm_oracle as A      # noqa: E402
from oracle import ensemble_oracle as E  # noqa: E402
from oracle import pde_oracle as P      # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def abm_case(seed, n, NT, net_k=None):
    rng = np.random.default_rng(seed)
    p, q = A.default_p(), A.default_q(n=n)
    x0 = rng.uniform(-2.2, 2.2, (n, 2))
    state = np.where(rng.random(n) > 0.5, 1, -1)
    f0, i0, m0 = A.abm_init(x0)
    xi = rng.standard_normal((NT, n, 2)); u1 = rng.random((NT, n)); u2 = rng.random((NT, n))
    net = None
    if net_k:
        cols = np.stack([np.append(rng.choice(np.delete(np.arange(n), j), size=net_k, replace=False), j) for j in range(n)])
        net = (np.arange(0, n * (net_k + 1) + 1, net_k + 1, dtype=np.int32), cols.reshape(-1).astype(np.int32))
    o = A.abm_solve(NT, x0, state, f0, i0, m0, xi, u1, u2, p, q, net)
    d = dict(n=n, NT=NT, x0=x0, state=state, follow0=f0, inf0=i0, media0=m0, xi=xi, u1=u1, u2=u2,
             xs=o["xs"], follows=o["follows"], infs=o["infs"], meds=o["meds"], min_margin=o["min_margin"])
    if net:
        d.update(rowptr=net[0], colidx=net[1])
    st = E.savetimes_abm(NT, 3)[-1:]          # one snapshot keeps the fixture small
    d["savesteps"] = st
    d["us"] = np.stack([E.kde_snapshot(o["xs"][t], o["follows"][t], state, 4, p["X"], p["Y"], n) for t in st], axis=-1)
    return d


def pde_case():
    p = P.pde_construct_coarse()
    q = A.default_q()
    u0, z0, y0, ci, cm = P.uniforminit(p, q)
    rng = np.random.default_rng(7)
    u = u0 * (1 + 0.3 * rng.random(u0.shape)) + 0.002 * rng.random(u0.shape)
    z = z0 + 0.1 * rng.standard_normal(z0.shape); y = y0 + 0.1 * rng.standard_normal(y0.shape)
    du, dz, dy = P.rhs(u, z, y, p, q)
    du0, dz0, dy0 = P.rhs(u0, z0, y0, p, q)
    u5 = np.concatenate([u, np.zeros_like(u[..., :1])], axis=3)
    y5 = np.concatenate([y, np.array([[0.2], [-0.3]])], axis=1)
    q5 = dict(q, L=5)
    ci5, cm5 = np.array([0, 3.0, 0, 0, 1.0]), np.array([0.0, 1.0])
    du5, dz5, dy5 = P.rhs(u5, z, y5, p, q5, ci5, cm5)
    ur, zr, yr = P.rk_solve(u0, z0, y0, 10, 1e-3, p, q)
    return dict(u=u, z=z, y=y, du=du, dz=dz, dy=dy, du0=du0, dz0=dz0, dy0=dy0, u5=u5, y5=y5, ci5=ci5, cm5=cm5,
                du5=du5, dz5=dz5, dy5=dy5, u_rk4_10=ur, z_rk4_10=zr, y_rk4_10=yr)


def adaptive_control_case():
    """Adaptive integrator (controller log + states) and one fixed-target control evaluation on the coarse grid."""
    from oracle import control_oracle as CO
    p = P.pde_construct_coarse()
    q = A.default_q()
    u0, z0, y0, ci, cm = P.uniforminit(p, q)
    save = np.array([0.0, 0.02, 0.05])
    states, log = P.bs3_adaptive(u0, z0, y0, save, p, q)
    qo = A.default_q(eta=1.0)
    uzy0, sl, q2, ci2, cm2, slm = CO.prep(0.02, p, qo, start="zero", savedt=0.01, dt=1e-3, countercontrol="inf")
    ts = np.array([0.0, 0.025, 0.05]); targets = np.array([[0.3, 0.3], [-0.5, 0.8]]).T
    fs, pen, (uf, zf, yf) = CO.fixedtargets(ts, targets, sl, p, q2, ci2, cm2, uzy0, savedt=0.01, dt=1e-3, countercontrol="inf",
                                            stubborntarget=np.array([1.5, 1.5]))
    return dict(save=save, log=np.array([(t, dt, n, float(a)) for t, dt, n, a in log]),
                u_adapt=np.stack([s_[0] for s_ in states], axis=-1), y_adapt=np.stack([s_[2] for s_ in states], axis=-1),
                ctrl_ts=ts, ctrl_targets=targets, ctrl_followersum=fs, ctrl_penalty=pen, ctrl_y_final=yf,
                ctrl_u_final_sum=uf.sum(axis=(0, 1)))


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "pde_adaptive_control.npz"), **adaptive_control_case())
    np.savez_compressed(os.path.join(HERE, "abm_dense_n60.npz"), **abm_case(101, 60, 25))
    np.savez_compressed(os.path.join(HERE, "abm_csr_n80.npz"), **abm_case(202, 80, 15, net_k=6))
    np.savez_compressed(os.path.join(HERE, "pde_coarse.npz"), **pde_case())
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
