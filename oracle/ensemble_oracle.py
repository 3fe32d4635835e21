"""Float64 NumPy restatement of the reference ensemble drivers (src/ensemble.jl) and of the
ensemble statistics in src/plotting.jl:361-363.

TEST INFRASTRUCTURE -- see oracle/__init__.py.  PARITY UNPINNED (no reference fixtures).
"""
import numpy as np

from . import abm_oracle as A


def savetimes_abm(NT, savepoints):
    """ensemble.jl:21: Int.(round.(LinRange(1,NT,savepoints))) -- 1-based indices into xs, so the
    0-based step indices are these minus one (the last save is step NT-1, never NT)."""
    return np.round(np.linspace(1, NT, savepoints)).astype(int) - 1


def kde_snapshot(x, follow, state, L, X, Y, n, sigma=0.1):
    """ensemble.jl:27-38 + :2-8 + pde.jl:122: us[:,:,i,j] = (1/n) sum_{agents with medium i and
    influencer j} gaussian(grid point, x_agent)."""
    N_x, N_y = X.shape
    us = np.zeros((N_x, N_y, 2, L))
    xa, ya = X[:, 0], Y[0, :]
    for i, s in enumerate((-1, 1)):
        for j in range(L):
            sel = (follow == j) & (state == s)
            if sel.any():
                ex = np.exp(-1 / (2 * sigma ** 2) * (xa[:, None] - x[sel, 0][None, :]) ** 2)
                ey = np.exp(-1 / (2 * sigma ** 2) * (ya[:, None] - x[sel, 1][None, :]) ** 2)
                us[:, :, i, j] = (1 / n) * (1 / (2 * np.pi * sigma ** 2)) * (ex @ ey.T)
    return us


def kde_snapshot_naive(x, follow, state, L, X, Y, n, sigma=0.1):
    """The same, written like the reference's comprehension (slow; for cross-checking)."""
    from .pde_oracle import gaussian
    N_x, N_y = X.shape
    us = np.zeros((N_x, N_y, 2, L))
    for i, s in enumerate((-1, 1)):
        for j in range(L):
            sel = (follow == j) & (state == s)
            for a in np.nonzero(sel)[0]:
                us[:, :, i, j] += gaussian(np.stack([X, Y], axis=-1), x[a], sigma)
    return us / n


def abm_ensemble(NT, runs, p, q, savepoints=4, sigma=0.1, net=None):
    """ABMsolveensemble (ensemble.jl:11-55).  ``runs`` is a list of dicts with the injected
    x0,state,follow0,inf0,media0,xi,u1,u2 of each realization (the reference seeds realization
    k with chosenseed=k, ensemble.jl:25).  Returns us[N_x,N_y,2,L,sp,N], zs[2,2,sp,N],
    ys[2,L,sp,N] and the ensemble means (plotting.jl:361-363)."""
    L, n = q["L"], q["n"]
    X, Y = p["X"], p["Y"]
    N = len(runs)
    st = savetimes_abm(NT, savepoints)
    us = np.zeros(X.shape + (2, L, savepoints, N))
    zs = np.zeros((2, 2, savepoints, N))
    ys = np.zeros((2, L, savepoints, N))
    for k, r in enumerate(runs):
        out = A.abm_solve(NT, r["x0"], r["state"], r["follow0"], r["inf0"], r["media0"],
                          r["xi"], r["u1"], r["u2"], p, q, net)
        for m, t in enumerate(st):
            us[..., m, k] = kde_snapshot(out["xs"][t], out["follows"][t], r["state"], L, X, Y, n, sigma)
            ys[:, :, m, k] = out["infs"][t].T          # ensemble.jl:39 (transpose, Q9)
            zs[:, :, m, k] = out["meds"][t].T          # ensemble.jl:41
    means = dict(av_u=us.sum(axis=5) / N, av_z=zs.sum(axis=3) / N, av_y=ys.sum(axis=3) / N)
    return us, zs, ys, means
