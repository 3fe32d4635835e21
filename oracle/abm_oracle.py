"""Float64 NumPy restatement of the reference ABM time-stepper (src/abm.jl).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  PARITY UNPINNED (no reference
fixtures, no Julia here); pinned to the closed-form KATs of SURVEY.md section 4.

Conventions (0-based here, 1-based in the reference):
  x[n,2], inf[L,2], media[M,2]   rows per entity, as in abm.jl:18,33
  follow[n] in 0..L-1            the column k with FolInfNet[j,k]==1 (abm.jl:27-30)
  state[n] in {-1,+1}            abm.jl:39; media row 0 <-> state -1, row 1 <-> +1 (abm.jl:74-78)
  network: None = all-to-all incl. self (abm.jl:42), or CSR (rowptr[n+1], colidx[nnz]);
           row j lists the i with IndNet[j,i]==1 (abm.jl:51), self-loop only if listed.
Random draws are *injected*: xi[n,2] ~ N(0,1) (abm.jl:173), u1[n], u2[n] ~ U[0,1)
(abm.jl:122,126), xi_inf[L,2], xi_med[M,2] (abm.jl:181-196; multiplied by
sigmahat/sigmatilde which default to 0, setting.jl:10-11).
"""
import numpy as np

THETA = 0.1  # abm.jl:94 (hard-coded threshold of the recommender function)


def default_q(**kw):
    """setting.jl:1-22 (model parameters)."""
    q = dict(L=4, M=2, n=250, eta=15.0, a=1.0, b=4.0, c=2.0, sigma=0.5, sigmahat=0.0,
             sigmatilde=0.0, frictionI=10.0, frictionM=100.0,
             controlspeed1=np.sqrt(1.5 ** 2 + 1.5 ** 2) / 5, controltarget1=np.array([1.5, 1.5]),
             controlspeed2=np.sqrt(1.5 ** 2 + 1.5 ** 2) / 5, controltarget2=np.array([1.5, 1.5]))
    q.update(kw)
    return q


def default_p(dt=0.01, dx=0.05, dy=None, domain=None):
    """setting.jl:30-43 (ABMconstruct)."""
    dy = dx if dy is None else dy
    domain = 2.2 * np.array([[-1.0, 1.0], [-1.0, 1.0]]) if domain is None else np.asarray(domain, float)
    xa = julia_range(domain[0, 0], dx, domain[0, 1])
    ya = julia_range(domain[1, 0], dy, domain[1, 1])
    X, Y = np.meshgrid(xa, ya, indexing="ij")
    return dict(dt=dt, dx=dx, dy=dy, domain=domain, X=X, Y=Y, dV=dx * dy)


def julia_range(a, step, b):
    """Values of the Julia range a:step:b (length floor((b-a)/step + tiny)+1, a + i*step)."""
    nsteps = int(np.floor((b - a) / step + 1e-9))
    return a + step * np.arange(nsteps + 1)


def abm_init(x, L=4):
    """abm.jl:8-45 given the uniform positions x (already scaled to the domain) and
    u_state[n] handled by the caller: quadrant followers (abm.jl:21-24), influencer = mean of
    its followers (abm.jl:33-36, NaN if it has none), media = [-1 -1; 1 1] (abm.jl:18)."""
    assert L == 4  # abm.jl:13
    follow = np.where(x[:, 1] > 0, np.where(x[:, 0] > 0, 0, 1), np.where(x[:, 0] > 0, 2, 3))
    inf = np.full((L, 2), np.nan)
    for i in range(L):
        sel = follow == i
        with np.errstate(invalid="ignore", divide="ignore"):
            inf[i] = x[sel].sum(axis=0) / sel.sum()
    media = np.array([[-1.0, -1.0], [1.0, 1.0]])
    return follow, inf, media


def attraction(x, net=None):
    """abm.jl:47-67.  force_j = sum_i w_ij (x_i-x_j) / sum_i w_ij, w_ij = exp(-|x_i-x_j|),
    i over the neighbours of j (self included when listed: +1 in the denominator)."""
    n = x.shape[0]
    if net is None:
        d = x[None, :, :] - x[:, None, :]            # d[j,i] = x_i - x_j   (abm.jl:58)
        w = np.exp(-np.sqrt(d[..., 0] ** 2 + d[..., 1] ** 2))
        return (w[..., None] * d).sum(axis=1) / w.sum(axis=1)[:, None]
    rowptr, colidx = net
    force = np.zeros((n, 2))
    deg = np.diff(rowptr)
    rows = np.repeat(np.arange(n), deg)
    d = x[colidx] - x[rows]
    w = np.exp(-np.sqrt(d[:, 0] ** 2 + d[:, 1] ** 2))
    wsum = np.bincount(rows, weights=w, minlength=n)
    fx = np.bincount(rows, weights=w * d[:, 0], minlength=n)
    fy = np.bincount(rows, weights=w * d[:, 1], minlength=n)
    nz = deg > 0                                      # abm.jl:52-53: no neighbours -> 0
    force[nz, 0] = fx[nz] / wsum[nz]
    force[nz, 1] = fy[nz] / wsum[nz]
    return force


def influence(x, media, inf, follow, state, b, c):
    """abm.jl:69-88."""
    force1 = media[(state == 1).astype(int)] - x      # row 1 (0-based) if state==1 else row 0
    force2 = inf[follow] - x
    return c * force1 + b * force2


def boundaryconditions(x, domain):
    """abm.jl:139-150: single reflection per coordinate, in place, returns its argument."""
    for d in range(2):
        hi = x[:, d] > domain[d, 1]
        lo = x[:, d] < domain[d, 0]
        x[hi, d] = -x[hi, d] + 2 * domain[d, 1]
        x[lo, d] = -x[lo, d] + 2 * domain[d, 0]
    return x


def follower_stats(x, follow, state, L):
    """cnt_i, S_i = sum of follower positions, statesum_i (abm.jl:96-100,179-180)."""
    cnt = np.bincount(follow, minlength=L).astype(float)
    S = np.stack([np.bincount(follow, weights=x[:, 0], minlength=L),
                  np.bincount(follow, weights=x[:, 1], minlength=L)], axis=1)
    ssum = np.bincount(follow, weights=state.astype(float), minlength=L)
    return cnt, S, ssum


def changeinfluencer(state, x, follow, inf, u1, u2, eta, dt, L):
    """abm.jl:90-137 with injected uniforms.  Returns (follow_new, lam, margin) where margin is
    the smallest distance of a decision variable to its threshold (diagnostic for FP32 parity)."""
    n = x.shape[0]
    cnt = np.bincount(follow, minlength=L).astype(float)
    ssum = np.bincount(follow, weights=state.astype(float), minlength=L)
    with np.errstate(invalid="ignore", divide="ignore"):
        fraction = ssum / cnt                          # abm.jl:99 (0/0 = NaN, Q5)
    d = x[:, None, :] - inf[None, :, :]
    dist = np.exp(-np.sqrt(d[..., 0] ** 2 + d[..., 1] ** 2))   # abm.jl:103-109
    g2 = state[:, None] * fraction[None, :]
    g2 = np.where(g2 < THETA, THETA, g2)               # abm.jl:115-118 (NaN < theta is false)
    attractive = eta * dist * g2                       # abm.jl:119
    lam = np.zeros(n)
    for i in range(L):                                 # sequential left-to-right sum (abm.jl:123)
        lam = lam + attractive[:, i]
    with np.errstate(invalid="ignore"):
        thr = 1.0 - np.exp(-lam * dt)
        switch = u1 < thr                              # abm.jl:124 (NaN -> False)
    follow_new = follow.copy()
    margin = np.abs(u1 - thr)
    if switch.any():
        p = attractive[switch] / lam[switch][:, None]  # abm.jl:125
        cs = np.zeros(p.shape[0])
        k = np.full(p.shape[0], L - 1)                 # Q7: clamp to L
        done = np.zeros(p.shape[0], bool)
        m2 = np.full(p.shape[0], np.inf)
        for i in range(L):                             # abm.jl:127-130: while sum(p[1:k]) < r2
            cs = cs + p[:, i]
            hit = (~done) & ~(cs < u2[switch])
            k[hit] = i
            done |= hit
            if i < L - 1:
                m2 = np.minimum(m2, np.abs(cs - u2[switch]))
        follow_new[switch] = k                         # abm.jl:131-132 (may re-pick the current one, Q6)
        margin[switch] = np.minimum(margin[switch], m2)
    return follow_new, lam, margin


def abm_step(x, inf, media, follow, state, xi, u1, u2, p, q, net=None, xi_inf=None, xi_med=None):
    """One pass of the loop body abm.jl:166-213.  Inputs are not modified.
    Returns x', inf', media', follow', margin."""
    dt, domain = p["dt"], p["domain"]
    L, M = q["L"], q["M"]
    xold = x
    force = q["a"] * attraction(xold, net) + influence(xold, media, inf, follow, state, q["b"], q["c"])
    xnew = xold + dt * force + np.sqrt(dt) * q["sigma"] * xi          # abm.jl:173

    inf = inf.copy()
    media = media.copy()
    cnt, S, _ = follower_stats(xold, follow, state, L)
    zi = np.zeros((L, 2)) if xi_inf is None else xi_inf
    zm = np.zeros((M, 2)) if xi_med is None else xi_med
    for i in range(L):                                                # abm.jl:177-185
        noise = 1 / q["frictionI"] * np.sqrt(dt) * q["sigmahat"] * zi[i]
        if cnt[i] > 0:
            mc = S[i] / cnt[i]
            inf[i] = inf[i] + dt / q["frictionI"] * (mc - inf[i]) + noise
        else:
            inf[i] = inf[i] + noise
    states = [-1, 1]
    for i in range(M):                                                # abm.jl:188-198
        sel = state == states[i]
        noise = 1 / q["frictionM"] * np.sqrt(dt) * q["sigmatilde"] * zm[i]
        if sel.sum() > 0:
            mc = xold[sel].sum(axis=0) / sel.sum()
            media[i] = media[i] + dt / q["frictionM"] * (mc - media[i]) + noise
        else:
            media[i] = media[i] + noise
    boundaryconditions(xnew, domain)                                  # abm.jl:201-203
    boundaryconditions(inf, domain)
    boundaryconditions(media, domain)
    # Q1: changeinfluencer sees the *updated, reflected* influencers (infold aliases inf, abm.jl:175,207),
    # the OLD agent positions and the OLD follower network.
    follow_new, _, margin = changeinfluencer(state, xold, follow, inf, u1, u2, q["eta"], dt, L)
    return xnew, inf, media, follow_new, margin


def abm_solve(NT, x0, state, follow0, inf0, media0, xi, u1, u2, p, q, net=None,
              xi_inf=None, xi_med=None):
    """ABMsolve (abm.jl:154-217) with injected initial condition and draws.
    xi[NT,n,2], u1[NT,n], u2[NT,n].  Returns dict with xs[NT+1,n,2], follows[NT+1,n] (0-based;
    the reference's stateinfs = follows+1 as Float64, abm.jl:164,212), infs[NT+1,L,2],
    meds[NT+1,M,2], min_margin, margins[NT,n] (distance of every agent's decision variables to their thresholds at
    each step: which FP32 decisions are near-ties of the Float64 arithmetic)."""
    x, inf, media, follow = x0.copy(), inf0.copy(), media0.copy(), follow0.copy()
    xs, fs, infs, meds = [x.copy()], [follow.copy()], [inf.copy()], [media.copy()]
    mm = np.inf
    margins = []
    for k in range(NT):
        x, inf, media, follow, margin = abm_step(
            x, inf, media, follow, state, xi[k], u1[k], u2[k], p, q, net,
            None if xi_inf is None else xi_inf[k], None if xi_med is None else xi_med[k])
        mm = min(mm, np.nanmin(margin)) if np.isfinite(margin).any() else mm
        margins.append(margin)
        xs.append(x.copy()); fs.append(follow.copy()); infs.append(inf.copy()); meds.append(media.copy())
    return dict(xs=np.array(xs), follows=np.array(fs), infs=np.array(infs), meds=np.array(meds),
                min_margin=mm, margins=np.array(margins).reshape(NT, -1))


def follower_numbers(follows, state, L):
    """plotting.jl:284-307: numbers[t,i,j] = #agents with medium i and influencer j."""
    T = follows.shape[0]
    out = np.zeros((T, 2, L), dtype=np.int64)
    for i, s in enumerate((-1, 1)):
        sel = state == s
        for t in range(T):
            out[t, i] = np.bincount(follows[t][sel], minlength=L)
    return out
