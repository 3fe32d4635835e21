"""Float64 NumPy restatement of the reference mean-field PDE (src/pde.jl, src/setting.jl:45-75).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  PARITY UNPINNED (no reference
fixtures, no Julia here); pinned to the closed-form KATs of SURVEY.md section 4.

Array conventions: u[ix,iy,i,j] (x index first, medium i, influencer j) exactly as the
reference's u[N_x,N_y,M,L] (pde.jl:96,189); z[2,M], y[2,L] columns per entity (pde.jl:109,115).
Time integration in the reference is delegated to DifferentialEquations.jl 7.6.0
(pde.jl:290-291; un-vendored, Manifest.toml:300-304); the build integrates with a fixed-step
explicit RK scheme, and this file carries the *same tableaus* (``rk_solve``) so trajectory
parity is scheme-for-scheme, plus an adaptive SciPy solve (``adaptive_solve``) as a
convergence check.
"""
import numpy as np

from .abm_oracle import default_q, julia_range  # noqa: F401  (re-exported)


AT_TARGET = 1e-12  # distance below which a controlled influencer/medium counts as arrived


def julia_round(v):
    """Julia ``round`` = ties-to-even (setting.jl:51-52, pde.jl:97-99); np.round agrees."""
    return int(np.round(v))


def second_derivative(N_x, dx):
    """pde.jl:68-77."""
    Mm = np.diag(np.full(N_x, -2.0)) + np.diag(np.ones(N_x - 1), 1) + np.diag(np.ones(N_x - 1), -1)
    Mm[0, 0] = -1.0
    Mm[-1, -1] = -1.0
    return Mm * (1 / dx ** 2)


def centered_difference(N_x, dx):
    """pde.jl:79-87."""
    C = 1 / (2 * dx) * (np.diag(np.ones(N_x - 1), 1) - np.diag(np.ones(N_x - 1), -1))
    C[0, 0:2] = 1 / dx * np.array([-1.0, 1.0])
    C[-1, -2:] = 1 / dx * np.array([-1.0, 1.0])
    return C


def generate_K(grid_points):
    """pde.jl:8-23 (dense; only for small grids)."""
    dxm = grid_points[None, :, 0] - grid_points[:, None, 0]   # [i,j] = g_j - g_i
    dym = grid_points[None, :, 1] - grid_points[:, None, 1]
    w = np.exp(-np.sqrt(dxm ** 2 + dym ** 2))
    return np.stack([w * dxm, w * dym], axis=-1), w


def pde_construct(dx=0.05, dy=None, domain=None, materialise=False):
    """setting.jl:45-67.  K_matrix/W_matrix only when ``materialise`` (1.5 GB at 89x89)."""
    dy = dx if dy is None else dy
    domain = 2.2 * np.array([[-1.0, 1.0], [-1.0, 1.0]]) if domain is None else np.asarray(domain, float)
    N_x = julia_round((domain[0, 1] - domain[0, 0]) / dx + 1)
    N_y = julia_round((domain[1, 1] - domain[1, 0]) / dy + 1)
    xa = domain[0, 0] + dx * np.arange(N_x)
    ya = domain[1, 0] + dy * np.arange(N_y)
    X, Y = np.meshgrid(xa, ya, indexing="ij")
    # vec(X): x index fastest (column-major) -- setting.jl:57
    grid_points = np.stack([X.ravel(order="F"), Y.ravel(order="F")], axis=1)
    p = dict(dx=dx, dy=dy, dV=dx * dy, X=X, Y=Y, N=N_x * N_y, N_x=N_x, N_y=N_y, domain=domain,
             grid_points=grid_points, xa=xa, ya=ya,
             C=centered_difference(N_x, dx), MassM=second_derivative(N_x, dx))
    if materialise:
        p["K_matrix"], p["W_matrix"] = generate_K(grid_points)
    return p


def pde_construct_coarse(**kw):
    """setting.jl:69-71."""
    return pde_construct(dx=0.1, dy=0.1, domain=2.3 * np.array([[-1.0, 1.0], [-1.0, 1.0]]), **kw)


def pde_construct_meso(**kw):
    """setting.jl:73-75."""
    return pde_construct(dx=44 / 600, dy=44 / 600, **kw)


def uniforminit(p, q):
    """pde.jl:90-119."""
    N_x, N_y, dV, domain, dx = p["N_x"], p["N_y"], p["dV"], p["domain"], p["dx"]
    assert q["M"] == 2 and q["L"] == 4
    rho = np.zeros((N_x, N_y, 2, 4))
    mid_y = julia_round(N_y / 2)
    mid_x = julia_round(N_x / 2)
    start_x = julia_round((domain[0, 1] - 2) / dx + 1)
    end_x = N_x - start_x + 1
    s, mx, my, e = start_x - 1, mid_x, mid_y, end_x       # 1-based inclusive -> python slices
    rho[s:mx, s:my, :, 3] = 1
    rho[s:mx, my:e, :, 1] = 1
    rho[mx:e, s:my, :, 2] = 1
    rho[mx:e, my:e, :, 0] = 1
    u0 = rho / (rho.sum() * dV)
    z0 = np.array([[-1.0, 1.0], [-1.0, 1.0]])
    y0 = np.array([[1.0, -1.0, 1.0, -1.0], [1.0, 1.0, -1.0, -1.0]])
    return u0, z0, y0, np.zeros(4), np.zeros(2)


def gaussian(x, center, sigma=0.1):
    """pde.jl:122."""
    d = np.asarray(x) - np.asarray(center)
    return 1 / (2 * np.pi * sigma ** 2) * np.exp(-1 / (2 * sigma ** 2) * (d ** 2).sum(axis=-1))


def randominit_from(p, q, random_pos, medium):
    """pde.jl:124-168 with the random draws injected: random_pos[n,2] in the domain,
    medium[n] in {0,1} (the reference's rand([1,2])-1)."""
    N_x, N_y, dV, gp = p["N_x"], p["N_y"], p["dV"], p["grid_points"]
    L = q["L"]
    rho = np.zeros((N_x, N_y, 2, L))
    counts = np.zeros((L, 2))
    y0 = np.zeros((2, L))
    for a in range(random_pos.shape[0]):
        px, py = random_pos[a]
        infi = (0 if px > 0 else 1) if py > 0 else (2 if px > 0 else 3)
        rho[:, :, medium[a], infi] += gaussian(gp, random_pos[a]).reshape((N_x, N_y), order="F")
        counts[infi, medium[a]] += 1
        y0[:, infi] += random_pos[a]
    u0 = rho / (rho.sum() * dV)
    z0 = np.array([[-1.0, 1.0], [-1.0, 1.0]])
    y0 = y0 / counts.sum(axis=1)[None, :]
    return u0, z0, y0, np.zeros(L), np.zeros(2)


def r_fun(x):
    """pde.jl:26-32: x>0.1 ? x : 0.1, so r(NaN)=0.1 (Q14)."""
    with np.errstate(invalid="ignore"):
        return np.where(x > 0.1, x, 0.1)


def agent_force_dense(rho, K_matrix, W_matrix, dV):
    """pde.jl:49-59 verbatim (gemv over the materialised matrices)."""
    v = rho.ravel(order="F")
    norms = dV * (W_matrix @ v)
    F = np.stack([(dV * (K_matrix[:, :, d] @ v) / norms).reshape(rho.shape, order="F") for d in range(2)], axis=-1)
    return F


def toeplitz_tables(N_x, N_y, dx, dy):
    """Appendix B: w[|dix|,|diy|] generates W, K (pde.jl:14-20 is translation invariant)."""
    ax = dx * np.arange(-(N_x - 1), N_x)
    ay = dy * np.arange(-(N_y - 1), N_y)
    DX, DY = np.meshgrid(ax, ay, indexing="ij")
    w = np.exp(-np.sqrt(DX ** 2 + DY ** 2))
    return w, w * DX, w * DY


def agent_force_toeplitz(rho, dx, dy):
    """Same operator as agent_force_dense, evaluated as three zero-padded correlations
    (Float64 FFT; 1e-15 relative to the largest term)."""
    from scipy.signal import fftconvolve
    N_x, N_y = rho.shape
    w, kx, ky = toeplitz_tables(N_x, N_y, dx, dy)
    # out[g] = sum_g' T[g'-g] rho[g']  = correlation = convolution with the flipped table
    def corr(T):
        full = fftconvolve(rho, T[::-1, ::-1], mode="full")
        return full[N_x - 1:2 * N_x - 1, N_y - 1:2 * N_y - 1]
    norms = corr(w)
    return np.stack([corr(kx) / norms, corr(ky) / norms], axis=-1)


def agent_force_direct(rho, dx, dy):
    """Direct O(N^2) evaluation of the Toeplitz sums row by row (exact summation, no FFT)."""
    N_x, N_y = rho.shape
    w, kx, ky = toeplitz_tables(N_x, N_y, dx, dy)
    F = np.zeros((N_x, N_y, 2))
    for ix in range(N_x):
        for iy in range(N_y):
            sl = (slice(N_x - 1 - ix, 2 * N_x - 1 - ix), slice(N_y - 1 - iy, 2 * N_y - 1 - iy))
            nrm = (w[sl] * rho).sum()
            F[ix, iy, 0] = (kx[sl] * rho).sum() / nrm
            F[ix, iy, 1] = (ky[sl] * rho).sum() / nrm
    return F


def gamma_rates(p, u, y, eta):
    """pde.jl:35-47.  Returns G[ix,iy,i,j2] (the rate into j2; independent of the source j)."""
    N_x, N_y, M, L = u.shape
    dV = p["dV"]
    m = dV * u.sum(axis=(0, 1))                                   # m[i,j]
    G = np.zeros((N_x, N_y, M, L))
    for i in range(M):
        io = (i + 1) % M                                           # mod1(i+1,M)
        for j2 in range(L):
            dist = np.sqrt((p["X"] - y[0, j2]) ** 2 + (p["Y"] - y[1, j2]) ** 2)
            with np.errstate(invalid="ignore", divide="ignore"):
                arg = (m[i, j2] - m[io, j2]) / (m[i, j2] + m[io, j2])
            G[:, :, i, j2] = eta * np.exp(-dist) * r_fun(arg)
    return G


def rhs(u, z, y, p, q, controlled_inf=None, controlled_med=None, force="toeplitz", parts=False):
    """The PDE/ODE right-hand side f (pde.jl:183-268).  Returns du, dz, dy."""
    a, b, c, sigma, eta = q["a"], q["b"], q["c"], q["sigma"], q["eta"]
    N_x, N_y, M, L = u.shape
    dx, dy_, dV, C, Mm = p["dx"], p["dy"], p["dV"], p["C"], p["MassM"]
    X, Y = p["X"], p["Y"]
    controlled_inf = np.zeros(L) if controlled_inf is None else controlled_inf
    controlled_med = np.zeros(M) if controlled_med is None else controlled_med
    D = sigma ** 2 * 0.5                                           # pde.jl:188
    rhosum = u.sum(axis=(2, 3))                                    # pde.jl:193
    if force == "dense":
        F = agent_force_dense(rhosum, p["K_matrix"], p["W_matrix"], dV)
    elif force == "direct":
        F = agent_force_direct(rhosum, dx, dy_)
    else:
        F = agent_force_toeplitz(rhosum, dx, dy_)
    G = gamma_rates(p, u, y, eta)
    du = np.zeros_like(u)
    dz = np.zeros_like(z)
    dyv = np.zeros_like(y)
    m_i = dV * u.sum(axis=(0, 1, 3))
    m_j = dV * u.sum(axis=(0, 1, 2))
    for i in range(2):                                             # pde.jl:204 (M hard-wired, Q13)
        for j in range(L):
            rho = u[:, :, i, j]
            fx = rho * (a * F[:, :, 0] + b * (y[0, j] - X) + c * (z[0, i] - X))   # pde.jl:214-215
            fy = rho * (a * F[:, :, 1] + b * (y[1, j] - Y) + c * (z[1, i] - Y))
            dive = C @ fx + fy @ C.T                               # pde.jl:216-217
            reac = np.zeros((N_x, N_y))
            for j2 in range(L):                                    # pde.jl:220-225
                if j2 != j:
                    reac += -G[:, :, i, j2] * rho + G[:, :, i, j] * u[:, :, i, j2]
            dif = D * (Mm @ rho + rho @ Mm.T)                      # pde.jl:228
            dif[0, :] += -D / dx * fx[0, :]                        # pde.jl:231-234 (Q12)
            dif[-1, :] += D / dx * fx[-1, :]
            dif[:, 0] += -D / dy_ * fy[:, 0]
            dif[:, -1] += D / dy_ * fy[:, -1]
            du[:, :, i, j] = dif - dive + reac                     # pde.jl:236
    for j in range(L):                                             # pde.jl:239-255
        ci = controlled_inf[j]
        if ci == 0:
            rj = u[:, :, :, j].sum(axis=2)
            with np.errstate(invalid="ignore", divide="ignore"):
                mean = 1 / m_j[j] * dV * np.array([(rj * X).sum(), (rj * Y).sum()])
            dyv[:, j] = 1 / q["frictionI"] * (mean - y[:, j])
        elif ci in (1, 3):
            tgt = np.ravel(q["controltarget1"] if ci == 1 else q["controltarget2"])
            spd = q["controlspeed1"] if ci == 1 else q["controlspeed2"]
            nrm = np.linalg.norm(tgt - y[:, j])
            # pde.jl:244,250 test `> 0`; AT_TARGET guards the rounding-decided case at arrival (see pde.cu)
            dyv[:, j] = spd * (tgt - y[:, j]) / nrm if nrm > AT_TARGET else 0.0
    for i in range(2):                                             # pde.jl:256-265
        if controlled_med[i] == 1:
            tgt = np.ravel(q["controltarget1"])
            nrm = np.linalg.norm(tgt - z[:, i])
            dz[:, i] = q["controlspeed1"] * (tgt - z[:, i]) / nrm if nrm > AT_TARGET else 0.0
        else:
            ri = u[:, :, i, :].sum(axis=2)
            mean = 1 / m_i[i] * dV * np.array([(ri * X).sum(), (ri * Y).sum()])
            dz[:, i] = 1 / q["frictionM"] * (mean - z[:, i])
    if parts:
        return du, dz, dyv, dict(F=F, G=G, rhosum=rhosum)
    return du, dz, dyv


# ---- fixed-step explicit Runge-Kutta tableaus shared with the CUDA integrator -------------
SCHEMES = {
    "euler": dict(a=[[]], b=[1.0], c=[0.0]),
    "heun": dict(a=[[], [1.0]], b=[0.5, 0.5], c=[0.0, 1.0]),
    "ssprk3": dict(a=[[], [1.0], [0.25, 0.25]], b=[1 / 6, 1 / 6, 2 / 3], c=[0.0, 1.0, 0.5]),
    "rk4": dict(a=[[], [0.5], [0.0, 0.5], [0.0, 0.0, 1.0]], b=[1 / 6, 1 / 3, 1 / 3, 1 / 6],
                c=[0.0, 0.5, 0.5, 1.0]),
}


def rk_step(u, z, y, dt, p, q, ci=None, cm=None, scheme="rk4", force="toeplitz"):
    tab = SCHEMES[scheme]
    ks = []
    for s in range(len(tab["b"])):
        us, zs, ys = u.copy(), z.copy(), y.copy()
        for r_, coef in enumerate(tab["a"][s]):
            if coef != 0.0:
                us += dt * coef * ks[r_][0]; zs += dt * coef * ks[r_][1]; ys += dt * coef * ks[r_][2]
        ks.append(rhs(us, zs, ys, p, q, ci, cm, force=force))
    un, zn, yn = u.copy(), z.copy(), y.copy()
    for s, bs in enumerate(tab["b"]):
        un += dt * bs * ks[s][0]; zn += dt * bs * ks[s][1]; yn += dt * bs * ks[s][2]
    return un, zn, yn


def rk_solve(u0, z0, y0, nsteps, dt, p, q, ci=None, cm=None, scheme="rk4", save_every=None, force="toeplitz"):
    """Fixed-step integration; returns final state, or list of (step,u,z,y) if save_every."""
    u, z, y = u0.copy(), z0.copy(), y0.copy()
    saved = [(0, u.copy(), z.copy(), y.copy())]
    for k in range(1, nsteps + 1):
        u, z, y = rk_step(u, z, y, dt, p, q, ci, cm, scheme, force)
        if save_every and k % save_every == 0:
            saved.append((k, u.copy(), z.copy(), y.copy()))
    return (u, z, y) if not save_every else saved


def adaptive_solve(u0, z0, y0, tmax, p, q, ci=None, cm=None, rtol=1e-8, atol=1e-10, method="DOP853", stats=None):
    """Independent adaptive integrators on the oracle right-hand side (SciPy): DOP853 at a tight tolerance shows that a
    fixed step is converged; method="RK23" is SciPy's own Bogacki-Shampine 3(2) pair (4 stages, FSAL, its own step
    controller) -- an implementation of the same family that shares no code or controller with csrc/pde.cu."""
    from scipy.integrate import solve_ivp
    shp = u0.shape
    nu = u0.size

    def fun(_, v):
        u = v[:nu].reshape(shp); z = v[nu:nu + 4].reshape(2, 2); y = v[nu + 4:].reshape(2, shp[3])
        du, dz, dy = rhs(u, z, y, p, q, ci, cm)
        return np.concatenate([du.ravel(), dz.ravel(), dy.ravel()])

    v0 = np.concatenate([u0.ravel(), z0.ravel(), y0.ravel()])
    sol = solve_ivp(fun, (0, tmax), v0, method=method, rtol=rtol, atol=atol)
    if stats is not None:
        stats.update(nfev=int(sol.nfev), steps=int(sol.t.size - 1))
    v = sol.y[:, -1]
    return v[:nu].reshape(shp), v[nu:nu + 4].reshape(2, 2), v[nu + 4:].reshape(2, shp[3])


# ---- adaptive embedded 3(2) pair: the controller of csrc/pde.cu restated ------------------
def bs3_adaptive(u0, z0, y0, savetimes, p, q, ci=None, cm=None, abstol=1e-6, reltol=1e-3, dt0=0.0, dtmin=0.0,
                 dtmax=0.0, force_dtmin=False, force="toeplitz", max_attempts=10 ** 6):
    """Adaptive integration with the semantics of the reference's solve calls (pde.jl:291;
    controlstrategies.jl:10,96: abstol, reltol, dtmin, force_dtmin, saveat) on the explicit 3-stage
    Bogacki-Shampine scheme c=(0,1/2,3/4), b=(2/9,1/3,4/9) with the midpoint solution b*=(0,1,0) as error
    estimator: norm = sqrt(mean((err/(abstol+reltol*max(|u_n|,|u_n+1|)))^2)) over all density values and ODE
    unknowns, accept if norm<=1, next dt = dt*clamp(fac, 0.2, 5) with fac = 0.9 norm^(-7/30) prev^(2/15) after an accepted step (PI control, prev = norm
    of the previous accepted step) and 0.9 norm^(-1/3) after a rejected one (no growth right after a rejection), steps clipped to land on the
    save times.  (DifferentialEquations.jl itself is absent from the reference tree; only its call-site
    keywords are restated.)  Returns (list of (u,z,y) at savetimes, log of (t, dt, norm, accepted))."""
    u, z, y = u0.copy(), z0.copy(), y0.copy()
    D = 0.5 * float(q["sigma"]) ** 2
    dt_prop = dt0 if dt0 > 0 else (0.1 * min(p["dx"], p["dy"]) ** 2 / D if D > 0 else 1e-3)
    t, out, log, attempts, rejected_last, errprev = 0.0, [], [], 0, False, 1e-4
    ntot = u.size + z.size + y.size
    for tend in savetimes:
        if dtmax > 0:
            dt_prop = min(dt_prop, dtmax)
        dt_prop = max(dt_prop, dtmin)
        while tend - t > 1e-13 * max(1.0, abs(tend)):
            rem = tend - t
            clipped = dt_prop >= rem
            dt = rem if clipped else dt_prop
            k1 = rhs(u, z, y, p, q, ci, cm, force=force)
            k2 = rhs(u + 0.5 * dt * k1[0], z + 0.5 * dt * k1[1], y + 0.5 * dt * k1[2], p, q, ci, cm, force=force)
            k3 = rhs(u + 0.75 * dt * k2[0], z + 0.75 * dt * k2[1], y + 0.75 * dt * k2[2], p, q, ci, cm, force=force)
            new = [w + dt * (2 / 9 * a + 1 / 3 * b + 4 / 9 * c) for w, a, b, c in zip((u, z, y), k1, k2, k3)]
            err = [dt * (2 / 9 * a + (1 / 3 - 1.0) * b + 4 / 9 * c) for a, b, c in zip(k1, k2, k3)]
            tot = sum(float(np.sum((e / (abstol + reltol * np.maximum(np.abs(w), np.abs(n)))) ** 2))
                      for e, w, n in zip(err, (u, z, y), new))
            norm = np.sqrt(tot / ntot)
            at_min = dt <= dtmin * (1 + 1e-12)
            accept = bool(norm <= 1.0) or (force_dtmin and at_min)
            if norm != norm:
                fac = 0.2
            elif norm > 0:
                if accept:   # PI control on accepted steps (beta1 = 7/30, beta2 = 2/15), I control on rejected ones
                    fac = 0.9 * norm ** (-7.0 / 30.0) * errprev ** (2.0 / 15.0)
                else:
                    fac = 0.9 * norm ** (-1.0 / 3.0)
                fac = min(5.0, max(0.2, fac))
            else:
                fac = 5.0
            if rejected_last:
                fac = min(fac, 1.0)      # no growth right after a rejection
            rejected_last = not accept
            log.append((t, dt, norm, accept))
            if accept:
                errprev = max(norm, 1e-4)
                t = tend if clipped else t + dt
                u, z, y = new
                dt_prop = max(dt_prop, dt * fac) if clipped else dt * fac
            else:
                dt_prop = dt * min(fac, 0.9)
                if at_min and not force_dtmin:
                    raise RuntimeError("error test failed at dt = dtmin")
            if dtmax > 0:
                dt_prop = min(dt_prop, dtmax)
            dt_prop = max(dt_prop, dtmin)
            attempts += 1
            if attempts > max_attempts:
                raise RuntimeError("too many attempts")
        t = tend
        out.append((u.copy(), z.copy(), y.copy()))
    return out, log
