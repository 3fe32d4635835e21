"""Parameter constructors of the reference (src/setting.jl), same names, keywords and defaults.

The reference returns Julia NamedTuples that are splatted and merged (pde.jl:178); ``Struct`` is
the Python equivalent (immutable attribute access + ``merge``).
"""
import numpy as np


class Struct(dict):
    """NamedTuple-like: q.n, q["n"], q.merge(eta=1.0)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        raise AttributeError("Struct is immutable; use .merge(...)")

    def merge(self, **kw):
        d = dict(self)
        d.update(kw)
        return Struct(d)


def parameters(*, L=4, M=2, n=250, eta=15.0, a=1.0, b=4.0, c=2.0, sigma=0.5, sigmahat=0.0, sigmatilde=0.0,
               frictionI=10.0, frictionM=100.0, controlspeed1=np.sqrt(1.5 ** 2 + 1.5 ** 2) / 5,
               controltarget1=(1.5, 1.5), controlspeed2=np.sqrt(1.5 ** 2 + 1.5 ** 2) / 5, controltarget2=(1.5, 1.5)):
    """setting.jl:1-22."""
    return Struct(n=n, L=L, M=M, frictionM=frictionM, frictionI=frictionI, a=a, b=b, c=c, eta=eta, sigma=sigma,
                  sigmahat=sigmahat, sigmatilde=sigmatilde, controlspeed1=controlspeed1,
                  controltarget1=np.asarray(controltarget1, float).reshape(1, 2), controlspeed2=controlspeed2,
                  controltarget2=np.asarray(controltarget2, float).reshape(1, 2))


def parametersstronginf():
    """setting.jl:24-27."""
    return parameters(eta=1.0)


def _default_domain():
    return 2.2 * np.array([[-1.0, 1.0], [-1.0, 1.0]])


def _julia_range(a, step, b):
    nsteps = int(np.floor((b - a) / step + 1e-9))
    return a + step * np.arange(nsteps + 1)


def ABMconstruct(*, dt=0.01, dx=0.05, dy=None, domain=None):
    """setting.jl:30-43."""
    dy = dx if dy is None else dy
    domain = _default_domain() if domain is None else np.asarray(domain, float)
    X, Y = np.meshgrid(_julia_range(domain[0, 0], dx, domain[0, 1]), _julia_range(domain[1, 0], dy, domain[1, 1]),
                       indexing="ij")
    return Struct(dt=dt, dx=dx, dy=dy, domain=domain, X=X, Y=Y, dV=dx * dy)


class _PdeP(Struct):
    """PDE ``p`` (setting.jl:64).  K_matrix / W_matrix (pde.jl:8-23) are *not* materialised -- the
    CUDA path generates the Toeplitz kernel on the fly -- but are available lazily on small grids
    for code that inspects them."""

    def __getattr__(self, k):
        if k in ("K_matrix", "W_matrix"):
            if self["N"] > 3600:
                raise MemoryError(f"{k} at N={self['N']} would need {self['N'] ** 2 * 24 / 1e9:.1f} GB; "
                                  "the B200 path never materialises it (Toeplitz kernel generated on the fly)")
            g = self["grid_points"]
            dxm = g[None, :, 0] - g[:, None, 0]
            dym = g[None, :, 1] - g[:, None, 1]
            w = np.exp(-np.sqrt(dxm ** 2 + dym ** 2))
            return np.stack([w * dxm, w * dym], axis=-1) if k == "K_matrix" else w
        return super().__getattr__(k)


def PDEconstruct(*, dx=0.05, dy=None, domain=None):
    """setting.jl:45-67."""
    dy = dx if dy is None else dy
    domain = _default_domain() if domain is None else np.asarray(domain, float)
    N_x = int(np.round((domain[0, 1] - domain[0, 0]) / dx + 1))
    N_y = int(np.round((domain[1, 1] - domain[1, 0]) / dy + 1))
    xa = domain[0, 0] + dx * np.arange(N_x)
    ya = domain[1, 0] + dy * np.arange(N_y)
    X, Y = np.meshgrid(xa, ya, indexing="ij")
    grid_points = np.stack([X.ravel(order="F"), Y.ravel(order="F")], axis=1)
    Mm = (np.diag(np.full(N_x, -2.0)) + np.diag(np.ones(N_x - 1), 1) + np.diag(np.ones(N_x - 1), -1))
    Mm[0, 0] = Mm[-1, -1] = -1.0
    Cm = 1 / (2 * dx) * (np.diag(np.ones(N_x - 1), 1) - np.diag(np.ones(N_x - 1), -1))
    Cm[0, 0:2] = 1 / dx * np.array([-1.0, 1.0])
    Cm[-1, -2:] = 1 / dx * np.array([-1.0, 1.0])
    return _PdeP(grid_points=grid_points, dx=dx, dy=dy, dV=dx * dy, X=X, Y=Y, N=N_x * N_y, N_x=N_x, N_y=N_y,
                 domain=domain, C=Cm, MassM=Mm * (1 / dx ** 2))


def PDEconstructcoarse():
    """setting.jl:69-71."""
    return PDEconstruct(dx=0.1, dy=0.1, domain=2.3 * np.array([[-1.0, 1.0], [-1.0, 1.0]]))


def PDEconstructmeso():
    """setting.jl:73-75."""
    return PDEconstruct(dx=44 / 600, dy=44 / 600)


def parameterscontrol(*, ntarg=3, speedbound=2, Tmax=5, tequil=5.0, dtmin=0.001, start="influencer",
                      maximize="follower", alpha=0.05, mindist=5.0):
    """setting.jl:77-82."""
    return Struct(ntarg=ntarg, speedbound=speedbound, Tmax=Tmax, tequil=tequil, dtmin=dtmin, start=start,
                  maximize=maximize, alpha=alpha, mindist=mindist)
