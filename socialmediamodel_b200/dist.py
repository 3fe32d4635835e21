"""One process per GPU: shard the realizations / PDE instances of an ensemble over the ranks of a
torch.distributed job and reduce the ensemble means (plotting.jl:361-363) with ONE all-reduce.

The ensemble members are independent (ensemble.jl:24-25,69-70), so there is no data-path
collective: rank r runs the contiguous block of members returned by ``shard_range`` and writes its
own slice; only the sums that form the ensemble mean cross NVLink (NCCL on GPUs; gloo in the CPU
tests of this host-side logic)."""
import numpy as np


def shard_range(N, rank, world):
    """Contiguous block [lo, hi) of the N members owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(N, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def rank_world():
    d = _dist()
    return (d.get_rank(), d.get_world_size()) if d else (0, 1)


def allreduce_sums(arrays, device=None):
    """In-place SUM all-reduce of a list of NumPy arrays (Fortran or C order) as one flat buffer."""
    d = _dist()
    if d is None:
        return arrays
    import torch
    flat = np.concatenate([np.ravel(a, order="K") for a in arrays])
    t = torch.from_numpy(flat)
    if device is not None:
        t = t.to(device)
    d.all_reduce(t, op=d.ReduceOp.SUM)
    flat = t.cpu().numpy()
    off = 0
    for a in arrays:
        n = a.size
        np.copyto(a, flat[off:off + n].reshape(a.shape, order="F" if a.flags.f_contiguous else "C"))
        off += n
    return arrays


def sharded_abm_ensemble(NT, N, *, p, q, savepoints=4, sigma=0.1, precision="f32", seed=0, runner=None, device=None):
    """ABMsolveensemble over all ranks.  Returns the local slice (us, zs, ys for members lo..hi-1,
    global realization indices lo+1..hi as in ensemble.jl:25) and the GLOBAL ensemble means.
    `runner(NT, R, first_realization=..., ...)` defaults to ensemble.abm_ensemble_raw (the CUDA path);
    the CPU tests inject a stand-in to exercise the sharding/reduction logic under gloo."""
    from .ensemble import savetimes_abm
    if runner is None:
        from .ensemble import abm_ensemble_raw as runner
    rank, world = rank_world()
    lo, hi = shard_range(N, rank, world)
    st = savetimes_abm(NT, savepoints)
    Nx, Ny = p.X.shape
    if hi > lo:
        out = runner(NT, hi - lo, p=p, q=q, savesteps=st, sigma=sigma, first_realization=lo + 1, seed=seed,
                     precision=precision, want=("us", "zs", "ys", "sums", "numbers"))
    else:  # more ranks than members: contribute zeros
        out = dict(us=np.zeros((Nx, Ny, 2, q.L, len(st), 0), order="F"), zs=np.zeros((2, 2, len(st), 0), order="F"),
                   ys=np.zeros((2, q.L, len(st), 0), order="F"),
                   sum_us=np.zeros((Nx, Ny, 2, q.L, len(st)), order="F"), sum_zs=np.zeros((2, 2, len(st)), order="F"),
                   sum_ys=np.zeros((2, q.L, len(st)), order="F"), numbers=np.zeros((q.M, q.L, NT + 1), np.int64, order="F"))
    numbers = out["numbers"].astype(np.float64)
    allreduce_sums([out["sum_us"], out["sum_zs"], out["sum_ys"], numbers], device)
    means = dict(av_u=out["sum_us"] / N, av_z=out["sum_zs"] / N, av_y=out["sum_ys"] / N,
                 numbers=np.round(numbers).astype(np.int64))
    return dict(lo=lo, hi=hi, us=out["us"], zs=out["zs"], ys=out["ys"], means=means)


def sharded_pde_ensemble(tmax, qs, *, p, u0, z0, y0, savepoints=4, precision="f32", adaptive=False, dt=None, runner=None,
                         device=None, **solver_kw):
    """PDEsolveensemble (ensemble.jl:58-85) as a parameter sweep over all ranks: rank r integrates the contiguous
    block of the instances ``qs`` returned by ``shard_range`` as one device-resident batch, saved at
    LinRange(0, tmax, savepoints); the ensemble means (plotting.jl:361-363) are one all-reduce of the per-rank sums.
    ``u0/z0/y0`` are the shared initial state (one instance).  `runner(p, qs, U, Z, Y, times, ...)` returning
    dict(us, zs, ys, sum_us, sum_zs, sum_ys) defaults to the CUDA solvers; the CPU tests inject a stand-in."""
    rank, world = rank_world()
    N = len(qs)
    lo, hi = shard_range(N, rank, world)
    times = np.linspace(0.0, tmax, savepoints) if savepoints > 1 else np.array([0.0])
    L = int(qs[0].L)
    Nx, Ny = p.N_x, p.N_y
    if runner is None:
        def runner(p, qloc, U, Z, Y, times, **kw):
            from .pde import default_dt, pde_solve_adaptive_raw, pde_solve_raw
            if adaptive:
                return pde_solve_adaptive_raw(p, qloc, U, Z, Y, savetimes=times, precision=precision, want_sums=True, **solver_kw)
            h = min(default_dt(p, qq) for qq in qloc) if dt is None else dt
            per = max(1, int(np.ceil(tmax / max(savepoints - 1, 1) / h - 1e-9)))
            nsteps = per * max(savepoints - 1, 1)
            st = (np.arange(savepoints) * per).astype(np.int32) if savepoints > 1 else np.array([0], np.int32)
            return pde_solve_raw(p, qloc, U, Z, Y, dt=tmax / nsteps, nsteps=nsteps, savesteps=st, precision=precision,
                                 want_sums=True, **solver_kw)
    if hi > lo:
        B = hi - lo
        U = np.stack([u0] * B, axis=-1); Z = np.stack([z0] * B, axis=-1); Y = np.stack([y0] * B, axis=-1)
        out = runner(p, list(qs[lo:hi]), U, Z, Y, times)
    else:  # more ranks than instances: contribute zeros
        out = dict(us=np.zeros((Nx, Ny, 2, L, len(times), 0), order="F"), zs=np.zeros((2, 2, len(times), 0), order="F"),
                   ys=np.zeros((2, L, len(times), 0), order="F"), sum_us=np.zeros((Nx, Ny, 2, L, len(times)), order="F"),
                   sum_zs=np.zeros((2, 2, len(times)), order="F"), sum_ys=np.zeros((2, L, len(times)), order="F"))
    allreduce_sums([out["sum_us"], out["sum_zs"], out["sum_ys"]], device)
    means = dict(av_u=out["sum_us"] / N, av_z=out["sum_zs"] / N, av_y=out["sum_ys"] / N)
    return dict(lo=lo, hi=hi, t=times, us=out["us"], zs=out["zs"], ys=out["ys"], means=means)
