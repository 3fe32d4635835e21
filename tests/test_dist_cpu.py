"""world_size-2 gloo test of the multi-GPU host logic (sharding + the single ensemble-mean
all-reduce).  The CUDA runner is replaced by a deterministic stand-in so the test runs without a
GPU; on the B200 box the same code path runs with NCCL (bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from socialmediamodel_b200 import dist as D


def test_shard_range_partitions_exactly():
    for N in (1, 7, 10, 1024, 1025):
        for world in (1, 2, 3, 8):
            spans = [D.shard_range(N, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == N
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _fake_runner(NT, R, *, p, q, savesteps, sigma, first_realization, seed, precision, want):
    Nx, Ny = p.X.shape
    sp = len(savesteps)
    us = np.zeros((Nx, Ny, 2, q.L, sp, R), order="F")
    ys = np.zeros((2, q.L, sp, R), order="F"); zs = np.zeros((2, 2, sp, R), order="F")
    for k in range(R):
        g = first_realization + k                      # global 1-based member index
        us[..., k] = g; ys[..., k] = 10 * g; zs[..., k] = 100 * g
    numbers = np.full((q.M, q.L, NT + 1), R, np.int64, order="F")
    return dict(us=us, zs=zs, ys=ys, sum_us=us.sum(axis=-1), sum_zs=zs.sum(axis=-1), sum_ys=ys.sum(axis=-1), numbers=numbers)


def _worker(rank, world, port, N, q_out):
    import torch.distributed as dist

    import socialmediamodel_b200 as smm
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = smm.ABMconstruct(dx=0.55)
    q = smm.parameters(n=10)
    res = D.sharded_abm_ensemble(8, N, p=p, q=q, savepoints=2, runner=_fake_runner)
    q_out.put((rank, res["lo"], res["hi"], float(res["means"]["av_u"].mean()), float(res["means"]["av_y"].mean()),
               float(res["means"]["av_z"].mean()), int(res["means"]["numbers"][0, 0, 0]), res["us"].shape[-1]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("N", [7, 1])
def test_sharded_ensemble_mean_two_ranks_gloo(N):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q_out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, N, q_out)) for r in range(2)]
    for pr in procs: pr.start()
    got = sorted(q_out.get(timeout=120) for _ in range(2))
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    mean_g = (N + 1) / 2                                 # mean of the member indices 1..N
    for rank, lo, hi, au, ay, az, num, nloc in got:
        assert (lo, hi) == D.shard_range(N, rank, 2) and nloc == hi - lo
        assert abs(au - mean_g) < 1e-12 and abs(ay - 10 * mean_g) < 1e-12 and abs(az - 100 * mean_g) < 1e-12
        assert num == N                                  # follower numbers are summed over ALL members


def _fake_pde_runner(p, qloc, U, Z, Y, times, **kw):
    B, sp, L = len(qloc), len(times), int(qloc[0].L)
    us = np.zeros((p.N_x, p.N_y, 2, L, sp, B), order="F"); zs = np.zeros((2, 2, sp, B), order="F"); ys = np.zeros((2, L, sp, B), order="F")
    for k, qq in enumerate(qloc):                      # the "solution" is the swept parameter itself
        us[..., k] = qq.a; zs[..., k] = 10 * qq.a; ys[..., k] = 100 * qq.a
    return dict(us=us, zs=zs, ys=ys, sum_us=us.sum(axis=-1), sum_zs=zs.sum(axis=-1), sum_ys=ys.sum(axis=-1))


def _pde_worker(rank, world, port, N, q_out):
    import torch.distributed as dist

    import socialmediamodel_b200 as smm
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = smm.PDEconstructcoarse()
    qs = [smm.parameters(a=float(k + 1)) for k in range(N)]
    uzy0, _ = smm.constructinitial("4inf", (p, qs[0]))
    res = D.sharded_pde_ensemble(0.1, qs, p=p, u0=uzy0.x[0], z0=uzy0.x[1], y0=uzy0.x[2], savepoints=3, runner=_fake_pde_runner)
    q_out.put((rank, res["lo"], res["hi"], float(res["means"]["av_u"].mean()), float(res["means"]["av_z"].mean()),
               float(res["means"]["av_y"].mean()), res["us"].shape[-1], len(res["t"])))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("N", [5, 1])
def test_sharded_pde_sweep_two_ranks_gloo(N):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q_out = ctx.Queue()
    procs = [ctx.Process(target=_pde_worker, args=(r, 2, port, N, q_out)) for r in range(2)]
    for pr in procs: pr.start()
    got = sorted(q_out.get(timeout=120) for _ in range(2))
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    mean_a = (N + 1) / 2
    for rank, lo, hi, au, az, ay, nloc, nt in got:
        assert (lo, hi) == D.shard_range(N, rank, 2) and nloc == hi - lo and nt == 3
        assert abs(au - mean_a) < 1e-12 and abs(az - 10 * mean_a) < 1e-12 and abs(ay - 100 * mean_a) < 1e-12
