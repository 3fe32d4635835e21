"""Size-independent properties at BASELINE.json's FULL sizes, where the oracle is too slow to run beside the CUDA path
(the per-value parity against the oracle is established at small sizes by the other -m gpu modules):

  C2  1 024 realizations x 200 steps x 250 agents (ensemble.jl:11-55): agents are conserved per step (follower numbers,
      plotting.jl:284-307), every KDE integrates to the mass that lies inside the grid, the run is bit-reproducible, and a
      shard of the ensemble equals the same realizations inside the whole ensemble (seed = global index, ensemble.jl:25)
  C4  instances at the 4x refined grid 353 x 353 (ensemble.jl:58-85): an instance of a batch = its solo run bit for bit
      (instances are independent, ensemble.jl:69-70); the interaction force is invariant under scaling the density
      (pde.jl:58: force ./ norms); total reaction is zero (pde.jl:220-225)
  C5  10^7 agents, 64 influencers, CSR network: agents conserved, positions inside the domain (abm.jl:139-150), two runs
      give the same bits (order-independent fixed-point class sums)
"""
import numpy as np
import pytest

import socialmediamodel_b200 as smm

pytestmark = pytest.mark.gpu


def test_c2_full_ensemble_properties(monkeypatch):
    NT, R, sp = 200, 1024, 4
    p, q = smm.ABMconstruct(), smm.parameters()
    st = smm.savetimes_abm(NT, sp)
    assert list(st) == [0, 66, 133, 199]                                           # ensemble.jl:21: 1,67,134,200 (1-based)
    a = smm.abm_ensemble_raw(NT, R, p=p, q=q, savesteps=st, first_realization=1, precision="f32", want=("zs", "ys", "sums", "numbers"))
    b = smm.abm_ensemble_raw(NT, R, p=p, q=q, savesteps=st, first_realization=1, precision="f32", want=("zs", "ys", "sums", "numbers"))
    for k in ("sum_us", "sum_ys", "sum_zs", "numbers", "ys"):
        np.testing.assert_array_equal(a[k], b[k])                                  # bit-reproducible
    np.testing.assert_array_equal(a["numbers"].sum(axis=(0, 1)), np.full(NT + 1, R * q.n))   # agents conserved at every step
    mass = a["sum_us"].sum(axis=(0, 1, 2, 3)) * p.dV / R
    assert np.all(mass > 0.97) and np.all(mass < 1.0 + 1e-6), mass                 # Gaussians leak a little over the domain edge
    assert np.all(a["sum_us"] >= 0)
    # the full ensemble runs every realization as a chain of 20-step CTAs (state handed over through global memory, so that
    # the SMs stay full to the end): same bits as one CTA per realization for the whole run
    monkeypatch.setenv("SMM_ABM_CHUNK_STEPS", "0")
    c = smm.abm_ensemble_raw(NT, R, p=p, q=q, savesteps=st, first_realization=1, precision="f32", want=("zs", "ys", "sums", "numbers"))
    monkeypatch.delenv("SMM_ABM_CHUNK_STEPS")
    for k in ("sum_us", "sum_ys", "sum_zs", "numbers", "ys", "zs"):
        np.testing.assert_array_equal(a[k], c[k])
    # realizations 513..640 as their own shard = the same realizations inside the ensemble
    sh = smm.abm_ensemble_raw(NT, 128, p=p, q=q, savesteps=st, first_realization=513, precision="f32", want=("ys", "zs"))
    np.testing.assert_array_equal(sh["ys"], a["ys"][..., 512:640])
    np.testing.assert_array_equal(sh["zs"], a["zs"][..., 512:640])
    # influencers start at their followers' means in the four quadrants and stay inside the domain
    assert np.abs(a["ys"]).max() <= 2.2 and np.isfinite(a["ys"]).all()


def test_c4_refined_grid_properties(monkeypatch):
    # the same kernels for the batch and the solo runs (kernel selection depends on the batch size): the throughput path
    monkeypatch.setenv("SMM_FFT_FAST", "1")
    monkeypatch.setenv("SMM_PDE_STAGE2", "1")
    p = smm.PDEconstruct(dx=0.0125)
    assert (p.N_x, p.N_y) == (353, 353)
    qs = [smm.parameters(a=0.5, sigma=0.75), smm.parameters(a=2.0, b=6.0, eta=1.0), smm.parameters(c=4.0, sigma=0.25)]
    inits = [smm.constructinitial("4inf", (p, q)) for q in qs]
    ql = [qq for _, qq in inits]
    rng = np.random.default_rng(3)
    U = np.stack([uz.x[0] * (1 + 0.2 * rng.random(uz.x[0].shape)) for uz, _ in inits], axis=-1)
    Z = np.stack([uz.x[1] for uz, _ in inits], axis=-1); Y = np.stack([uz.x[2] for uz, _ in inits], axis=-1)
    dt = 5e-5
    batch = smm.pde_solve_raw(p, ql, U, Z, Y, dt=dt, nsteps=8, savesteps=[8], precision="f32")
    for b in range(3):                                                             # instance of a batch == solo run
        solo = smm.pde_solve_raw(p, [ql[b]], U[..., b:b + 1], Z[..., b:b + 1], Y[..., b:b + 1], dt=dt, nsteps=8, savesteps=[8],
                                 precision="f32")
        np.testing.assert_array_equal(solo["us"][..., 0, 0], batch["us"][..., 0, b])
        np.testing.assert_array_equal(solo["ys"][..., 0, 0], batch["ys"][..., 0, b])
    # reaction moves mass between influencers only: the rate of change of sum_j u, summed over the grid, is the flux part
    # alone, and the whole right-hand side is invariant under nothing else we can check cheaply -- but the interaction
    # force is scale-invariant (F = (K rho) / (W rho)): doubling the density doubles du exactly where only advection acts
    du1, _, _ = smm.pde_rhs(U[..., 0], Z[..., 0], Y[..., 0], p=p, q=ql[0].merge(eta=0.0), precision="f32")
    du2, _, _ = smm.pde_rhs(2 * U[..., 0], Z[..., 0], Y[..., 0], p=p, q=ql[0].merge(eta=0.0), precision="f32")
    assert np.abs(du2 - 2 * du1).max() <= 2e-6 * np.abs(du2).max()
    # total reaction is zero at every grid point (pde.jl:220-225): with and without it the sum over influencers agrees
    dur, _, _ = smm.pde_rhs(U[..., 0], Z[..., 0], Y[..., 0], p=p, q=ql[0], precision="f32")
    assert np.abs(dur.sum(axis=3) - du1.sum(axis=3)).max() <= 1e-5 * np.abs(du1).max()
    assert np.isfinite(batch["us"]).all()


def test_c4_queued_steps_reproducible_and_close_to_float64(monkeypatch):
    """Eight RK4 steps queued back to back (no host synchronisation in between) on the throughput path: three runs give the
    same bits and stay within 2e-6 of the Float64 kernels.  This is the configuration in which a missing barrier between the
    table loads and the first pass of the fast FFT kernels showed up as run-to-run differences of up to 7e-5;
    with one step per call it never did."""
    monkeypatch.setenv("SMM_FFT_FAST", "1")
    monkeypatch.setenv("SMM_PDE_STAGE2", "1")
    monkeypatch.setenv("SMM_PDE_PDL", "0")       # programmatic dependent launch hid the race: test without it
    p = smm.PDEconstruct(dx=0.0125)
    qs = [smm.parameters(a=0.5, sigma=0.75), smm.parameters(a=2.0, b=6.0, eta=1.0), smm.parameters(c=4.0, sigma=0.25)]
    inits = [smm.constructinitial("4inf", (p, q)) for q in qs]
    ql = [qq for _, qq in inits]
    rng = np.random.default_rng(3)
    U = np.stack([uz.x[0] * (1 + 0.2 * rng.random(uz.x[0].shape)) for uz, _ in inits], axis=-1)
    Z = np.stack([uz.x[1] for uz, _ in inits], axis=-1); Y = np.stack([uz.x[2] for uz, _ in inits], axis=-1)
    runs = [smm.pde_solve_raw(p, ql, U, Z, Y, dt=5e-5, nsteps=8, savesteps=[8], precision="f32")["us"] for _ in range(3)]
    np.testing.assert_array_equal(runs[1], runs[0])
    np.testing.assert_array_equal(runs[2], runs[0])
    ref = smm.pde_solve_raw(p, ql, U, Z, Y, dt=5e-5, nsteps=8, savesteps=[8], precision="f64")["us"]
    assert np.abs(runs[0] - ref).max() <= 2e-6 * np.abs(ref).max()


def test_c5_full_population_properties():
    import torch
    n, L, k = 10_000_000, 64, 16
    p, q = smm.ABMconstruct(), smm.parameters(n=n, L=L)
    rng = np.random.default_rng(0)
    x0 = rng.uniform(-2.2, 2.2, (n, 2))
    state = np.where(rng.random(n) > 0.5, 1, -1).astype(np.int32)
    g = 8
    cx = np.minimum(g - 1, ((x0[:, 0] + 2.2) / 4.4 * g).astype(np.int64)); cy = np.minimum(g - 1, ((x0[:, 1] + 2.2) / 4.4 * g).astype(np.int64))
    follow0 = (cy * g + cx).astype(np.int32)
    cnt = np.bincount(follow0, minlength=L)
    inf0 = np.stack([np.bincount(follow0, weights=x0[:, d], minlength=L) / cnt for d in range(2)], axis=1)
    media0 = np.array([[-1.0, -1.0], [1.0, 1.0]])
    gen = torch.Generator(device="cuda"); gen.manual_seed(0)
    nb = torch.randint(0, n - 1, (n, k), device="cuda", generator=gen, dtype=torch.int32)
    me = torch.arange(n, device="cuda", dtype=torch.int32)[:, None]
    cols = torch.cat([nb + (nb >= me).to(torch.int32), me], dim=1).contiguous().cpu().numpy().reshape(-1)
    rowptr = np.arange(0, n * (k + 1) + 1, k + 1, dtype=np.int32)
    outs = []
    for _ in range(2):
        pop = smm.ABMPopulation(q, p, x0, state, follow0, inf0, media0, (rowptr, cols), seed=0, precision="f32")
        pop.step(5)
        outs.append(pop.read())
        pop.close()
    a, b = outs
    np.testing.assert_array_equal(a["follow"], b["follow"])                        # bit-reproducible
    np.testing.assert_array_equal(a["x"], b["x"])
    np.testing.assert_array_equal(a["inf"], b["inf"])
    assert int(a["numbers"].sum()) == n                                            # agents conserved
    np.testing.assert_array_equal(a["numbers"].sum(axis=0), np.bincount(a["follow"], minlength=L))
    assert np.abs(a["x"]).max() <= 2.2 + 1e-6 and np.isfinite(a["x"]).all()        # single reflection keeps everybody inside
    assert (a["follow"] != follow0).sum() > 1000                                   # the run exercises switching
