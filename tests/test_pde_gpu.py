"""GPU parity of the PDE right-hand side and fixed-step integrator against the Float64 oracle,
through the C ABI.  Bar: densities within 1e-5 (FP32 vs the reference's Float64), per evaluation
(relative to max|du|) and at final time (relative to max|u|) against the oracle running the same
RK tableau."""
import numpy as np
import pytest

import socialmediamodel_b200 as smm
from oracle import abm_oracle as A
from oracle import pde_oracle as P
from oracle.abm_oracle import default_q as A_default_q

pytestmark = pytest.mark.gpu

TOL32 = 1e-5
TOL64 = 1e-11


def oracle_p(p):
    return P.pde_construct(dx=p.dx, dy=p.dy, domain=p.domain)


def rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


@pytest.fixture(scope="module")
def coarse():
    p = smm.PDEconstructcoarse()
    q = smm.parameters()
    uzy0, q2 = smm.constructinitial("4inf", (p, q))
    return p, q2, uzy0


def perturbed_state(p, q, uzy0, seed=0):
    rng = np.random.default_rng(seed)
    u0, z0, y0 = uzy0.x
    u = u0 * (1 + 0.3 * rng.random(u0.shape)) + 0.002 * rng.random(u0.shape)
    return u, z0 + 0.1 * rng.standard_normal(z0.shape), y0 + 0.1 * rng.standard_normal(y0.shape)


@pytest.mark.parametrize("precision,tol", [("f64", TOL64), ("f32", TOL32)])
def test_rhs_parity_coarse(coarse, precision, tol):
    p, q, uzy0 = coarse
    po = oracle_p(p)
    for u, z, y in (uzy0.x, perturbed_state(p, q, uzy0)):
        du, dz, dy = smm.pde_rhs(u, z, y, p=p, q=q, precision=precision)
        dur, dzr, dyr = P.rhs(u, z, y, po, dict(q))
        assert rel(du, dur) < tol, rel(du, dur)
        assert np.abs(dz - dzr).max() < max(tol, 1e-9) * max(np.abs(dzr).max(), 1e-3)
        assert np.abs(dy - dyr).max() < max(tol, 1e-9) * max(np.abs(dyr).max(), 1e-3)


def test_rhs_matches_materialised_K_matrix(coarse):
    """The Toeplitz kernel reproduces the reference's dense K_matrix/W_matrix gemv (pde.jl:49-59)."""
    p, q, uzy0 = coarse
    po = P.pde_construct_coarse(materialise=True)
    u, z, y = perturbed_state(p, q, uzy0, 3)
    du, dz, dy = smm.pde_rhs(u, z, y, p=p, q=q, precision="f64")
    dur, _, _ = P.rhs(u, z, y, po, dict(q), force="dense")
    assert rel(du, dur) < 1e-10


def test_rhs_default_grid_f32_and_reference_f_signature():
    p = smm.PDEconstruct()
    q = smm.parameters()
    uzy0, q2 = smm.constructinitial("4inf", (p, q))
    duzy = smm.ArrayPartition(*[np.zeros_like(a) for a in uzy0.x])
    assert smm.f(duzy, uzy0, (p, q2), 0.0) is None
    dur, dzr, dyr = P.rhs(*uzy0.x, oracle_p(p), dict(q))
    assert rel(duzy.x[0], dur) < TOL32
    assert abs(np.abs(dur).max() - 9.61) < 0.02                     # survey probe: max|du| = 9.61 at 89^2
    np.testing.assert_allclose(duzy.x[1], -uzy0.x[1] / 100.0, atol=1e-7)   # dz = -z/frictionM at t = 0


def test_rhs_controlled_L5_zero_mass_and_batch(coarse):
    p, q, uzy0 = coarse
    u0, z0, y0 = uzy0.x
    u = np.concatenate([u0, np.zeros_like(u0[..., :1])], axis=3)
    y = np.concatenate([y0, np.array([[0.2], [-0.3]])], axis=1)
    q5 = q.merge(L=5, controlled_inf=np.array([0, 0, 0, 0, 1.0]), controlled_med=np.array([0.0, 1.0]))
    q6 = q5.merge(controlled_inf=np.array([0, 3.0, 0, 0, 1.0]), controlled_med=np.array([0.0, 0.0]), eta=1.0, a=2.0,
                  controltarget2=np.array([[-1.0, 0.5]]), controlspeed2=0.7)
    po = oracle_p(p)
    uB = np.stack([u, u * 1.1], axis=-1); zB = np.stack([z0, z0], axis=-1); yB = np.stack([y, y], axis=-1)
    du, dz, dy = smm.pde_rhs(uB, zB, yB, p=p, q=[q5, q6], precision="f64")
    for k, (qq, uu) in enumerate(((q5, u), (q6, u * 1.1))):
        dur, dzr, dyr = P.rhs(uu, z0, y, po, dict(qq), qq.controlled_inf, qq.controlled_med)
        assert rel(du[..., k], dur) < TOL64
        np.testing.assert_allclose(dy[..., k], dyr, atol=1e-12)
        np.testing.assert_allclose(dz[..., k], dzr, atol=1e-12)
    assert np.isfinite(du).all()


@pytest.mark.parametrize("scheme", ["euler", "heun", "ssprk3", "rk4"])
def test_fixed_step_trajectory_parity_f64(coarse, scheme):
    p, q, uzy0 = coarse
    po = oracle_p(p)
    u0, z0, y0 = uzy0.x
    dt, nsteps = 1e-3, 6
    out = smm.pde_solve_raw(p, [q], u0, z0, y0, dt=dt, nsteps=nsteps, savesteps=[0, 3, 6], scheme=scheme, precision="f64")
    saved = P.rk_solve(u0, z0, y0, nsteps, dt, po, dict(q), scheme=scheme, save_every=3)
    for m, (k, ur, zr, yr) in enumerate(saved):
        assert rel(out["us"][..., m, 0], ur) < 1e-10, (scheme, k)
        np.testing.assert_allclose(out["zs"][..., m, 0], zr, atol=1e-11)
        np.testing.assert_allclose(out["ys"][..., m, 0], yr, atol=1e-11)


def test_rk4_trajectory_parity_f32_and_convergence(coarse):
    p, q, uzy0 = coarse
    po = oracle_p(p)
    u0, z0, y0 = uzy0.x
    T, n = 0.05, 50
    out = smm.pde_solve_raw(p, [q], u0, z0, y0, dt=T / n, nsteps=n, savesteps=[n], precision="f32")
    ur, zr, yr = P.rk_solve(u0, z0, y0, n, T / n, po, dict(q))
    assert rel(out["us"][..., 0, 0], ur) < TOL32                      # same tableau, FP32 vs Float64
    assert np.abs(out["ys"][..., 0, 0] - yr).max() < 1e-6
    ua, za, ya = P.adaptive_solve(u0, z0, y0, T, po, dict(q))
    assert rel(out["us"][..., 0, 0], ua) < 5e-5                       # fixed step is converged (cf. tests/test_oracle_kats)


def test_pdesolve_reference_api(coarse):
    p, q, _ = coarse
    sol, (p2, q2) = smm.PDEsolve(0.02, p=p, q=smm.parameters(), precision="f64")
    assert sol.t[0] == 0.0 and abs(sol.t[-1] - 0.02) < 1e-15
    assert "controlled_inf" in q2 and "controlled_med" in q2          # pde.jl:178
    u, z, y = smm.sol2uyz(sol, 0.02)
    assert u.shape == (47, 47, 2, 4) and z.shape == (2, 2) and y.shape == (2, 4)
    um = sol(0.013).x[0]                                              # dense output between saved steps
    assert np.isfinite(um).all()
    po = oracle_p(p)
    uzy0, _ = smm.constructinitial("4inf", (p, smm.parameters()))
    n = len(sol.t) - 1
    ur, _, _ = P.rk_solve(*uzy0.x, n, 0.02 / n, po, dict(smm.parameters()))
    assert rel(u, ur) < 1e-10


def test_pdesolveensemble_sweep_matches_single_solves(coarse):
    p, q, _ = coarse
    qs = [smm.parameters(a=a, eta=eta) for a, eta in ((0.5, 15.0), (2.0, 1.0), (1.0, 15.0))]
    us, zs, ys, P_ = smm.PDEsolveensemble(0.012, 3, savepoints=3, p=p, qs=qs, save=False, dt=1e-3, precision="f64")
    assert us.shape == (47, 47, 2, 4, 3, 3) and zs.shape == (2, 2, 3, 3) and ys.shape == (2, 4, 3, 3)
    po = oracle_p(p)
    uzy0, _ = smm.constructinitial("4inf", (p, qs[0]))
    for k, qq in enumerate(qs):
        ur, zr, yr = P.rk_solve(*uzy0.x, 12, 1e-3, po, dict(qq))
        assert rel(us[..., 2, k], ur) < 1e-10
        np.testing.assert_allclose(ys[..., 2, k], yr, atol=1e-11)
    np.testing.assert_array_equal(us[..., 0, 0], uzy0.x[0])            # save point 1 is t = 0 (ensemble.jl:65)


def test_batch_resident_and_sums(coarse):
    import torch
    p, q, uzy0 = coarse
    qs = [q.merge(sigma=s) for s in (0.25, 0.5, 0.75, 1.0)]
    pb = smm.PDEBatch(p, qs, *uzy0.x, precision="f32")
    pb.step(5e-4, 10)
    u, z, y = pb.read()
    s = torch.zeros(u[..., 0].size, dtype=torch.float64, device="cuda")
    pb.sum_dev(s.data_ptr())
    torch.cuda.synchronize()
    np.testing.assert_allclose(s.cpu().numpy().reshape(u[..., 0].shape, order="F"), u.sum(axis=-1), rtol=1e-12, atol=1e-12)
    po = oracle_p(p)
    for k in (0, 3):
        ur, _, _ = P.rk_solve(*uzy0.x, 10, 5e-4, po, dict(qs[k]))
        assert rel(u[..., k], ur) < TOL32
    pb.close()


def test_direct_toeplitz_kernel_agrees_with_fft_path(coarse, monkeypatch):
    """SMM_PDE_CONV=direct selects the O(N^2) smem-tiled Toeplitz kernel; default is the FFT path."""
    p, q, uzy0 = coarse
    u, z, y = perturbed_state(p, q, uzy0, 5)
    du_fft, _, _ = smm.pde_rhs(u, z, y, p=p, q=q, precision="f64")
    monkeypatch.setenv("SMM_PDE_CONV", "direct")
    du_dir, _, _ = smm.pde_rhs(u, z, y, p=p, q=q, precision="f64")
    du32, _, _ = smm.pde_rhs(u, z, y, p=p, q=q, precision="f32")
    monkeypatch.delenv("SMM_PDE_CONV")
    assert rel(du_fft, du_dir) < 1e-11
    dur, _, _ = P.rhs(u, z, y, oracle_p(p), dict(q))
    assert rel(du32, dur) < TOL32


@pytest.mark.parametrize("dx,N", [(0.025, 177), (0.0125, 353)])
def test_rhs_refined_grids_f32(dx, N):
    """C4 grids: the reference cannot build K_matrix here (248 GB at 353^2); parity is against the
    oracle's Float64 FFT evaluation of the same Toeplitz sums."""
    p = smm.PDEconstruct(dx=dx)
    assert p.N_x == N
    q = smm.parameters()
    uzy0, q2 = smm.constructinitial("4inf", (p, q))
    du, dz, dy = smm.pde_rhs(*uzy0.x, p=p, q=q2, precision="f32")
    dur, dzr, dyr = P.rhs(*uzy0.x, oracle_p(p), dict(q))
    assert rel(du, dur) < TOL32, rel(du, dur)
    assert np.abs(dy - dyr).max() < 1e-7 and np.abs(dz - dzr).max() < 1e-7


def test_randominit_uses_gpu_kde_and_matches_oracle():
    """randominit (pde.jl:124-168): density = KDE of sampled agents (GPU kernel K4), y0 = class means."""
    p = smm.PDEconstructcoarse()
    q = smm.parameters(n=80)
    uzy0, ci, cm = smm.randominit((p, q), np.random.default_rng(4))
    rng = np.random.default_rng(4)
    pos = rng.random((80, 2)) * (p.domain[0, 1] - p.domain[0, 0]) + p.domain[0, 0]
    med = rng.integers(0, 2, 80)
    ur, zr, yr, _, _ = P.randominit_from(oracle_p(p), dict(q), pos, med)
    assert rel(uzy0.x[0], ur) < 1e-10
    np.testing.assert_allclose(uzy0.x[2], yr, rtol=1e-12)
    assert abs(uzy0.x[0].sum() * p.dV - 1) < 1e-12
    sol, _ = smm.PDEsolve(0.004, init="random", p=p, q=q, rng=np.random.default_rng(4), precision="f64")
    assert np.isfinite(sol(0.004).x[0]).all()


def test_runensembles_shapes():
    us, zs, ys, (p, q), us2, zs2, ys2, _ = smm.runensembles(3, NT=10, tmax=0.01, savepoints=2, save=False)
    assert us.shape == (89, 89, 2, 4, 2, 1) and us2.shape == (89, 89, 2, 4, 2, 3)
    assert np.isfinite(us).all() and np.isfinite(us2).all()


def test_c3_full_horizon_fp32_drift():
    """C3: PDEsolve on the default 89x89 grid, RK4 dt=1e-3 to tmax=2.0 (2 000 steps).  FP32 state must stay
    within 1e-5 of the Float64 run at the final time (the Float64 path is oracle-exact, see above)."""
    p = smm.PDEconstruct()
    q = smm.parameters()
    uzy0, q2 = smm.constructinitial("4inf", (p, q))
    o32 = smm.pde_solve_raw(p, [q2], *uzy0.x, dt=1e-3, nsteps=2000, savesteps=[200, 2000], precision="f32")
    o64 = smm.pde_solve_raw(p, [q2], *uzy0.x, dt=1e-3, nsteps=2000, savesteps=[200, 2000], precision="f64")
    for m in range(2):
        assert rel(o32["us"][..., m, 0], o64["us"][..., m, 0]) < TOL32
        assert np.abs(o32["ys"][..., m, 0] - o64["ys"][..., m, 0]).max() < 1e-6
    assert abs(o64["us"][..., 1, 0].sum() * p.dV - 1.0113) < 2e-3     # the scheme is not mass-conserving (SURVEY hard part 7)


@pytest.mark.parametrize("dx,N", [(0.1, 47), (0.05, 89), (0.025, 177), (0.0125, 353)])
def test_fast_fft_path_against_generic_path_and_oracle(dx, N, monkeypatch):
    """FP32 interaction force by the register-resident FFT kernels (fft_fast.cu: P = 96 / 192 / 384 / 768, both CTA
    shapes) against the generic shared-memory FFT kernels (fft_conv.cu) on a non-symmetric state, several instances
    (ragged last row / column groups), and against the Float64 oracle."""
    p = smm.PDEconstructcoarse() if N == 47 else smm.PDEconstruct(dx=dx)
    assert p.N_x == N
    rng = np.random.default_rng(N)
    B = 3
    qs = [smm.parameters(a=a) for a in (0.5, 1.0, 2.0)]
    uzy0, _ = smm.constructinitial("4inf", (p, qs[0]))
    u = np.stack([uzy0.x[0] * (1 + 0.3 * rng.random(uzy0.x[0].shape)) for _ in range(B)], axis=-1)
    z = np.stack([uzy0.x[1]] * B, axis=-1)
    y = np.stack([uzy0.x[2] + 0.05 * rng.standard_normal(uzy0.x[2].shape) for _ in range(B)], axis=-1)
    ql = [smm.constructinitial("4inf", (p, q))[1] for q in qs]
    res = {}
    for name, env in (("generic", {"SMM_FFT_FAST": "0"}), ("fast4", {"SMM_FFT_FAST": "1", "SMM_FFT_PAIRS": "4"}),
                      ("fast8", {"SMM_FFT_FAST": "1", "SMM_FFT_PAIRS": "8"}), ("fast4c", {"SMM_FFT_FAST": "1", "SMM_FFT_CHUNK_MB": "1"})):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        res[name] = smm.pde_rhs(u, z, y, p=p, q=ql, precision="f32")
        for k in env:
            monkeypatch.delenv(k)
    ref = res["generic"][0]
    for name in ("fast4", "fast8", "fast4c"):   # fast4c: one instance per chunk
        assert rel(res[name][0], ref) < 2e-6, (name, rel(res[name][0], ref))
        assert np.abs(res[name][2] - res["generic"][2]).max() < 1e-6
    if N <= 177:
        for b in range(B):
            dur, _, _ = P.rhs(u[..., b], z[..., b], y[..., b], oracle_p(p), dict(ql[b]))
            assert rel(res["fast4"][0][..., b], dur) < TOL32


@pytest.mark.parametrize("N", [47, 89, 177])
@pytest.mark.parametrize("scheme", ["rk4", "ssprk3", "heun", "euler"])
def test_packed_stage_kernel_against_generic_and_f64(N, scheme, monkeypatch):
    """pde_stage2_kernel (two grid points per thread, FP32x2; edge-only tiles at 47, interior tiles at 177) against the
    one-point-per-thread kernel and the Float64 path (oracle-exact, see above): right-hand side and 6 steps of every
    scheme (all RK accumulator modes), batch of 2 with different parameters, L = 4."""
    p = smm.PDEconstructcoarse() if N == 47 else smm.PDEconstruct(dx={89: 0.05, 177: 0.025}[N])
    assert p.N_x == N
    rng = np.random.default_rng(N)
    qs = [smm.parameters(a=0.7, sigma=0.4), smm.parameters(a=2.0, eta=5.0)]
    inits = [smm.constructinitial("4inf", (p, q)) for q in qs]
    ql = [qq for _, qq in inits]
    U = np.stack([uz.x[0] * (1 + 0.3 * rng.random(uz.x[0].shape)) for uz, _ in inits], axis=-1)
    Z = np.stack([uz.x[1] for uz, _ in inits], axis=-1); Y = np.stack([uz.x[2] for uz, _ in inits], axis=-1)
    dt = min(smm.default_dt(p, q) for q in ql)
    run = lambda prec: (smm.pde_rhs(U, Z, Y, p=p, q=ql, precision=prec),
                        smm.pde_solve_raw(p, ql, U, Z, Y, dt=dt, nsteps=6, savesteps=[3, 6], scheme=scheme, precision=prec))
    monkeypatch.setenv("SMM_PDE_STAGE2", "0")
    rhs1, sol1 = run("f32")
    monkeypatch.setenv("SMM_PDE_STAGE2", "1")
    rhs2, sol2 = run("f32")
    monkeypatch.delenv("SMM_PDE_STAGE2")
    rhs64, sol64 = run("f64")
    assert rel(rhs2[0], rhs1[0]) < 2e-6 and rel(rhs2[0], rhs64[0]) < TOL32
    assert rel(sol2["us"], sol1["us"]) < 2e-6 and rel(sol2["us"], sol64["us"]) < TOL32
    assert np.abs(sol2["ys"] - sol64["ys"]).max() < 1e-6 and np.abs(sol2["zs"] - sol64["zs"]).max() < 1e-6


def test_packed_stage_kernel_controlled_L5(monkeypatch):
    """L = 5 with a controlled, zero-mass influencer (controlstrategies.jl:16-30; r(NaN) = 0.1, pde.jl:26-32)."""
    from oracle import control_oracle as CO
    p = smm.PDEconstructcoarse()
    q = smm.parameters(eta=1.0)
    po, qo = oracle_p(p), A_default_q(eta=1.0)
    uzy0, sl, q2, ci, cm, slm = CO.prep(0.01, po, qo, start="zero", savedt=0.01, dt=1e-3, countercontrol="no")
    u, z, y = uzy0
    qq = q.merge(L=5, controlled_inf=ci, controlled_med=cm)
    monkeypatch.setenv("SMM_PDE_STAGE2", "1")
    du, dz, dy = smm.pde_rhs(u, z, y, p=p, q=qq, precision="f32")
    monkeypatch.delenv("SMM_PDE_STAGE2")
    dur, dzr, dyr = P.rhs(u, z, y, po, dict(q2), ci, cm)
    assert rel(du, dur) < TOL32 and np.abs(dy - dyr).max() < 1e-6


def test_c4_grid_fp32_drift_500_steps():
    """C4 as bench.py times it (SURVEY hard part 5): 353x353 grid, RK4 with the sweep's dt = 5e-5 (D/dx^2 up to 3 200),
    500 steps, two lattice corners of the parameter sweep (weak / strong diffusion).  FP32 state vs the Float64 run of
    the same scheme (the Float64 path is oracle-exact per evaluation and per step, see above): <= 1e-5 of max|u|."""
    p = smm.PDEconstruct(dx=0.0125)
    assert p.N_x == 353
    qs = [smm.parameters(a=0.5, b=1.0, c=1.0, eta=1.0, sigma=0.25), smm.parameters(a=3.0, b=6.0, c=4.0, eta=15.0, sigma=1.0)]
    inits = [smm.constructinitial("4inf", (p, q)) for q in qs]
    qlist = [qq for _, qq in inits]
    U = np.stack([uz.x[0] for uz, _ in inits], axis=-1)
    Z = np.stack([uz.x[1] for uz, _ in inits], axis=-1)
    Y = np.stack([uz.x[2] for uz, _ in inits], axis=-1)
    dt = min(smm.default_dt(p, q) for q in qlist)
    assert abs(dt - 5e-5) < 1e-12
    o32 = smm.pde_solve_raw(p, qlist, U, Z, Y, dt=dt, nsteps=500, savesteps=[100, 500], precision="f32")
    o64 = smm.pde_solve_raw(p, qlist, U, Z, Y, dt=dt, nsteps=500, savesteps=[100, 500], precision="f64")
    for b in range(2):
        for m in range(2):
            e = rel(o32["us"][..., m, b], o64["us"][..., m, b])
            assert e < TOL32, (b, m, e)
            assert np.abs(o32["ys"][..., m, b] - o64["ys"][..., m, b]).max() < 1e-6
    assert np.isfinite(o64["us"]).all() and rel(o64["us"][..., 1, 0], o64["us"][..., 0, 0]) > 1e-3   # the state moved


def test_pde_against_golden_fixtures():
    """Committed golden vectors (tests/golden/pde_coarse.npz, oracle-generated)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pde_coarse.npz"))
    p, q = smm.PDEconstructcoarse(), smm.parameters()
    for precision, tol in (("f64", TOL64), ("f32", TOL32)):
        du, dz, dy = smm.pde_rhs(g["u"], g["z"], g["y"], p=p, q=q, precision=precision)
        assert rel(du, g["du"]) < tol
        assert np.abs(dy - g["dy"]).max() < max(tol, 1e-9) and np.abs(dz - g["dz"]).max() < max(tol, 1e-9)
        q5 = q.merge(L=5, controlled_inf=g["ci5"], controlled_med=g["cm5"])
        du5, dz5, dy5 = smm.pde_rhs(g["u5"], g["z"], g["y5"], p=p, q=q5, precision=precision)
        assert rel(du5, g["du5"]) < tol and np.abs(dy5 - g["dy5"]).max() < max(tol, 1e-7)
        uzy0, q2 = smm.constructinitial("4inf", (p, q))
        out = smm.pde_solve_raw(p, [q2], *uzy0.x, dt=1e-3, nsteps=10, savesteps=[10], precision=precision)
        assert rel(out["us"][..., 0, 0], g["u_rk4_10"]) < max(tol, 1e-10)


def test_sharded_pde_sweep_single_rank_equals_direct(coarse):
    """dist.sharded_pde_ensemble without a process group = the whole sweep on this GPU; its means are the sums of
    smm_pde_solve divided by the number of instances, fixed-step and adaptive."""
    from socialmediamodel_b200 import dist as D
    p, q, uzy0 = coarse
    qs = [q.merge(a=a) for a in (0.5, 1.0, 1.5)]
    u0, z0, y0 = uzy0.x
    for adaptive in (False, True):
        res = D.sharded_pde_ensemble(0.02, qs, p=p, u0=u0, z0=z0, y0=y0, savepoints=3, precision="f64", adaptive=adaptive)
        assert (res["lo"], res["hi"]) == (0, 3) and res["us"].shape == (p.N_x, p.N_y, 2, 4, 3, 3)
        np.testing.assert_allclose(res["means"]["av_u"], res["us"].sum(axis=-1) / 3, rtol=1e-13, atol=1e-15)
        np.testing.assert_allclose(res["means"]["av_y"], res["ys"].sum(axis=-1) / 3, rtol=1e-13)
        np.testing.assert_array_equal(res["us"][..., 0, 0], u0)
    us, zs, ys, _ = smm.PDEsolveensemble(0.02, 3, savepoints=3, p=p, qs=[smm.parameters(a=a) for a in (0.5, 1.0, 1.5)],
                                         save=False, precision="f64", adaptive=True)
    assert rel(res["us"], us) < 1e-12


# ---- whole RK steps inside one thread-block cluster (csrc/pde_cluster.cu): small FP32 batches ------------------------------
def _cluster(monkeypatch, on):
    monkeypatch.setenv("SMM_PDE_CLUSTER", "1" if on else "0")


@pytest.mark.parametrize("scheme", ["euler", "heun", "ssprk3", "rk4"])
@pytest.mark.parametrize("grid", ["coarse", "default"])
def test_cluster_steps_against_oracle_and_kernel_chain(grid, scheme, monkeypatch):
    """pde.jl:183-268 + the RK tableau, one instance resident in a 16-CTA cluster: against the Float64 oracle on the coarse
    grid, against the Float64 kernel chain on the default grid (the oracle takes minutes there), and against the FP32 chain."""
    p = smm.PDEconstructcoarse() if grid == "coarse" else smm.PDEconstruct()
    uzy0, q = smm.constructinitial("4inf", (p, smm.parameters()))
    u0, z0, y0 = perturbed_state(p, q, uzy0, 5) if grid == "coarse" else uzy0.x
    dt, n = (1e-3, 6)
    _cluster(monkeypatch, True)
    a = smm.pde_solve_raw(p, [q], u0, z0, y0, dt=dt, nsteps=n, savesteps=[0, 3, 6], scheme=scheme, precision="f32")
    launches = smm.Context.default().launches
    smm.pde_solve_raw(p, [q], u0, z0, y0, dt=dt, nsteps=40, savesteps=[40], scheme=scheme, precision="f32")
    per_solve = smm.Context.default().launches - launches
    _cluster(monkeypatch, False)
    c = smm.pde_solve_raw(p, [q], u0, z0, y0, dt=dt, nsteps=n, savesteps=[0, 3, 6], scheme=scheme, precision="f32")
    launches = smm.Context.default().launches
    smm.pde_solve_raw(p, [q], u0, z0, y0, dt=dt, nsteps=40, savesteps=[40], scheme=scheme, precision="f32")
    assert per_solve < (smm.Context.default().launches - launches) - 40      # the cluster path really ran: 1 launch, not >= 8 per step
    r = smm.pde_solve_raw(p, [q], u0, z0, y0, dt=dt, nsteps=n, savesteps=[0, 3, 6], scheme=scheme, precision="f64")
    assert rel(a["us"], r["us"]) < TOL32 and rel(a["us"], c["us"]) < 2e-6
    assert np.abs(a["ys"] - r["ys"]).max() < 1e-6 and np.abs(a["zs"] - r["zs"]).max() < 1e-6
    if grid == "coarse":
        saved = P.rk_solve(u0, z0, y0, n, dt, oracle_p(p), dict(q), scheme=scheme, save_every=3)
        for m, (k, ur, zr, yr) in enumerate(saved):
            assert rel(a["us"][..., m, 0], ur) < TOL32, (scheme, k)
            assert np.abs(a["ys"][..., m, 0] - yr).max() < 1e-6 and np.abs(a["zs"][..., m, 0] - zr).max() < 1e-6


def test_cluster_steps_batch_resume_and_controlled_influencer(monkeypatch):
    """(a) three instances = three clusters, each equal to its solo run bit for bit; (b) 6 steps in one launch = 2 + 4 steps in
    two launches (the state leaves and re-enters shared memory); (c) L = 5 with a controlled, zero-mass influencer
    (controlstrategies.jl:16-30) against the oracle."""
    from oracle import control_oracle as CO
    _cluster(monkeypatch, True)
    p = smm.PDEconstructcoarse()
    qs = [smm.parameters(), smm.parameters(a=1.5, eta=10.0), smm.parameters(c=3.0, sigma=0.4)]
    inits = [smm.constructinitial("4inf", (p, q)) for q in qs]
    ql = [qq for _, qq in inits]
    U = np.stack([uz.x[0] for uz, _ in inits], axis=-1); Z = np.stack([uz.x[1] for uz, _ in inits], axis=-1); Y = np.stack([uz.x[2] for uz, _ in inits], axis=-1)
    batch = smm.pde_solve_raw(p, ql, U, Z, Y, dt=1e-3, nsteps=6, savesteps=[6], precision="f32")
    split = smm.pde_solve_raw(p, ql, U, Z, Y, dt=1e-3, nsteps=6, savesteps=[2, 6], precision="f32")
    np.testing.assert_array_equal(split["us"][..., 1, :], batch["us"][..., 0, :])
    np.testing.assert_array_equal(split["ys"][..., 1, :], batch["ys"][..., 0, :])
    for b in range(3):
        solo = smm.pde_solve_raw(p, [ql[b]], U[..., b:b + 1], Z[..., b:b + 1], Y[..., b:b + 1], dt=1e-3, nsteps=6, savesteps=[6], precision="f32")
        np.testing.assert_array_equal(solo["us"][..., 0, 0], batch["us"][..., 0, b])
    # (c)
    q = smm.parameters(eta=1.0)
    po, qo = oracle_p(p), A_default_q(eta=1.0)
    uzy0, sl, q2, ci, cm, slm = CO.prep(0.01, po, qo, start="zero", savedt=0.01, dt=1e-3, countercontrol="no")
    u, z, y = uzy0
    qq = q.merge(L=5, controlled_inf=ci, controlled_med=cm)
    out = smm.pde_solve_raw(p, [qq], u, z, y, dt=1e-3, nsteps=5, savesteps=[5], precision="f32")
    ur, zr, yr = P.rk_solve(u, z, y, 5, 1e-3, po, dict(q2), ci, cm)
    assert rel(out["us"][..., 0, 0], ur) < TOL32 and np.abs(out["ys"][..., 0, 0] - yr).max() < 1e-6
    # (d) the same controlled L = 5 instance on the default 89 x 89 grid (the largest shape the cluster form holds), against
    # the Float64 kernel chain
    pd = smm.PDEconstruct()
    uzy0, sl, q2, ci, cm, slm = CO.prep(0.01, oracle_p(pd), qo, start="zero", savedt=0.01, dt=1e-3, countercontrol="no")
    u, z, y = uzy0
    l0 = smm.Context.default().launches
    out = smm.pde_solve_raw(pd, [qq], u, z, y, dt=1e-3, nsteps=8, savesteps=[8], precision="f32")
    assert smm.Context.default().launches - l0 < 8 * 8                               # one launch for the 8 steps, not a chain
    ref = smm.pde_solve_raw(pd, [qq], u, z, y, dt=1e-3, nsteps=8, savesteps=[8], precision="f64")
    assert rel(out["us"], ref["us"]) < TOL32 and np.abs(out["ys"] - ref["ys"]).max() < 1e-6
