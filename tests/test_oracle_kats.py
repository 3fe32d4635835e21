"""Known-answer tests that pin the oracle to the mathematics of the reference source
(invariants 1-6) -- the reference ships no golden vectors."""
import numpy as np
import pytest

from oracle import abm_oracle as A
from oracle import ensemble_oracle as E
from oracle import pde_oracle as P
from oracle import philox


def test_philox_random123_kat():
    z4 = np.zeros((1, 4), np.uint32); z2 = np.zeros((1, 2), np.uint32)
    assert [int(v) for v in philox.philox4x32(z4, z2)[0]] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f4 = np.full((1, 4), 0xffffffff, np.uint32); f2 = np.full((1, 2), 0xffffffff, np.uint32)
    assert [int(v) for v in philox.philox4x32(f4, f2)[0]] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    c = np.array([[0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344]], np.uint32)
    k = np.array([[0xa4093822, 0x299f31d0]], np.uint32)
    assert [int(v) for v in philox.philox4x32(c, k)[0]] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_attraction_two_agent_closed_form():
    # invariant 6: force_1 = w d/(1+w), w = exp(-|d|) (self term in the denominator, abm.jl:42,51-63)
    x = np.array([[0.3, -0.2], [1.1, 0.7]])
    d = x[1] - x[0]
    w = np.exp(-np.linalg.norm(d))
    f = A.attraction(x)
    np.testing.assert_allclose(f[0], w * d / (1 + w), rtol=1e-15)
    np.testing.assert_allclose(f[1], -w * d / (1 + w), rtol=1e-15)


def test_attraction_csr_equals_dense():
    rng = np.random.default_rng(1)
    n = 40
    x = rng.uniform(-2, 2, (n, 2))
    rowptr = np.arange(0, n * n + 1, n)
    colidx = np.tile(np.arange(n), n)
    np.testing.assert_allclose(A.attraction(x, (rowptr, colidx)), A.attraction(x), rtol=1e-13, atol=1e-15)
    # empty neighbourhood -> zero force (abm.jl:52-53)
    rp = np.zeros(n + 1, dtype=np.int64)
    assert np.all(A.attraction(x, (rp, np.zeros(0, dtype=np.int64))) == 0)


def _random_run(seed, n=60, NT=20, L=4):
    rng = np.random.default_rng(seed)
    p = A.default_p(); q = A.default_q(n=n)
    x0 = rng.uniform(-2.2, 2.2, (n, 2))
    state = np.where(rng.random(n) > 0.5, 1, -1)
    follow0, inf0, media0 = A.abm_init(x0)
    return p, q, dict(x0=x0, state=state, follow0=follow0, inf0=inf0, media0=media0,
                      xi=rng.standard_normal((NT, n, 2)), u1=rng.random((NT, n)), u2=rng.random((NT, n)))


def test_abm_frozen_dynamics_and_no_switch():
    # invariant 5: a=b=c=sigma=0 -> positions constant; eta=0 -> no switches
    p, q, r = _random_run(0)
    q.update(a=0.0, b=0.0, c=0.0, sigma=0.0, eta=0.0)
    out = A.abm_solve(20, r["x0"], r["state"], r["follow0"], r["inf0"], r["media0"], r["xi"], r["u1"], r["u2"], p, q)
    assert np.all(out["xs"] == r["x0"][None])
    assert np.all(out["follows"] == r["follow0"][None])


def test_abm_one_influencer_each_and_nan_lock():
    # invariant 4 + Q5: an influencer without followers gives fraction=NaN -> nobody switches again
    p, q, r = _random_run(3)
    out = A.abm_solve(20, r["x0"], r["state"], r["follow0"], r["inf0"], r["media0"], r["xi"], r["u1"], r["u2"], p, q)
    assert out["follows"].min() >= 0 and out["follows"].max() <= 3
    assert (out["follows"] != out["follows"][0]).any()          # switching happens at eta=15
    f0 = r["follow0"].copy(); f0[f0 == 2] = 0                    # influencer 2 loses everyone
    inf0 = r["inf0"].copy()
    out2 = A.abm_solve(5, r["x0"], r["state"], f0, inf0, r["media0"], r["xi"], r["u1"], r["u2"], p, q)
    assert np.all(out2["follows"] == f0[None])
    np.testing.assert_array_equal(out2["infs"][:, 2], np.broadcast_to(inf0[2], (6, 2)))  # noise-only, sigmahat=0


def test_abm_reflection_single():
    x = np.array([[2.5, -2.3], [0.0, 2.2]])
    A.boundaryconditions(x, A.default_p()["domain"])
    np.testing.assert_allclose(x, [[1.9, -2.1], [0.0, 2.2]], rtol=1e-15)


def test_abm_savetimes():
    assert list(E.savetimes_abm(200, 4) + 1) == [1, 67, 134, 200]
    assert list(E.savetimes_abm(200, 5) + 1) == [1, 51, 100, 150, 200]
    assert list(E.savetimes_abm(10, 2) + 1) == [1, 10]


def test_kde_separable_equals_naive_and_mass():
    p, q, r = _random_run(5, n=30)
    us = E.kde_snapshot(r["x0"], r["follow0"], r["state"], 4, p["X"], p["Y"], 30)
    un = E.kde_snapshot_naive(r["x0"], r["follow0"], r["state"], 4, p["X"], p["Y"], 30)
    np.testing.assert_allclose(us, un, rtol=1e-12, atol=1e-14)
    inner = np.all(np.abs(r["x0"]) < 1.7, axis=1)                # Gaussians fully inside the grid
    if inner.all():
        assert abs(us.sum() * p["dV"] - 1) < 1e-6


@pytest.fixture(scope="module")
def coarse():
    p = P.pde_construct_coarse(materialise=True)
    q = A.default_q()
    return p, q, P.uniforminit(p, q)


def test_grid_sizes_and_uniforminit_indices():
    assert P.pde_construct_coarse()["N_x"] == 47
    assert P.pde_construct_meso()["N_x"] == 61
    assert P.pde_construct()["N_x"] == 89
    assert P.pde_construct(dx=0.025)["N_x"] == 177
    assert P.pde_construct(dx=0.0125)["N_x"] == 353
    for dxv, (s, m, e) in {0.05: (5, 44, 85), 0.025: (9, 88, 169)}.items():
        p = P.pde_construct(dx=dxv)
        u0 = P.uniforminit(p, A.default_q())[0]
        nzx = np.nonzero(u0[:, :, 0, 3].sum(axis=1))[0] + 1
        assert (nzx[0], nzx[-1]) == (s, m)
        nzx1 = np.nonzero(u0[:, :, 0, 0].sum(axis=1))[0] + 1
        assert (nzx1[0], nzx1[-1]) == (m + 1, e)
        assert abs(u0.sum() * p["dV"] - 1) < 1e-13


def test_K_W_structure(coarse):
    p, _, _ = coarse
    K, W = p["K_matrix"], p["W_matrix"]
    assert np.array_equal(W, W.T) and np.all(np.diag(W) == 1)
    assert np.array_equal(K[:, :, 0], -K[:, :, 0].T) and np.all(np.diag(K[:, :, 1]) == 0)


def test_toeplitz_force_equals_dense_and_direct(coarse):
    p, q, (u0, z0, y0, ci, cm) = coarse
    rng = np.random.default_rng(0)
    rho = u0.sum(axis=(2, 3)) + 0.01 * rng.random((47, 47))
    Fd = P.agent_force_dense(rho, p["K_matrix"], p["W_matrix"], p["dV"])
    Ft = P.agent_force_toeplitz(rho, p["dx"], p["dy"])
    np.testing.assert_allclose(Ft, Fd, rtol=1e-11, atol=1e-13)
    small = rho[:12, :12]
    np.testing.assert_allclose(P.agent_force_direct(small, 0.1, 0.1), P.agent_force_toeplitz(small, 0.1, 0.1),
                               rtol=1e-11, atol=1e-13)


def test_rhs_invariants_at_t0(coarse):
    p, q, (u0, z0, y0, ci, cm) = coarse
    du, dz, dy, parts = P.rhs(u0, z0, y0, p, q, ci, cm, force="dense", parts=True)
    # invariant 3: boundary rows are zero at t=0 -> sum du dV = 0 to round-off
    assert abs(du.sum() * p["dV"]) < 1e-12
    # dz_i = -z_i/frictionM (medium densities have zero mean)
    np.testing.assert_allclose(dz, -z0 / q["frictionM"], atol=1e-12)
    assert np.abs(dy).max() <= 5e-3 + 1e-12
    # invariant 1: sum_j reac = 0 -- check through the rates: sum over j of du equals the
    # reaction-free du summed over j
    q0 = dict(q, eta=0.0)
    du0, _, _ = P.rhs(u0, z0, y0, p, q0, ci, cm, force="dense")
    np.testing.assert_allclose(du.sum(axis=3), du0.sum(axis=3), atol=1e-10)
    # survey probe: max|du| = 3.84 at 47^2, max u = 0.0297
    assert abs(np.abs(du).max() - 3.84) < 0.01 and abs(u0.max() - 0.0297) < 1e-4


def test_rhs_controlled_and_zero_mass(coarse):
    p, q, (u0, z0, y0, ci, cm) = coarse
    q = dict(q, L=5)
    u = np.concatenate([u0, np.zeros_like(u0[..., :1])], axis=3)          # zero-mass 5th influencer (Q14)
    y = np.concatenate([y0, np.array([[0.0], [0.0]])], axis=1)
    ci5 = np.array([0, 0, 0, 0, 1.0])
    du, dz, dy = P.rhs(u, z0, y, p, q, ci5, np.array([0.0, 1.0]), force="dense")
    assert np.isfinite(du).all() and np.isfinite(dy).all()
    t = np.array([1.5, 1.5])
    np.testing.assert_allclose(dy[:, 4], q["controlspeed1"] * t / np.linalg.norm(t), rtol=1e-14)
    np.testing.assert_allclose(dz[:, 1], q["controlspeed1"] * (t - z0[:, 1]) / np.linalg.norm(t - z0[:, 1]), rtol=1e-14)
    assert (du[..., 4] > 0).any()                                          # mass flows to the new influencer


def test_rk4_converged_against_adaptive(coarse):
    p, q, (u0, z0, y0, ci, cm) = coarse
    T = 0.02
    u1, z1, y1 = P.rk_solve(u0, z0, y0, 20, T / 20, p, q, ci, cm)
    u2, z2, y2 = P.rk_solve(u0, z0, y0, 40, T / 40, p, q, ci, cm)
    ua, za, ya = P.adaptive_solve(u0, z0, y0, T, p, q, ci, cm)
    s = np.abs(ua).max()
    assert np.abs(u1 - u2).max() / s < 1e-5
    assert np.abs(u2 - ua).max() / s < 1e-6


def test_bs3_adaptive_controller(coarse):
    """The embedded 3(2) pair and its controller (what csrc/pde.cu runs on the device): lands on the save
    times, error follows the tolerance against the SciPy DOP853 solve, forced dtmin steps terminate (a step
    shortened below dtmin to reach the save time is accepted too), an unforced failure at dtmin raises."""
    p, q, (u0, z0, y0, ci, cm) = coarse
    T = 0.04
    ua, za, ya = P.adaptive_solve(u0, z0, y0, T, p, q, ci, cm)
    s = np.abs(ua).max()
    errs, work = [], []
    for rtol in (1e-2, 1e-4):
        states, log = P.bs3_adaptive(u0, z0, y0, [0.0, T / 2, T], p, q, ci, cm, abstol=1e-3 * rtol, reltol=rtol)
        np.testing.assert_array_equal(states[0][0], u0)                    # save_start
        errs.append(np.abs(states[-1][0] - ua).max() / s)
        work.append(len(log))
        ts = [t for t, dt, n, acc in log if acc]
        assert any(abs(t + dt - T / 2) < 1e-15 for t, dt, n, acc in log if acc)     # clipped onto the save time
        assert all(n <= 1.0 for t, dt, n, acc in log if acc) and all(n > 1.0 for t, dt, n, acc in log if not acc)
    assert errs[1] < errs[0] < 5e-2 and errs[1] < 1e-3 and work[1] > work[0]
    states, log = P.bs3_adaptive(u0, z0, y0, [T], p, q, ci, cm, dtmin=4e-3, dt0=4e-3, force_dtmin=True)
    assert len(log) == 10 and all(acc for *_, acc in log)
    with pytest.raises(RuntimeError, match="dtmin"):
        P.bs3_adaptive(u0, z0, y0, [T], p, q, ci, cm, dtmin=4e-3, dt0=4e-3, force_dtmin=False)
