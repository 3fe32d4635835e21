"""GPU parity of the adaptive PDE integrator (smm_pde_batch_advance / smm_pde_solve_adaptive) against the
oracle restatement of the same embedded 3(2) pair and controller (oracle/pde_oracle.bs3_adaptive), through
the C ABI.  Float64 mode must reproduce the oracle's accepted / rejected step sequence and states; FP32
mode must agree within the requested tolerance (its accept / reject decisions may differ where the error
norm is within rounding of 1)."""
import numpy as np
import pytest

import socialmediamodel_b200 as smm
from oracle import pde_oracle as P

pytestmark = pytest.mark.gpu


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


SAVE = np.array([0.0, 0.03, 0.05, 0.1])


@pytest.fixture(scope="module")
def oracle_run(coarse):
    p, q, uzy0 = coarse
    return P.bs3_adaptive(*uzy0.x, SAVE, oracle_p(p), dict(q))


def test_adaptive_f64_reproduces_oracle_step_sequence(coarse, oracle_run):
    p, q, uzy0 = coarse
    states, log = oracle_run
    out = smm.pde_solve_adaptive_raw(p, [q], *uzy0.x, savetimes=SAVE, precision="f64")
    assert int(out["naccept"][0]) == sum(1 for l in log if l[3])
    assert int(out["nreject"][0]) == sum(1 for l in log if not l[3])
    assert int(out["nreject"][0]) >= 1          # the first trial step 0.1 dx^2/D is too long for the tolerance
    np.testing.assert_array_equal(out["us"][..., 0, 0], uzy0.x[0])   # save_start = true: the initial state itself
    for m, (u, z, y) in enumerate(states):
        assert rel(out["us"][..., m, 0], u) < 1e-10
        assert np.abs(out["zs"][..., m, 0] - z).max() < 1e-10
        assert np.abs(out["ys"][..., m, 0] - y).max() < 1e-10


def test_adaptive_f32_within_tolerance(coarse, oracle_run):
    p, q, uzy0 = coarse
    states, log = oracle_run
    out = smm.pde_solve_adaptive_raw(p, [q], *uzy0.x, savetimes=SAVE, precision="f32")
    n_or = len(log)
    assert abs(int(out["naccept"][0] + out["nreject"][0]) - n_or) <= max(3, n_or // 5)
    for m, (u, z, y) in enumerate(states):
        assert rel(out["us"][..., m, 0], u) < 3e-3          # both are reltol = 1e-3 solutions
        assert np.abs(out["ys"][..., m, 0] - y).max() < 3e-3


def test_adaptive_solution_converges_to_fine_fixed_step(coarse):
    """Tightening the tolerance brings the adaptive solution to the converged fixed-step RK4 one."""
    p, q, uzy0 = coarse
    ref = smm.pde_solve_raw(p, [q], *uzy0.x, dt=2.5e-4, nsteps=400, savesteps=[400], precision="f64")["us"][..., 0, 0]
    errs, work = [], []
    for rtol in (1e-2, 1e-3, 1e-5):
        out = smm.pde_solve_adaptive_raw(p, [q], *uzy0.x, savetimes=[0.1], abstol=rtol * 1e-3, reltol=rtol, precision="f64")
        errs.append(rel(out["us"][..., 0, 0], ref))
        work.append(int(out["naccept"][0] + out["nreject"][0]))
    assert errs[0] > errs[1] > errs[2] and errs[2] < 1e-4, errs   # error follows the tolerance (about 3 x reltol)
    assert errs[1] < 1e-2 and errs[0] < 1e-1
    assert work[0] <= work[1] <= work[2], work


def test_adaptive_against_independent_integrators(coarse):
    """Not circular: the device integrator against two integrators it shares nothing with -- SciPy's RK23 (the standard
    4-stage FSAL Bogacki-Shampine pair with SciPy's controller) at the same tolerances, and DOP853 at a tight tolerance
    as the truth -- all on the Float64 oracle right-hand side.  Same tolerance semantics as the reference's
    solve(..., abstol, reltol) calls (pde.jl:291, controlstrategies.jl:10,96): the error of our solution is of the order of
    the requested tolerance and no worse than a few times RK23's, at a comparable number of right-hand-side evaluations."""
    p, q, uzy0 = coarse
    po, qo = oracle_p(p), dict(q)
    T = 0.03
    truth = P.adaptive_solve(*uzy0.x, T, po, qo, rtol=1e-7, atol=1e-10)[0]
    st = {}
    rk23 = P.adaptive_solve(*uzy0.x, T, po, qo, rtol=1e-3, atol=1e-6, method="RK23", stats=st)[0]
    out = smm.pde_solve_adaptive_raw(p, [q], *uzy0.x, savetimes=[T], abstol=1e-6, reltol=1e-3, precision="f64")
    ours = out["us"][..., 0, 0]
    e_ours, e_rk23 = rel(ours, truth), rel(rk23, truth)
    assert e_ours < 5e-3 and e_rk23 < 1e-2, (e_ours, e_rk23)   # measured: 2.3e-4 (72 evaluations) and 4.1e-3 (29)
    assert e_ours < 5 * max(e_rk23, 2e-4), (e_ours, e_rk23)
    evals = 3 * int(out["naccept"][0] + out["nreject"][0])
    assert evals <= 2 * st["nfev"] + 30, (evals, st)


def test_adaptive_per_instance_step_control(coarse):
    """Instances of one batch choose their own steps: each reproduces its solo run, and the step counts
    differ between the instances."""
    p, q, uzy0 = coarse
    qs = []
    for sigma in (0.5, 0.7, 1.0):     # (sigma = 0.3 is under-resolved on this grid: negative densities, step collapse)
        qs.append(q.merge(sigma=sigma))
    u0, z0, y0 = uzy0.x
    U = np.stack([u0] * 3, axis=-1); Z = np.stack([z0] * 3, axis=-1); Y = np.stack([y0] * 3, axis=-1)
    save = [0.02, 0.06]
    for precision, tol in (("f64", 1e-10), ("f32", 1e-4)):
        out = smm.pde_solve_adaptive_raw(p, qs, U, Z, Y, savetimes=save, precision=precision, max_attempts=5000)
        assert len(set(out["naccept"].tolist())) > 1          # different step sequences inside one batch
        for b, qq in enumerate(qs):
            solo = smm.pde_solve_adaptive_raw(p, [qq], u0, z0, y0, savetimes=save, precision=precision, max_attempts=5000)
            assert rel(out["us"][..., b], solo["us"][..., 0]) < tol
            assert np.abs(out["ys"][..., b] - solo["ys"][..., 0]).max() < tol
            slack = 0 if precision == "f64" else 2
            assert abs(int(out["naccept"][b]) - int(solo["naccept"][0])) <= slack
            assert abs(int(out["nreject"][b]) - int(solo["nreject"][0])) <= slack


def test_adaptive_dtmin_force_and_failure(coarse):
    """dtmin / force_dtmin as in controlstrategies.jl:10: with force_dtmin the steps of size dtmin are taken
    whatever their error; without it the solve reports the failure instead of continuing."""
    p, q, uzy0 = coarse
    dtmin = 4e-3            # above what the tolerance allows (about 2e-3 on this grid), below the stability limit
    out = smm.pde_solve_adaptive_raw(p, [q], *uzy0.x, savetimes=[0.04], dtmin=dtmin, dt0=dtmin, force_dtmin=True, precision="f64", max_attempts=2000)
    assert int(out["naccept"][0]) == 10 and int(out["nreject"][0]) == 0
    states, _ = P.bs3_adaptive(*uzy0.x, [0.04], oracle_p(p), dict(q), dtmin=dtmin, dt0=dtmin, force_dtmin=True)
    assert rel(out["us"][..., 0, 0], states[0][0]) < 1e-10
    with pytest.raises(smm.SmmError, match="dtmin"):
        smm.pde_solve_adaptive_raw(p, [q], *uzy0.x, savetimes=[0.04], dtmin=dtmin, dt0=dtmin, force_dtmin=False, precision="f64", max_attempts=2000)


def test_adaptive_batch_clock_and_mixing_with_fixed_steps(coarse):
    p, q, uzy0 = coarse
    b = smm.PDEBatch(p, [q], *uzy0.x, precision="f64")
    try:
        b.step(1e-3, 10)                                   # clock = 0.01
        na, nr, dn = b.advance(0.05)
        assert na[0] > 0 and dn[0] > 0
        u, z, y = b.read()
        na2, nr2, _ = b.advance(0.05)                      # already there: nothing to do
        assert na2[0] == 0 and nr2[0] == 0
        u2, _, _ = b.read()
        np.testing.assert_array_equal(u, u2)
        with pytest.raises(smm.SmmError, match="before the batch clock"):
            b.advance(0.01)
    finally:
        b.close()
    po = oracle_p(p)
    s = P.rk_solve(*uzy0.x, 10, 1e-3, po, dict(q))
    states, _ = P.bs3_adaptive(*s, [0.04], po, dict(q))
    assert rel(u[..., 0], states[0][0]) < 1e-10


def test_adaptive_controlled_influencer_f64(coarse):
    """L = 5 with a controlled influencer (flag 1) moving to its target: the kink at arrival is handled by the
    step control; Float64 parity with the oracle."""
    p, q, uzy0 = coarse
    u0, z0, y0 = uzy0.x
    L = 5
    u5 = np.zeros(u0.shape[:3] + (L,)); u5[..., :4] = u0
    y5 = np.zeros((2, L)); y5[:, :4] = y0; y5[:, 4] = (0.2, -0.1)
    qq = q.merge(L=L, controlled_inf=np.array([0, 0, 0, 0, 1.0]), controlled_med=np.zeros(2),
                 controltarget1=np.array([[0.25, -0.05]]), controlspeed1=2.0)
    save = [0.02, 0.06]
    out = smm.pde_solve_adaptive_raw(p, [qq], u5, z0, y5, savetimes=save, precision="f64")
    states, log = P.bs3_adaptive(u5, z0, y5, save, oracle_p(p), dict(qq), qq.controlled_inf, qq.controlled_med)
    assert int(out["naccept"][0]) == sum(1 for l in log if l[3])
    for m, (u, z, y) in enumerate(states):
        assert rel(out["us"][..., m, 0], u) < 1e-10
        assert np.abs(out["ys"][..., m, 0] - y).max() < 1e-10
    assert np.abs(out["ys"][:, 4, -1, 0] - qq.controltarget1.ravel()).max() < 1e-2     # arrived (chatters within speed * dt of the target)


def test_PDEsolve_adaptive_keyword(coarse):
    p, q, _ = coarse
    sol, (_, q2) = smm.PDEsolve(0.05, p=p, q=smm.parameters(), adaptive=True, saveat=np.arange(0, 0.051, 0.01), precision="f32")
    assert len(sol.t) == 6 and sol.naccept > 0
    fixed, _ = smm.PDEsolve(0.05, p=p, q=smm.parameters(), precision="f32")
    assert rel(sol(0.05).x[0], fixed(0.05).x[0]) < 5e-3
    assert np.abs(sol(0.025).x[2] - fixed(0.025).x[2]).max() < 5e-3


def test_PDEsolveensemble_adaptive_parameter_sweep(coarse):
    """ensemble.jl:58-85 on the adaptive integrator: a parameter sweep saved at LinRange(0, tmax, savepoints) agrees
    with the fixed-step sweep within the tolerance."""
    p, q, _ = coarse
    qs = [smm.parameters(a=a) for a in (0.5, 1.0, 2.0)]
    ua, za, ya, _ = smm.PDEsolveensemble(0.06, 3, savepoints=3, p=p, qs=qs, save=False, adaptive=True, precision="f32")
    uf, zf, yf, _ = smm.PDEsolveensemble(0.06, 3, savepoints=3, p=p, qs=qs, save=False, precision="f32")
    assert ua.shape == uf.shape == (p.N_x, p.N_y, 2, 4, 3, 3)
    np.testing.assert_array_equal(ua[..., 0, :], uf[..., 0, :])            # t = 0
    assert rel(ua, uf) < 5e-3 and np.abs(ya - yf).max() < 5e-3
    assert rel(ua[..., -1, 0], ua[..., -1, 2]) > 1e-2                      # the sweep parameter matters


def test_adaptive_and_control_against_golden_fixture(coarse):
    """tests/golden/pde_adaptive_control.npz (oracle-generated): controller log, states, control objective."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pde_adaptive_control.npz"))
    p, q, uzy0 = coarse
    out = smm.pde_solve_adaptive_raw(p, [q], *uzy0.x, savetimes=g["save"], precision="f64")
    assert int(out["naccept"][0]) == int(g["log"][:, 3].sum()) and int(out["nreject"][0]) == int((1 - g["log"][:, 3]).sum())
    assert rel(out["us"][..., 0], g["u_adapt"]) < 1e-10
    assert np.abs(out["ys"][..., 0] - g["y_adapt"]).max() < 1e-10
    r = smm.parameterscontrol(ntarg=2, start="zero", Tmax=0.05, tequil=0.02)
    fs, pen, sols, Ps = smm.PDEsolvefixedtargets(g["ctrl_targets"].reshape(-1, order="F"), p=p, q=smm.parametersstronginf(), r=r,
                                                 countercontrol="inf", stubborntarget=[1.5, 1.5], savedt=0.01, dt=1e-3,
                                                 precision="f64")
    assert abs(fs - float(g["ctrl_followersum"])) < 1e-10 * max(1.0, abs(fs)) + 1e-12
    assert abs(pen - float(g["ctrl_penalty"])) < 1e-12
