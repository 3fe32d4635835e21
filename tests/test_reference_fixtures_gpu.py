"""CUDA path against the REFERENCE's own outputs (tests/golden/ref/, written by baseline/julia_fixtures.jl from the
reference's Julia code with logged draws).  Skipped with "julia absent" until somebody has run that script once; the log
replay itself is exercised on every run with an emulated log."""
import numpy as np
import pytest

import ref_fixtures as RF
import socialmediamodel_b200 as smm
from oracle import abm_oracle as A

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not RF.available(), reason=RF.SKIP_REASON)


def run_cuda(r, NT, precision):
    p, q = smm.ABMconstruct(), smm.parameters(n=r["x0"].shape[0])
    init = dict(x0=r["x0"][..., None], state=r["state"][:, None], follow0=r["follow0"][:, None], inf0=r["inf0"][..., None],
                media0=r["media0"][..., None])
    inj = dict(xi=np.transpose(r["xi"], (1, 2, 0))[..., None], u1=r["u1"].T[..., None], u2=r["u2"].T[..., None])
    out = smm.abm_solve_raw(NT, p=p, q=q, precision=precision, init_state=init, injected=inj)
    return np.transpose(out["xs"][:, :, :, 0], (2, 0, 1)), out["follows"][:, :, 0].T, np.transpose(out["infs"][:, :, :, 0], (2, 0, 1))


def test_emulated_reference_log_through_cuda():
    p, q = A.default_p(), A.default_q(n=64)
    uni, nrm, xs_ref, fs_ref = RF.emulate_reference_log(8, 30, p, q)
    r = RF.replay_abm(RF.Log(uni), RF.Log(nrm), 30, p, q)
    xs, fol, _ = run_cuda(r, 30, "f64")
    np.testing.assert_array_equal(fol, fs_ref)
    assert np.abs(xs - xs_ref).max() < 1e-11


@needs_ref
@pytest.mark.parametrize("tag", ["abm_a", "abm_b", "abm_c"])
def test_cuda_against_reference_abmsolve(tag):
    from parity_util import check_history
    m = RF.manifest()[tag]
    p, q = A.default_p(), A.default_q()
    r = RF.replay_abm(RF.Log(RF.load(tag + "_uniform_log")), RF.Log(RF.load(tag + "_normal_log")), m["NT"], p, q)
    ref_xs = np.transpose(RF.load(tag + "_xs"), (2, 0, 1))
    ref_fol = (RF.load(tag + "_stateinfs").T - 1).astype(np.int64)
    o = dict(follows=ref_fol, xs=ref_xs, margins=A.abm_solve(m["NT"], r["x0"], r["state"], r["follow0"], r["inf0"], r["media0"],
                                                            r["xi"], r["u1"], r["u2"], p, q)["margins"])
    for precision, tol in (("f64", 1e-11), ("f32", 1e-5)):
        xs, fol, _ = run_cuda(r, m["NT"], precision)
        check_history(fol, o, exact=precision == "f64", label=f"reference {tag} {precision}", xs=xs, rtol=tol)


@needs_ref
@pytest.mark.parametrize("tag", ["pde_coarse", "pde_default"])
def test_cuda_against_reference_pde_rhs(tag):
    if tag not in RF.manifest():
        pytest.skip("the reference's PDE part did not load when the fixtures were generated")
    p = smm.PDEconstructcoarse() if tag == "pde_coarse" else smm.PDEconstruct()
    q = smm.parameters()
    _, q2 = smm.constructinitial("4inf", (p, q))
    for sub in ("init", "pert"):
        u, z, y = (RF.load(f"{tag}_{sub}_{nm}") for nm in ("u", "z", "y"))
        ref = RF.load(f"{tag}_{sub}_du")
        for precision, tol in (("f64", 1e-10), ("f32", 1e-5)):
            du, dz, dy = smm.pde_rhs(u, z, y, p=p, q=q2, precision=precision)
            assert np.abs(du - ref).max() <= tol * np.abs(ref).max()
            assert np.abs(dy - RF.load(f"{tag}_{sub}_dy")).max() < max(tol, 1e-9)
