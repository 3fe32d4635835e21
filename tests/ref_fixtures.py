"""Reference fixtures written by baseline/julia_fixtures.jl (the reference's own Julia code driven with logged random draws).

They cannot be produced in this image (no Julia); when tests/golden/ref/ exists the CPU suite checks the oracle against
them and the GPU suite checks the CUDA path, otherwise those tests skip with "julia absent".  The logs are replayed in the
reference's consumption order:
    uniform log:  rand(n,2) [2n, column-major] | rand(n) | per step, per agent j = 1..n: u1_j, then u2_j only if the agent
                  enters the switch branch (abm.jl:122-126)
    normal  log:  per step: randn(n,2) [2n, column-major] | L x randn(2,1) | M x randn(2,1)
`emulate_reference_log` builds logs of exactly that layout from an oracle run: it is how the replay itself is tested
without Julia (tests/test_golden_cpu.py::test_reference_log_replay_roundtrip)."""
import json
import os

import numpy as np

from oracle import abm_oracle as A

REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref")
SKIP_REASON = ("julia absent: tests/golden/ref/ has not been generated (run baseline/julia_fixtures.jl once with Julia 1.8.5 "
               "and the reference checkout); the oracle stays pinned to closed-form KATs only -- parity unpinned")


def available():
    return os.path.exists(os.path.join(REF_DIR, "manifest.json"))


def manifest():
    return json.load(open(os.path.join(REF_DIR, "manifest.json")))


def load(name):
    return np.load(os.path.join(REF_DIR, name + ".npy"))


class Log:
    def __init__(self, values):
        self.v, self.i = np.asarray(values, float), 0

    def take(self, k):
        out = self.v[self.i:self.i + k]
        assert out.size == k, "log exhausted: the replay consumed more draws than the reference"
        self.i += k
        return out

    def one(self):
        return float(self.take(1)[0])


def replay_abm(uni, nrm, NT, p, q):
    """One ABMsolve call (abm.jl:154-217) replayed from the logs with the oracle's arithmetic.  Returns the injected-draw
    form every backend accepts (x0, state, follow0, inf0, media0, xi[NT,n,2], u1[NT,n], u2[NT,n], xi_inf, xi_med) plus the
    oracle's run.  Consumes the logs in place (ensembles replay several calls from one log)."""
    n, L, M, dt = q["n"], q["L"], q["M"], p["dt"]
    dom = p["domain"]
    r = uni.take(2 * n)
    x0 = np.stack([r[:n], r[n:]], axis=1) * (dom[0, 1] - dom[0, 0]) + dom[0, 0]          # abm.jl:16
    state = np.where(uni.take(n) > 0.5, 1, -1)                                            # abm.jl:39
    follow0, inf0, media0 = A.abm_init(x0)
    xi = np.zeros((NT, n, 2)); u1 = np.zeros((NT, n)); u2 = np.full((NT, n), 0.5)
    xi_inf = np.zeros((NT, L, 2)); xi_med = np.zeros((NT, M, 2))
    x, inf, media, follow = x0.copy(), inf0.copy(), media0.copy(), follow0.copy()
    xs, fs, infs, meds = [x.copy()], [follow.copy()], [inf.copy()], [media.copy()]
    for k in range(NT):
        z = nrm.take(2 * n)
        xi[k] = np.stack([z[:n], z[n:]], axis=1)                                          # abm.jl:173
        xi_inf[k] = nrm.take(2 * L).reshape(L, 2)                                         # abm.jl:181/183, one randn(2,1) per influencer
        xi_med[k] = nrm.take(2 * M).reshape(M, 2)                                         # abm.jl:194/196
        # the switch rates depend on quantities known before the agent loop: one oracle evaluation gives every threshold
        xn, inf_n, media_n, _, _ = A.abm_step(x, inf, media, follow, state, xi[k], np.ones(n), np.zeros(n), p, q,
                                              None, xi_inf[k], xi_med[k])
        _, lam, _ = A.changeinfluencer(state, x, follow, inf_n, np.ones(n), np.zeros(n), q["eta"], dt, L)
        with np.errstate(invalid="ignore"):
            thr = 1.0 - np.exp(-lam * dt)
        for j in range(n):                                                                # abm.jl:121-134
            u1[k, j] = uni.one()
            if u1[k, j] < thr[j]:
                u2[k, j] = uni.one()
        x, inf, media, follow, _ = A.abm_step(x, inf, media, follow, state, xi[k], u1[k], u2[k], p, q, None, xi_inf[k], xi_med[k])
        xs.append(x.copy()); fs.append(follow.copy()); infs.append(inf.copy()); meds.append(media.copy())
    return dict(x0=x0, state=state, follow0=follow0, inf0=inf0, media0=media0, xi=xi, u1=u1, u2=u2, xi_inf=xi_inf, xi_med=xi_med,
                xs=np.array(xs), follows=np.array(fs), infs=np.array(infs), meds=np.array(meds))


def emulate_reference_log(seed, NT, p, q):
    """What baseline/julia_fixtures.jl would log for one ABMsolve call, produced from an oracle run with NumPy draws."""
    rng = np.random.default_rng(seed)
    n, L, M = q["n"], q["L"], q["M"]
    dom = p["domain"]
    uni, nrm = [], []
    r = rng.random((n, 2)); uni += list(r[:, 0]) + list(r[:, 1])
    x = r * (dom[0, 1] - dom[0, 0]) + dom[0, 0]
    s = rng.random(n); uni += list(s)
    state = np.where(s > 0.5, 1, -1)
    follow, inf, media = A.abm_init(x)
    xs, fs = [x.copy()], [follow.copy()]
    for k in range(NT):
        z = rng.standard_normal((n, 2)); nrm += list(z[:, 0]) + list(z[:, 1])
        zi = rng.standard_normal((L, 2)); nrm += list(zi.reshape(-1))
        zm = rng.standard_normal((M, 2)); nrm += list(zm.reshape(-1))
        _, inf_n, _, _, _ = A.abm_step(x, inf, media, follow, state, z, np.ones(n), np.zeros(n), p, q, None, zi, zm)
        _, lam, _ = A.changeinfluencer(state, x, follow, inf_n, np.ones(n), np.zeros(n), q["eta"], p["dt"], L)
        with np.errstate(invalid="ignore"):
            thr = 1.0 - np.exp(-lam * p["dt"])
        u1 = np.zeros(n); u2 = np.full(n, 0.5)
        for j in range(n):
            u1[j] = rng.random(); uni.append(u1[j])
            if u1[j] < thr[j]:
                u2[j] = rng.random(); uni.append(u2[j])
        x, inf, media, follow, _ = A.abm_step(x, inf, media, follow, state, z, u1, u2, p, q, None, zi, zm)
        xs.append(x.copy()); fs.append(follow.copy())
    return np.array(uni), np.array(nrm), np.array(xs), np.array(fs)
