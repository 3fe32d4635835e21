"""Counted, bounded follower-history parity for the FP32 mode (tests and __graft_entry__.smoke()).

The FP32 kernels evaluate the switch decision of abm.jl:122-133 in Float64 from FP32 positions, so a decision can only
differ from the Float64 oracle where the oracle's own decision variable sits within FP32 position error of its
threshold.  Instead of waving such a run through, the check below
  * finds the first step at which the histories differ,
  * requires that at most MAX_FLIPS agents differ there and that EACH of them is a near-tie in the oracle
    (margin < MARGIN at that step),
  * still compares every position / influencer / medium up to and including that step (the state at step t does not
    depend on the decisions taken at step t), and
  * counts how often this happened (LEDGER); test_fp32_divergence_budget bounds the count over the whole module.
Float64 mode must be bit-exact; no allowance."""
import numpy as np

MARGIN = 1e-5      # north_star tolerance: FP32 state vs the reference's Float64
MAX_FLIPS = 2      # near-ties are isolated events; more than this at one step means a wrong kernel
LEDGER = dict(cases=0, diverged=0, flips=0, log=[])


def relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def check_history(fol, o, *, exact, label, xs=None, infs=None, meds=None, rtol=MARGIN):
    """fol[NT+1,n] against the oracle run `o` (abm_oracle.abm_solve).  Returns the number of leading history
    entries that were compared (NT+1 when the histories are identical)."""
    ref = o["follows"]
    assert fol.shape == ref.shape, (fol.shape, ref.shape)
    LEDGER["cases"] += 1
    diff = fol != ref
    upto = ref.shape[0]
    if diff.any():
        assert not exact, f"{label}: Float64 follower history differs first at step {np.nonzero(diff.any(axis=1))[0][0]}"
        t = int(np.nonzero(diff.any(axis=1))[0][0])
        agents = np.nonzero(diff[t])[0]
        assert t >= 1, f"{label}: initial followers differ"
        margins = o["margins"][t - 1, agents]
        assert agents.size <= MAX_FLIPS, f"{label}: {agents.size} agents differ at step {t}"
        assert np.all(margins < MARGIN), f"{label}: decision flipped at step {t} with oracle margins {margins}"
        LEDGER["diverged"] += 1
        LEDGER["flips"] += int(agents.size)
        LEDGER["log"].append((label, t, agents.tolist(), margins.tolist()))
        upto = t + 1
    for got, key in ((xs, "xs"), (infs, "infs"), (meds, "meds")):
        if got is not None:
            e = relerr(got[:upto], o[key][:upto])
            assert e < rtol, f"{label}: {key} rel. error {e:.3e} over the first {upto} history entries"
    return upto
