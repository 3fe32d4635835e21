"""Philox4x32-10 counter-based generator (Salmon et al., SC'11), NumPy restatement.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference draws from Julia's
task-local Xoshiro256++ (abm.jl:155 ``Random.seed!``); that stream cannot be
reproduced on a GPU in parallel, so the build replaces it by a counter-based
stream (key = seed, counter = (agent, step, stream, realization)) and parity
is established by *injecting identical draws* into oracle and kernels.  This
file restates the integer part of that generator bit-exactly so the uniform
draws (u1,u2) can be cross-checked on the CPU; the Gaussian draws are produced
on the device in FP32 and dumped through ``smm_abm_draws`` for injection.
Known-answer vectors: Random123 kat_vectors (philox4x32-10), tested in
tests/test_oracle_kats.py.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32(ctr, key, rounds=10):
    """ctr: (..., 4) uint32, key: (..., 2) uint32 -> (..., 4) uint32."""
    c = [np.asarray(ctr[..., i], dtype=np.uint64) for i in range(4)]
    k0 = np.asarray(key[..., 0], dtype=np.uint64)
    k1 = np.asarray(key[..., 1], dtype=np.uint64)
    for _ in range(rounds):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c = [(hi1 ^ c[1] ^ k0) & MASK, lo1, (hi0 ^ c[3] ^ k1) & MASK, lo0]
        k0 = (k0 + np.uint64(W0)) & MASK
        k1 = (k1 + np.uint64(W1)) & MASK
    return np.stack([x.astype(np.uint32) for x in c], axis=-1)


STREAM_AGENT = 0   # per (agent, step): xi_x, xi_y, u1, u2
STREAM_ENTITY = 1  # per (influencer|medium, step): noise on influencers / media
STREAM_INIT = 2    # per agent: x0, y0, state


def raw_draws(seed, realization, step, agents, stream=STREAM_AGENT):
    """uint32[len(agents), 4] for counter (agent, step, stream, realization), key = seed."""
    agents = np.asarray(agents, dtype=np.uint32)
    ctr = np.stack([agents,
                    np.full_like(agents, step),
                    np.full_like(agents, stream),
                    np.full_like(agents, realization)], axis=-1)
    key = np.empty(agents.shape + (2,), dtype=np.uint32)
    key[..., 0] = np.uint32(seed & 0xFFFFFFFF)
    key[..., 1] = np.uint32((seed >> 32) & 0xFFFFFFFF)
    return philox4x32(ctr, key)


def uniform24(bits):
    """[0,1) with 24 random bits -- exactly representable in FP32 and FP64."""
    return (np.asarray(bits, dtype=np.uint32) >> np.uint32(8)).astype(np.float64) * 2.0 ** -24
