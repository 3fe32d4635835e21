#!/usr/bin/env python
"""Yardstick only (never linked into the library, not on any product path): the interaction-force convolution of one RK
stage (pde.jl:49-59: three zero-padded P x P circular convolutions of the density sum with the kernel images) done with
cuFFT through torch.fft, timed with CUDA events beside this library's hand-written FFT kernels on the same batch.

    python tools/cufft_yardstick.py            -> one JSON line per workload
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import socialmediamodel_b200 as smm  # noqa: E402


def cufft_conv_ms(N, P, B, reps=20):
    dev = torch.device("cuda")
    rho = torch.rand(B, N, N, device=dev, dtype=torch.float32)
    khat = torch.randn(3, P, P // 2 + 1, device=dev, dtype=torch.complex64)   # three kernel spectra (values do not matter for time)

    def conv():
        a = torch.fft.rfft2(rho, s=(P, P))                                     # zero-padded forward transform
        g = torch.fft.irfft2(a[:, None] * khat[None], s=(P, P))[..., :N, :N]  # three inverse transforms, kept part
        return g[:, 1] / g[:, 0], g[:, 2] / g[:, 0]                            # force ./ norms

    for _ in range(3):
        conv()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps):
        conv()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / reps


def ours_conv_ms(p, B, reps=3):
    """The three FFT kernels of one right-hand-side evaluation: device-timer class 3 of smm_pde_rhs (CUDA events around
    the convolution alone)."""
    ctx = smm.Context.default()
    q = smm.parameters()
    uzy0, q2 = smm.constructinitial("4inf", (p, q))
    u, z, y = uzy0.x
    U = np.repeat(u[..., None], B, axis=-1); Z = np.repeat(z[..., None], B, axis=-1); Y = np.repeat(y[..., None], B, axis=-1)
    ms = []
    for _ in range(reps):
        smm.pde_rhs(U, Z, Y, p=p, q=[q2] * B, precision="f32", ctx=ctx)
        ms.append(ctx.kernel_ms(3))
    return float(np.min(ms))


def main():
    for N, dx, P, B in ((353, 0.0125, 768, 64), (353, 0.0125, 768, 16), (177, 0.025, 384, 16), (89, 0.05, 192, 1)):
        p = smm.PDEconstruct(dx=dx)
        assert p.N_x == N
        out = dict(workload=f"{B} x {N}x{N}, P = {P}: W*rho, Kx*rho, Ky*rho and the division", cufft_ms=cufft_conv_ms(N, P, B))
        try:
            out["ours_ms"] = ours_conv_ms(p, B)
            out["ours_over_cufft"] = out["ours_ms"] / out["cufft_ms"]
        except Exception as e:  # noqa: BLE001
            out["ours_error"] = str(e)[:200]
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
