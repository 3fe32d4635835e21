"""C4 shard (64 instances 353^2, RK4): one batch of 64 on one stream vs two batches of 32 on two streams of the same GPU
(instances are independent, ensemble.jl:69-70): do the HBM-bound stage kernel of one half and the FFT kernels of the other overlap?"""
import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
import socialmediamodel_b200 as smm
import bench_pde

def make(ctx, qs, p):
    uzy0, _ = smm.constructinitial("4inf", (p, qs[0]))
    return smm.PDEBatch(p, qs, *uzy0.x, precision="f32", shared_init=True, ctx=ctx)

p = smm.PDEconstruct(dx=0.0125)
qs = bench_pde._sweep_qs(smm, 64)
dt = min(smm.default_dt(p, q) for q in qs)
steps = 10
c0 = smm.Context.default()
pb = make(c0, qs, p)
pb.step(dt, 2); c0.sync()
ts = []
for _ in range(3):
    t0 = time.perf_counter(); pb.step(dt, steps); c0.sync(); ts.append((time.perf_counter() - t0) / steps * 1e3)
print("one batch of 64, one stream: ms per RK4 step", round(min(ts), 4))
u1 = pb.read(want_u=False)
pb.close()
for parts in (2, 4):
    ctxs = [smm.Context(device=0) for _ in range(parts)]
    B = 64 // parts
    pbs = [make(c, qs[i * B:(i + 1) * B], p) for i, c in enumerate(ctxs)]
    for b, c in zip(pbs, ctxs): b.step(dt, 2)
    for c in ctxs: c.sync()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        for s in range(steps):                      # interleave the enqueues step by step
            for b in pbs: b.step(dt, 1)
        for c in ctxs: c.sync()
        ts.append((time.perf_counter() - t0) / steps * 1e3)
    print(f"{parts} batches of {B}, {parts} streams: ms per RK4 step", round(min(ts), 4))
    for b in pbs: b.close()
