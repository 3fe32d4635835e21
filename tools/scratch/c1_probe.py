import time, numpy as np, os, sys
sys.path.insert(0, "/root/repo")
import socialmediamodel_b200 as smm
ctx = smm.Context.default()
for th in ("128", "256", "512", None):
    if th: os.environ["SMM_ABM_THREADS"] = th
    else: os.environ.pop("SMM_ABM_THREADS", None)
    smm.ABMsolve(100, chosenseed=0, ctx=ctx)
    t0 = time.perf_counter()
    for k in range(20):
        out = smm.ABMsolve(100, chosenseed=k, ctx=ctx)
    wall = (time.perf_counter() - t0) / 20 * 1e3
    print("threads", th, "wall ms/call", round(wall, 3), "kernel ms", round(ctx.kernel_ms(0), 3))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for k in range(20): smm.ABMsolve(100, chosenseed=k, ctx=ctx)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
