"""Small fixed workloads for ncu captures: python tools_profile.py {pde_c3|pde_sweep|pop|kde}"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np

import socialmediamodel_b200 as smm


def main(which):
    ctx = smm.Context.default()
    if which == "pde_c3":
        import bench_pde
        print(bench_pde.run_pde_c3(ctx, steps=20))
    elif which == "pde_sweep":
        import bench_pde
        print(bench_pde.run_pde_sweep(ctx, dx=0.025, B=16, steps=2))
    elif which == "pde_sweep353":
        import bench_pde
        print(bench_pde.run_pde_sweep(ctx, dx=0.0125, B=16, steps=2))
    elif which == "pde_sweep353x64":
        import bench_pde
        print(bench_pde.run_pde_sweep(ctx, dx=0.0125, B=64, steps=2))
    elif which == "pop_long":
        import bench
        print(bench.run_abm_pop(None, ctx, n=10_000_000, steps=20, warmup=3))
    elif which == "pop":
        import bench
        print(bench.run_abm_pop(None, ctx, n=10_000_000, steps=3, warmup=1))


if __name__ == "__main__":
    main(sys.argv[1])
