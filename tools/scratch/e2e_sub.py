import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import bench
import socialmediamodel_b200 as smm
ctx = smm.Context(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
ens = bench.AbmEnsemble(ctx, stream, 1024, 1)
for sub in sys.argv[1:]:
    os.environ["SMM_HOST_WIDEN_SUB"] = sub
    r = ens.time_e2e(3, 1, mode="f64", pinned=False)
    print("sub", sub, "ms", round(r["ms_per_step"], 2), flush=True)
