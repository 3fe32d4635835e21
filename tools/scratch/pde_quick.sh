#!/bin/bash
# PDE parity tests + the three PDE bench workloads, three repetitions (tools/scratch)
timeout 300 python -m pytest tests/test_pde_gpu.py tests/test_pde_adaptive_gpu.py tests/test_control_gpu.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python - <<PY
import sys; sys.path.insert(0, ".")
import socialmediamodel_b200 as smm, bench_pde
ctx = smm.Context.default()
for rep in range(3):
    r = bench_pde.run_pde_workloads(ctx)
    print(rep, "C3 us/step %.1f" % r["pde_C3"]["us_per_step"], "177 ms/step %.4f" % r["pde_sweep_177"]["ms_per_step"],
          "353 ms/step %.3f" % r["pde_C4_shard_353"]["ms_per_step"])
PY
