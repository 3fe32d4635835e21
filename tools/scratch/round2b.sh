#!/bin/bash
# bench + ncu captures of the kernels that changed (dense ABM, cluster PDE); every ncu run follows a clean run of the same command
set -x
python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
python bench.py --no-extra --steps 2 --warmup 1 > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2b_launches_c2.csv python bench.py --no-extra --steps 2 --warmup 1 > gpurun_out/ncu_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:abm_dense -c 1 -o gpurun_out/r2b_abm_dense -f python bench.py --no-extra --steps 1 --warmup 0 > gpurun_out/ncu_dense.log 2>&1
python tools/profile_workloads.py pde_c3 > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2b_launches_pde_c3.csv python tools/profile_workloads.py pde_c3 > gpurun_out/ncu_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pde_cluster -c 1 -o gpurun_out/r2b_pde_cluster -f python tools/profile_workloads.py pde_c3 > gpurun_out/ncu_cl.log 2>&1
python tools/profile_workloads.py pde_sweep353x64 > gpurun_out/r2b_sweep353x64.txt 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2b_launches_pde_353x64.csv python tools/profile_workloads.py pde_sweep353x64 > gpurun_out/ncu_353.log 2>&1
