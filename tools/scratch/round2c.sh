#!/bin/bash
# ncu captures of the dense ABM kernel after the step chunks; every ncu run follows a clean run of the same command
python bench.py --no-extra --steps 2 --warmup 1 > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2c_launches_c2.csv python bench.py --no-extra --steps 2 --warmup 1 > gpurun_out/ncu_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:abm_dense -c 1 -o gpurun_out/r2c_abm_dense_chunks -f python bench.py --no-extra --steps 1 --warmup 0 > gpurun_out/ncu_dense3.log 2>&1
