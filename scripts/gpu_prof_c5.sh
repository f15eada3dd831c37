#!/bin/bash
B="python bench.py --points 2e6 --dims 8 --comps 256 --cov-type diag --steps 2 --warmup 3 --no-e2e --no-cpu"
$B > gpurun_out/plain_prof_c5.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_emd" -s 4 -c 1 -o gpurun_out/prof_r01t_c5 $B > gpurun_out/ncu_full_tc5.log 2>&1
tail -2 gpurun_out/ncu_full_tc5.log | cut -c1-200
