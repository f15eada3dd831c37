#!/bin/bash
# ncu full capture of the fused sweep at N=4e6 (run plain first, as the recipe demands)
B="python bench.py --points 4e6 --steps 2 --warmup 3 --no-e2e --no-cpu"
$B > gpurun_out/plain_prof.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_em16' -s 4 -c 1 -o gpurun_out/prof_r01t $B > gpurun_out/ncu_full_t.log 2>&1
tail -3 gpurun_out/ncu_full_t.log
