#!/bin/bash
# ncu full capture of the quadratic-form sweep on config 2 (N=1e6, d=2, K=8) -- plain run first
B="python bench.py --points 1e6 --dims 2 --comps 8 --sigma-c 6 --steps 2 --warmup 3 --no-e2e --no-cpu"
$B > gpurun_out/plain_prof2.log 2>&1 || { tail -5 gpurun_out/plain_prof2.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:'k_em16' -s 8 -c 1 -o gpurun_out/prof_r02l_cfg2 $B > gpurun_out/ncu_full_cfg2.log 2>&1
tail -3 gpurun_out/ncu_full_cfg2.log
