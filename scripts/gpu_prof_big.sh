#!/bin/bash
# ncu full capture of the large-shape kernels at d=128, K=32 (config-4 shape, reduced N)
B="python bench.py --points 4e5 --dims 128 --comps 32 --steps 2 --warmup 1 --no-e2e --no-cpu"
$B > gpurun_out/plain_prof_big.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_estep_big|k_mstep_big' -s 4 -c 2 -o gpurun_out/prof_r01r_big $B > gpurun_out/ncu_full_r.log 2>&1
tail -3 gpurun_out/ncu_full_r.log
bash scripts/gpu_cfg4.sh
