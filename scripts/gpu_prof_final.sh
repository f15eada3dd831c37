#!/bin/bash
# final build: full GPU suite, then ncu --set full of the quadratic-form sweep at the headline shape (N=4e6; plain run first)
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r02p.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_r02p.log; tail -3 gpurun_out/pytest_gpu_r02p.log
B="python bench.py --points 4e6 --steps 2 --warmup 3 --no-e2e --no-cpu"
$B > gpurun_out/plain_prof.log 2>&1 || { tail -5 gpurun_out/plain_prof.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:'k_em16' -s 8 -c 1 -o gpurun_out/prof_r02p $B > gpurun_out/ncu_full_p.log 2>&1
tail -2 gpurun_out/ncu_full_p.log
