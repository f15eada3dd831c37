#!/bin/bash
# full GPU suite, then ncu: launch list of a short bench run and one full capture of the quadratic-form sweep (plain run first)
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r02g.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_r02g.log
tail -4 gpurun_out/pytest_gpu_r02g.log
B="python bench.py --points 4e6 --steps 2 --warmup 3 --no-e2e --no-cpu"
$B > gpurun_out/plain_prof.log 2>&1 || { tail -5 gpurun_out/plain_prof.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02g_n4e6.csv $B > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_em16' -s 8 -c 1 -o gpurun_out/prof_r02g $B > gpurun_out/ncu_full_g.log 2>&1
tail -3 gpurun_out/ncu_full_g.log
