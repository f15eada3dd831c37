#!/bin/bash
# GPU call: parity tests, small bench, ncu launch list, ncu full capture of the two hot kernels
B="python bench.py --points 4e6 --steps 2 --warmup 1 --no-e2e --no-cpu"
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
$B > gpurun_out/plain_small.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_small.csv $B > gpurun_out/ncu_launch.log 2>&1
$B > gpurun_out/plain_small2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_estep|k_mstep' -s 2 -c 2 -o gpurun_out/prof_r01a $B > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
