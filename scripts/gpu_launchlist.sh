#!/bin/bash
B="python bench.py --points 4e6 --steps 2 --warmup 3 --no-e2e --no-cpu"
$B > gpurun_out/plain_ll.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv $B > gpurun_out/ncu_ll.log 2>&1
tail -2 gpurun_out/ncu_ll.log | cut -c1-300
