#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_c.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_c.log
tail -25 gpurun_out/pytest_gpu_c.log
python bench.py --points 4e6 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_c_small.json 2> gpurun_out/bench_c_small.err
MLF_B200_LEGACY=1 python bench.py --points 4e6 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_c_small_legacy.json 2> gpurun_out/bench_c_small_legacy.err
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_c_full.json 2> gpurun_out/bench_c_full.err
tail -2 gpurun_out/bench_c_small.err gpurun_out/bench_c_full.err
