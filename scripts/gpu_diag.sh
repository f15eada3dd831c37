#!/bin/bash
# diagonal CUDA-core kernel: parity tests, then the per-GPU shard of config 5 with and without it
python -m pytest tests -m gpu -x -q -k "diagonal or config5 or 64_point" > gpurun_out/pytest_diag.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_diag.log
tail -8 gpurun_out/pytest_diag.log
bash scripts/gpu_cfg5.sh
MLF_B200_EMD=0 python bench.py --points 1.25e8 --dims 8 --comps 256 --cov-type diag --steps 3 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('EMD=0 ms/step %.2f'%j['ms_per_step'], j['roofline']['per_iteration_ms'])"
