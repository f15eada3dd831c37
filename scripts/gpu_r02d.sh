#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r02d.log 2>&1; tail -5 gpurun_out/pytest_gpu_r02d.log
for sc in 1 0.5; do MLF_B200_DEBUG=1 python bench.py --sigma-c $sc --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/dbg_sc$sc.json 2> gpurun_out/dbg_sc$sc.err; grep debug gpurun_out/dbg_sc$sc.err | cut -c1-250 | head -30; done
bash scripts/gpu_overlap.sh
python bench.py --mode infer > gpurun_out/infer_r02d.json 2> gpurun_out/infer_r02d.err; tail -2 gpurun_out/infer_r02d.err; cat gpurun_out/infer_r02d.json
