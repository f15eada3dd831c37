#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_s.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_s.log
tail -5 gpurun_out/pytest_gpu_s.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_s_full.json 2> gpurun_out/bench_s_full.err
tail -3 gpurun_out/bench_s_full.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/bench_s_full.json").read().strip().splitlines()[-1]); r=j["roofline"]
print("ms/step %.2f iter/s %.3f frac %.3f"%(j["ms_per_step"], j["em_iter_per_s"], r["frac"]), r["per_iteration_ms"])
print("e2e", j["e2e"]); print("cpu", j["cpu_baseline"])
PY
